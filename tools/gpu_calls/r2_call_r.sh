#!/bin/bash
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2r_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2r_gputests.log
timeout 900 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2r_bench_reference.json 2> gpurun_out/r2r_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/r2r_bench_reference.json | head -c 600
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2r_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'us',d['roofline']['us_per_launch'],'whole',d['roofline']['whole_pair_frac'],'traffic',d['roofline']['traffic'],d['roofline'].get('traffic_source'),'clk',d['clocks'])
for k,v in d['e2e_plugin'].items(): print(k, v.get('ms_per_render'), v.get('passes_ms_per_render'))
for k,v in d['plugins'].items(): print(k, v.get('value'), v.get('bytes_differing_from_cv2', v.get('labels_differing_from_cv2')))
print(d['parity']); print(d['cpu_baseline']); print(d['e2e'])
PY
python -c "import __graft_entry__ as g; g.smoke()"
