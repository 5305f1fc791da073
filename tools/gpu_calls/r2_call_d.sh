#!/bin/bash
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2d_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2d_gputests.log
timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --workload inpaint_ns_4k --steps 3 --warmup 3 --no-cpu > gpurun_out/r2d_c4_n1.json 2> gpurun_out/r2d_c4_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'whole',d['roofline']['whole_pair_frac'],'clk',d['clocks'])
for k,v in d['e2e_plugin'].items(): print(k, v.get('ms_per_render'), v.get('passes_ms_per_render'))
for k,v in d['plugins'].items(): print(k, v.get('value'), v.get('bytes_differing_from_cv2', v.get('labels_differing_from_cv2')))
c=json.load(open('gpurun_out/r2d_c4_n1.json')); print('C4 n1', c['value'], c['e2e']['value'], c['single_frame'], c['roofline']['us_per_launch'])
PY
