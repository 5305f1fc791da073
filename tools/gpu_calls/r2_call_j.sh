#!/bin/bash
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2j_gputests.log
timeout 300 python tools/inpaint_sched.py 2>&1 | tee gpurun_out/r2j_inpaint_sched.log
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2j_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'us',d['roofline']['us_per_launch'],'whole',d['roofline']['whole_pair_frac'],'clk',d['clocks'])
for k,v in d['e2e_plugin'].items(): print(k, v.get('ms_per_render'), v.get('passes_ms_per_render'))
for k,v in d['plugins'].items(): print(k, v.get('value'), v.get('bytes_differing_from_cv2', v.get('labels_differing_from_cv2')))
print(d['parity'])
PY
