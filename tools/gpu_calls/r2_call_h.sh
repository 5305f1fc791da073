#!/bin/bash
timeout 600 python -m pytest tests/test_ofx_plugins.py tests/test_convert_gpu.py -x -q -m gpu 2>&1 | tail -2
for nt in 1 0 1 0; do
OFXCV_XFER_NT=$nt timeout 300 python bench.py --plugin-leg > gpurun_out/r2h_plugin_leg_nt$nt.json 2> gpurun_out/r2h_plugin_leg_nt$nt.err; python -c "
import json; d=json.load(open('gpurun_out/r2h_plugin_leg_nt$nt.json'))
print('NT=$nt', {k: (round(v['ms_per_render'],2), v.get('passes_ms_per_render')) for k,v in d.items()})"
done
OFXCV_TRACE=1 timeout 300 python tools/plugin_render_time.py 2>&1 | sed -n 7,13p
