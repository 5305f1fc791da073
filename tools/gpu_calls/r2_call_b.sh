#!/bin/bash
# incremental inpaint fill: parity, then fill times against the staged kernel; plugin leg stand-alone
timeout 600 python -m pytest tests/test_inpaint_gpu.py -x -q -m gpu > gpurun_out/r2b_inpaint_tests.log 2>&1; echo "inpaint tests rc=$?"
tail -3 gpurun_out/r2b_inpaint_tests.log
for v in 1 3 0; do
  echo "== OFXCV_IP_FILL_INC=$v"; OFXCV_IP_FILL_INC=$v timeout 300 python tools/inpaint_chain.py 2>&1 | tee -a gpurun_out/r2b_chain_inc$v.log
done
timeout 300 python bench.py --plugin-leg > gpurun_out/r2b_plugin_leg.json 2> gpurun_out/r2b_plugin_leg.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_plugin_leg.json'))
for k,v in d.items(): print(k, round(v['ms_per_render'],2))"
OFXCV_XFER_THREADS=8 timeout 300 python bench.py --plugin-leg > gpurun_out/r2b_plugin_leg_t8.json 2> gpurun_out/r2b_plugin_leg_t8.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_plugin_leg_t8.json'))
for k,v in d.items(): print('t8', k, round(v['ms_per_render'],2))"
nproc; lscpu | grep -i "model name\|^CPU(s)\|NUMA"
