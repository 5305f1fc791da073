#!/bin/bash
for inc in 3 2 4; do for wa in 1 2 0; do
  OFXCV_IP_FILL_INC=$inc OFXCV_IP_FILL_WAITALL=$wa timeout 120 python tools/inpaint_fill_modes.py 2>&1 | tee -a gpurun_out/r2c_fill_modes.log
done; done
OFXCV_IP_FILL_INC=0 timeout 120 python tools/inpaint_fill_modes.py 2>&1 | tee -a gpurun_out/r2c_fill_modes.log
for wa in 1 2 0; do OFXCV_IP_FILL_WAITALL=$wa timeout 600 python -m pytest tests/test_inpaint_gpu.py -x -q -m gpu 2>&1 | tail -1; done
