#!/bin/bash
timeout 600 python -m pytest tests/test_farneback_gpu.py -x -q -m gpu 2>&1 | tail -2
for lanes in 1 2; do timeout 200 python tools/seq_rate.py 3840 2160 3 8 $lanes 2>&1 | tail -1; done | tee gpurun_out/r2f_fb_fast.log
timeout 200 python tools/seq_rate.py 1920 1080 3 16 4 2>&1 | tail -1 | tee -a gpurun_out/r2f_fb_fast.log
timeout 200 python tools/quick_fb.py 3840 2160 2>&1 | tail -1 | tee -a gpurun_out/r2f_fb_fast.log
