#!/bin/bash
# L1 warm-up distance of the band kernel: variant libraries built with -DFB_WARM_ROWS=1/2/3
for v in 1 2 3 1 2 3; do
  cp tools/tmp_libs/libofxcv_w$v.so openfx-opencv_b200/libofxcv_b200.so
  for lanes in 2 1; do echo -n "WARM_ROWS=$v | "; timeout 200 python tools/seq_rate.py 3840 2160 3 8 $lanes 2>&1 | tail -1; done
done | tee gpurun_out/r2v_fb_warm.log
