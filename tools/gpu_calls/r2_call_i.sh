#!/bin/bash
for q in auto 0 1; do
  if [ $q = auto ]; then timeout 300 python tools/inpaint_sched.py; else OFXCV_IP_READYQ=$q timeout 300 python tools/inpaint_sched.py; fi
done 2>&1 | tee gpurun_out/r2i_inpaint_sched.log
timeout 300 python tools/prof_fb.py 3840 2160 3 4 2>&1 | tee gpurun_out/r2i_prof_fb_4k.log
