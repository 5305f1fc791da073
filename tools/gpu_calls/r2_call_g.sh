#!/bin/bash
for pf in 2 0 1 3; do for lanes in 1 2; do echo -n "PREFETCH=$pf "; OFXCV_FB_PREFETCH=$pf timeout 200 python tools/seq_rate.py 3840 2160 3 8 $lanes 2>&1 | tail -1; done; done | tee gpurun_out/r2g_fb_pf.log
