#!/bin/bash
r() { echo -n "$1 | "; env $1 timeout 200 python tools/seq_rate.py 3840 2160 3 8 $2 2>&1 | tail -1; }
{
r "X=0" 2; r "X=0" 1
r "OFXCV_FB_OCC=24" 2; r "OFXCV_FB_OCC=24" 1
r "OFXCV_FB_WARPS_PER_SM=16" 2; r "OFXCV_FB_WARPS_PER_SM=12" 2; r "OFXCV_FB_WARPS_PER_SM=6" 2
r "X=0" 3
r "OFXCV_FB_WARPS_PER_SM=20" 1
} | tee gpurun_out/r2k_fb_sweep.log
