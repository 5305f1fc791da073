#!/bin/bash
r() { echo -n "$1 | "; env $1 timeout 200 python tools/seq_rate.py 3840 2160 3 8 $2 2>&1 | tail -1; }
{
r "OFXCV_FB_PAD_SMEM=0" 2; r "OFXCV_FB_PAD_SMEM=8192" 2; r "OFXCV_FB_PAD_SMEM=16384" 2
r "OFXCV_FB_PAD_SMEM=0" 1; r "OFXCV_FB_PAD_SMEM=16384" 1
r "OFXCV_FB_PIPE=3" 2; r "OFXCV_FB_PIPE=3" 1
} | tee gpurun_out/r2n_fb_l1.log
