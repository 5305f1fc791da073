#!/bin/bash
r() { echo -n "$1 | "; env $1 timeout 200 python tools/seq_rate.py $3 $4 3 8 $2 2>&1 | tail -1; }
{
r "OFXCV_FB_PFL1=0" 2 3840 2160; r "OFXCV_FB_PFL1=1" 2 3840 2160
r "OFXCV_FB_PFL1=0" 1 3840 2160; r "OFXCV_FB_PFL1=1" 1 3840 2160
} | tee gpurun_out/r2p_fb_pfl1.log
