#!/bin/bash
r() { echo -n "$1 | "; env $1 timeout 200 python tools/seq_rate.py $3 $4 3 8 $2 2>&1 | tail -1; }
{
r "OFXCV_FB_OCC=16" 2 3840 2160; r "OFXCV_FB_OCC=20" 2 3840 2160
r "OFXCV_FB_OCC=16" 1 3840 2160; r "OFXCV_FB_OCC=20" 1 3840 2160
r "OFXCV_FB_OCC=16" 4 1920 1080; r "OFXCV_FB_OCC=20" 4 1920 1080
} | tee gpurun_out/r2o_fb_occ.log
