#!/bin/bash
timeout 600 python -m pytest tests/test_farneback_gpu.py -x -q -m gpu 2>&1 | tail -3
r() { echo -n "$1 | "; env $1 timeout 200 python tools/seq_rate.py $3 $4 3 8 $2 2>&1 | tail -1; }
{
r "OFXCV_FB_PIPE=2" 2 3840 2160; r "OFXCV_FB_PIPE=1" 2 3840 2160
r "OFXCV_FB_PIPE=2" 1 3840 2160; r "OFXCV_FB_PIPE=1" 1 3840 2160
r "OFXCV_FB_PIPE=2" 4 1920 1080; r "OFXCV_FB_PIPE=1" 4 1920 1080
} | tee gpurun_out/r2m_fb_pipe.log
