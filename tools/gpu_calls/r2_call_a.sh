#!/bin/bash
# one GPU call: phase trace of the plugin renders, the N=1 point of the config 4 / 5 curves, configs 1-3
OFXCV_TRACE=1 timeout 300 python tools/plugin_render_time.py > gpurun_out/r2_plugin_trace.log 2>&1
tools/scale_sweep.sh 1
for wl in farneback_1080p inpaint_telea_vga watershed_4k; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-plugins > gpurun_out/r2_n1_$wl.json 2> gpurun_out/r2_n1_$wl.err
  tail -c 600 gpurun_out/r2_n1_$wl.json
done
