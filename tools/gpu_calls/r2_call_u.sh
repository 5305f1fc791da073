#!/bin/bash
timeout 600 ncu --clock-control none --set full --import-source on --kernel-name regex:^fb_band3$ --launch-skip 52 --launch-count 1 -o gpurun_out/r2u_fb_band3 -f python tools/seq_rate.py 3840 2160 3 2 1 > gpurun_out/r2u_ncu.log 2>&1
python tools/ncu_to_profile.py gpurun_out/r2u_fb_band3.ncu-rep gpurun_out/r2u_ncu_fb_band3 3840 2160 | head -32
python tools/ncu_kernel.py gpurun_out/r2u_fb_band3.ncu-rep 8 2>&1 | tail -10
