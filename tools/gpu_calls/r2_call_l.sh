#!/bin/bash
# final-code profiles: full capture of the dominant kernel and of the inpaint fill, per-kernel DRAM table, launch list of the bench command
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on --kernel-name regex:^fb_band3$ --launch-skip 52 --launch-count 1 -o gpurun_out/r2z_fb_band3 -f python tools/seq_rate.py 3840 2160 3 2 1 > gpurun_out/r2z_ncu1.log 2>&1
python tools/ncu_to_profile.py gpurun_out/r2z_fb_band3.ncu-rep gpurun_out/r2z_ncu_fb_band3 3840 2160 > /dev/null 2>&1; cat gpurun_out/r2z_ncu_fb_band3.json | head -30
timeout 900 $NCU --set full --import-source on --kernel-name regex:^ip_fill_inc$ --launch-skip 1 --launch-count 1 -o gpurun_out/r2z_ip_fill_inc -f python tools/run_inpaint_once.py 3840 2160 0 > gpurun_out/r2z_ncu2.log 2>&1
python tools/ncu_kernel.py gpurun_out/r2z_ip_fill_inc.ncu-rep 14 > gpurun_out/r2z_ip_fill_inc.txt 2>&1; head -45 gpurun_out/r2z_ip_fill_inc.txt
timeout 900 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r2z_all_kernels.csv python tools/run_all_once.py > gpurun_out/r2z_ncu3.log 2>&1
python tools/ncu_table.py gpurun_out/r2z_all_kernels.csv > gpurun_out/r2z_all_kernels.md 2>&1; head -30 gpurun_out/r2z_all_kernels.md
timeout 900 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file gpurun_out/r2z_launches_farneback_4k.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-plugins --no-parity > gpurun_out/r2z_ncu4.log 2>&1
python tools/ncu_summary.py gpurun_out/r2z_launches_farneback_4k.csv > gpurun_out/r2z_launches_farneback_4k.md 2>&1; cat gpurun_out/r2z_launches_farneback_4k.md
