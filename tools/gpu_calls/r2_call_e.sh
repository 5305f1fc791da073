#!/bin/bash
# L2 policy / band order experiment of the Farneback band kernel
for m in 0 1 3 7 5 2; do for lanes in 1 2; do
  echo -n "OFXCV_FB_L2=$m "; OFXCV_FB_L2=$m timeout 200 python tools/seq_rate.py 3840 2160 3 8 $lanes 2>&1 | tail -1
done; done | tee gpurun_out/r2e_fb_l2.log
OFXCV_FB_L2=7 timeout 600 python -m pytest tests/test_farneback_gpu.py -x -q -m gpu 2>&1 | tail -2
for m in 0 7; do
OFXCV_FB_L2=$m timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --kernel-name regex:^fb_band3$ --launch-skip 50 --launch-count 6 --csv --log-file gpurun_out/r2e_ncu_l2_$m.csv python tools/seq_rate.py 3840 2160 3 2 1 > /dev/null 2>&1
done
python - <<'PY'
import csv
for m in (0,7):
    rows=[r for r in csv.reader(open('gpurun_out/r2e_ncu_l2_%d.csv'%m)) if len(r)>10 and r[0].isdigit()]
    out={}
    for r in rows: out.setdefault(r[0],{})[r[-3]]=r[-1]
    for k,v in out.items(): print(m,k,v)
PY
