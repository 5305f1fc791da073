#!/bin/bash
for w in fb ip; do
  echo "## racecheck $w"
  timeout 100 compute-sanitizer --tool racecheck --print-limit 400 python tools/sanitize_tiny.py $w > gpurun_out/r2z_race_$w.log 2>&1
  grep -c "hazard detected" gpurun_out/r2z_race_$w.log
  grep -A3 "hazard detected" gpurun_out/r2z_race_$w.log | grep -E "hazard detected|in .*\(|at .*cu" | sed 's/0x[0-9a-f]*/ADDR/g; s/thread ([0-9,]*)/thread/g; s/block ([0-9,]*)/block/g' | sort | uniq -c | sort -rn | head -12
  tail -2 gpurun_out/r2z_race_$w.log
done
