#!/bin/bash
{
echo "# compute-sanitizer 12.9, tools/sanitize_small.py with the final kernels of round 2 (incremental inpaint fill, source-colour table, run counting, interior-trip band kernel with the L1 warm-up copies)"
for tool in memcheck racecheck; do
  echo "## --tool $tool"
  timeout 420 compute-sanitizer --tool $tool python tools/sanitize_small.py 2>&1 | grep -v "^$" | tail -4
done
} | tee gpurun_out/r2y_sanitizer.txt
