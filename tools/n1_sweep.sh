#!/bin/bash
# every BASELINE.json config on ONE GPU: bench lines into gpurun_out/r2_n1_<workload>.json
for wl in farneback_4k farneback_1080p farneback_8k_l5 inpaint_ns_4k inpaint_telea_vga watershed_4k; do
  timeout 900 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r2_n1_$wl.json 2> gpurun_out/r2_n1_$wl.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_n1_$wl.json"))
    print("$wl value %.2f e2e %.2f frac %.3f parity %s cpu %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d.get("parity", {}).get("ok"), d.get("cpu_baseline", {}).get("value")))
except Exception as e:
    print("$wl FAILED", e)
PY
done
tools/scale_sweep.sh 1
