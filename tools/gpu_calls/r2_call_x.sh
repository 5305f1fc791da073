#!/bin/bash
for wl in farneback_1080p inpaint_telea_vga; do
  timeout 400 python bench.py --workload $wl --steps 5 --warmup 3 --no-plugins > gpurun_out/r2x_$wl.json 2> gpurun_out/r2x_$wl.err
  python -c "
import json; d=json.load(open('gpurun_out/r2x_$wl.json'))
print('$wl', 'value %.1f e2e %.1f frac %.4f parity %s cpu %.1f clk %s' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity'].get('ok'), d['cpu_baseline']['value'], d['clocks']['sm_mhz']), d.get('single_frame'))"
done
