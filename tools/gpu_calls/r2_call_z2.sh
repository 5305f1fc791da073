#!/bin/bash
timeout 100 compute-sanitizer --tool racecheck python tools/sanitize_tiny.py fb 2>&1 | tail -2
for lanes in 2 1; do timeout 200 python tools/seq_rate.py 3840 2160 3 8 $lanes 2>&1 | tail -1; done
timeout 200 python -m pytest tests/test_farneback_gpu.py -x -q -m gpu -k "720 or strid or tiny or small" 2>&1 | tail -1
