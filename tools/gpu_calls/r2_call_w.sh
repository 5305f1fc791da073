#!/bin/bash
timeout 600 python -m pytest tests/test_inpaint_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/inpaint_fill_modes.py 2>&1 | tee gpurun_out/r2w_fill_telea_weights.log
