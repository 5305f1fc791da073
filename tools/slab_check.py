"""Farneback 4K flow to a .npy (to compare schedules bit for bit). usage: slab_check.py out.npy [W H]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (3840, 2160)
ctx = p.Context(0)
base = s.gray(s.texture(H, W, seed=2000))
nxt = s.shift_bilinear(base, 2.5, -1.5)
np.save(sys.argv[1], ctx.farneback(base, nxt))
