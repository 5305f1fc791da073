"""One watershed call (ncu target). usage: ws_once.py [W H]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
ctx = p.Context(0)
img = s.texture(H, W, seed=4); mk = s.seed_markers(H, W, 256, 5)
lab = ctx.watershed(img, mk)
print("pops", ctx.watershed_stats())
