"""The kernels that changed in round 2, at tiny sizes, for compute-sanitizer --tool racecheck (hazards listed per kernel)."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
p = importlib.import_module("openfx-opencv_b200"); s = importlib.import_module("openfx-opencv_b200.synth")
ctx = p.Context(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "fb"):
    a, b = s.flow_pair(61, 62, seed=3)
    ctx.farneback(a, b, p.FbParams(levels=1, iterations=2))
if which in ("all", "ip"):
    img = s.texture(40, 56, 1)
    for m in (p.INPAINT_NS, p.INPAINT_TELEA):
        ctx.inpaint(img, s.iid_mask(40, 56, 2, 0.2), 3, m)
ctx.close()
print("sanitize_tiny: done")
