"""The C-ABI library loads on a CPU-only box and exports every symbol include/ofxcv_abi.h declares; host-side
planning helpers give the documented numbers; nothing computes without a GPU (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    L = pkg.lib()
    names = pkg.declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), "libofxcv_b200.so does not export " + n
    # and the python binding declares a signature for each of them
    assert set(names) == set(L._ofxcv_sigs)
    assert L.ofxcv_abi_version() == 1


def test_only_abi_symbols_are_exported(pkg):
    out = subprocess.check_output(["nm", "-D", "--defined-only", pkg.LIB_PATH]).decode()
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    extra = [s for s in syms if not s.startswith("ofxcv_") and s not in ("_init", "_fini")]
    assert not extra, extra


def test_planning_helpers(pkg):
    p = pkg.FbParams()
    assert pkg.farneback_scales(1920, 1080, p) == 4
    assert pkg.farneback_scales(7680, 4320, pkg.FbParams(levels=5)) == 6
    assert pkg.farneback_scales(100, 100, p) == 2
    # 3.83 GB per 1080p pair, 15.33 GB per 4K pair, 61.7 GB per 8K/5-level pair
    assert abs(pkg.farneback_algorithmic_bytes(1920, 1080, p) / 1e9 - 3.83) < 0.01
    assert abs(pkg.farneback_algorithmic_bytes(3840, 2160, p) / 1e9 - 15.33) < 0.01
    assert abs(pkg.farneback_algorithmic_bytes(7680, 4320, pkg.FbParams(levels=5)) / 1e9 - 61.7) < 0.1
    L = pkg.lib()
    d = pkg.FbParams(0, 0, 0, 0, 0, 0, 9)
    L.ofxcv_fb_default_params(C.byref(d))
    assert (d.pyr_scale, d.levels, d.winsize, d.iterations, d.poly_n, d.poly_sigma, d.flags) == (0.5, 3, 3, 15, 5, 1.1, 0)
    assert L.ofxcv_status_string(0) == b"ok" and b"device" in L.ofxcv_status_string(-2)
    assert L.ofxcv_farneback_scales(1920, 1080, C.byref(pkg.FbParams(winsize=7))) == -5


def test_no_device_means_no_compute(pkg):
    L = pkg.lib()
    if L.ofxcv_device_count() > 0:
        pytest.skip("a GPU is visible")
    assert L.ofxcv_create(0) is None
    with pytest.raises(pkg.OfxcvError):
        pkg.Context(0)
    # every op refuses a NULL context instead of falling back to the CPU
    assert L.ofxcv_farneback_u8_host(None, None, None, 0, 0, 0, None, 0, None) == -2
    assert L.ofxcv_inpaint_u8_host(None, None, 0, 3, None, 0, None, 0, 0, 0, 3.0, 1) == -2
    assert L.ofxcv_watershed_u8c3_host(None, None, 0, None, 0, 0, 0) == -2


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the package, include/ or the OFX glue may reference it."""
    bad = []
    for base in ("openfx-opencv_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep):
                continue
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".c")):
                    continue
                text = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"^\s*(import|from)\s+oracle\b|libofxcv_oracle|oracle/", text, re.M):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad


@pytest.mark.gpu
def test_concurrent_contexts_are_independent(pkg, synth):
    """kOfxImageEffectRenderFullySafe (VectorGenerator.cpp:108): render threads use separate contexts concurrently; every
    thread must get the bits a lone context gives."""
    import threading
    prev, nxt = synth.flow_pair(270, 480, seed=3)
    img = synth.texture(120, 160, 1)
    mask = synth.iid_mask(120, 160, 2, 0.1)
    mk = synth.seed_markers(120, 160, 9, 5)
    with pkg.Context(0) as c:
        ref = (c.farneback(prev, nxt), c.inpaint(img, mask, 3, pkg.INPAINT_TELEA), c.watershed(img, mk))
    results, errors = {}, []

    def work(i):
        try:
            with pkg.Context(0) as c:
                for _ in range(3):
                    out = (c.farneback(prev, nxt), c.inpaint(img, mask, 3, pkg.INPAINT_TELEA), c.watershed(img, mk))
                results[i] = out
        except Exception as e:  # surfaced below
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i in range(4):
        for a, b in zip(results[i], ref):
            assert np.array_equal(a, b)


def test_median25_network_is_a_median():
    """tv_median5 (Dual TV-L1) finds the 5x5 median with a 99-exchange min/max network.  Zero-one principle: a comparator
    network selects the median of every input iff it does so for all 2^25 inputs of zeros and ones — checked here,
    bit-parallel, on the exact list compiled into csrc/tvl1.cu."""
    import numpy as np
    src = open(os.path.join(ROOT, "openfx-opencv_b200", "csrc", "tvl1.cu")).read()
    body = src[src.index("#define TV_MED25_NET"):src.index("__device__ __forceinline__ float tv_median25")]
    pairs = [(int(a), int(b)) for a, b in re.findall(r"X\((\d+), (\d+)\)", body)]
    assert len(pairs) == 99 and all(0 <= a < b < 25 for a, b in pairs)
    n = 1 << 25
    idx = np.arange(n, dtype=np.uint32)
    wires, ones = [], np.zeros(n, np.uint8)
    for k in range(25):
        bit = ((idx >> k) & 1).astype(np.uint8)
        ones += bit
        wires.append(np.packbits(bit, bitorder="little").view(np.uint64))
    for a, b in pairs:
        wires[a], wires[b] = wires[a] & wires[b], wires[a] | wires[b]     # (min, max)
    expect = np.packbits((ones >= 13).astype(np.uint8), bitorder="little").view(np.uint64)
    assert np.array_equal(wires[12], expect)
