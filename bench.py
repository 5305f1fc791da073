#!/usr/bin/env python
"""bench.py — 4K frames/s of the Farneback optical-flow hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU; torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's OpenCV CPU path, rank 0 only

A "step" = one pass of the hot path over one batch of synthetic 3840x2160 frame pairs per GPU (default 16 pairs,
default plugin parameters: levels 3, winsize 3, 15 iterations, polyN 5, sigma 1.1).  Frames of the sequence are
sharded one block per GPU with no data-path collective (weak scaling: per-GPU work is fixed).
  value  = flow fields (frame pairs) per second, whole job, frames already resident in HBM, CUDA-event timed,
           max over ranks, barrier + synchronize on both sides;
  e2e    = the same metric through the host-buffer C-ABI call (ofxcv_farneback_sequence_u8_host) with page-locked host
           frames: H2D of every frame and D2H of every flow field inside the timed region, every step;
  roofline = the dominant kernel, fb_band3<ITER> at full resolution (14 of the 16 band launches of scale 0):
           algorithmic bytes per launch (88 B per pixel: SURVEY.md 8d) / its average CUDA-event duration inside the
           timed region, against the measured HBM copy peak in MEASURED_PEAKS.json (fallback 6650 GB/s); `traffic` =
           dram bytes read+written per launch from the committed `ncu --set full` capture (profiles/*.json);
  plugins = the other plugin bodies at 4K on the same GPU (rank 0, N=1 only), frames resident in HBM: NS / Telea inpaint
           of a 10 % mask (one frame, and a 32-frame clip with 8 frames in flight through ofxcv_inpaint_sequence_u8),
           watershed with 256 seeds (1 and 512 frames in flight), Dual TV-L1 with the plugin defaults (+ its iteration
           kernel alone); the CPU reference (cv2; for TV-L1 the CPU port on a 1/36-area sample) timed beside them;
  cpu_baseline = the reference arm run once on this box's host cores on a bounded sample (rank 0, N=1 only).
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W4K, H4K = 3840, 2160
METRIC = "4K frames/sec, Farneback optical flow (VectorGenerator plugin body), frames resident in HBM"
UNIT = "frames/s"


def hbm_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json %s)" % k
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


def ncu_traffic(W, H):
    """dram bytes (read + write) per launch of the dominant kernel from the newest committed ncu summary."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_fb_band3.json"))):
        try:
            d = json.load(open(f))
            if d.get("width") == W and d.get("height") == H:
                best = (float(d["dram_bytes_read"]) + float(d["dram_bytes_write"]), os.path.relpath(f, ROOT))
        except Exception:
            pass
    return best if best else (None, None)


# ------------------------------------------------------------------------------------------------ reference arm
def _ref_worker(args):
    """One 4K pair through the reference's CPU body in a worker process (OpenCV is single-threaded here)."""
    kind, seed = args
    import numpy as np
    if seed < 0:  # warm-up: import the library and touch the code path on a small pair (a CPU arm has nothing else to warm)
        a = np.random.default_rng(0).integers(0, 256, (256, 256), dtype=np.uint8)
        if kind == "reference":
            import cv2
            cv2.setNumThreads(1)
            cv2.calcOpticalFlowFarneback(a, a, None, 0.5, 3, 3, 15, 5, 1.1, 0)
        else:
            import oracle
            oracle.farneback(a, a)
        return 0.0
    synth = importlib.import_module("openfx-opencv_b200.synth")
    rng = np.random.default_rng(seed)
    # cheap synthetic pair (workers must not spend their time in the generator): smooth noise + translation
    base = rng.integers(0, 256, (H4K // 8 + 2, W4K // 8 + 2), dtype=np.uint8).astype(np.float32)
    base = np.kron(base, np.ones((8, 8), np.float32))[:H4K + 8, :W4K + 8]
    base = (base[:-8, :-8] + base[8:, 8:] + base[4:-4, 4:-4] * 2) / 4
    prev = base.astype(np.uint8)
    nxt = synth.shift_bilinear(prev, 2.5, -1.5)
    t0 = time.perf_counter()
    if kind == "reference":
        import cv2
        cv2.setNumThreads(1)
        cv2.calcOpticalFlowFarneback(prev, nxt, None, 0.5, 3, 3, 15, 5, 1.1, 0)
    else:
        import oracle
        oracle.farneback(prev, nxt)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    try:
        import cv2  # noqa: F401  (the OpenCV the reference plugin calls; un-vendored dependency, pinned 4.13)
        kind = "reference"
        what = "cv2.calcOpticalFlowFarneback (OpenCV %s)" % cv2.__version__
    except Exception:
        kind = "port"
        what = "oracle/farneback.c (C restatement)"
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, args.ref_workers))
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        def step(i):
            t0 = time.perf_counter()
            pool.map(_ref_worker, [(kind, 1000 * i + k) for k in range(workers)])
            return time.perf_counter() - t0
        for i in range(args.warmup):
            pool.map(_ref_worker, [(kind, -1)] * workers)
        times = [step(100 + i) for i in range(args.steps)]
    total = sum(times)
    value = workers * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "farneback_4k", "width": W4K, "height": H4K, "levels": 3, "iterations": 15, "poly_n": 5,
                   "poly_sigma": 1.1, "winsize": 3, "pairs_per_step": workers},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind,
                         "sample": "%d worker processes x 1 pair 3840x2160 per step, %s, 1 thread each (host has %d cores); warm-up steps run a 256x256 pair per worker" % (workers, what, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    args.out.emit(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------ our arm
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this framework has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    pkg = importlib.import_module("openfx-opencv_b200")
    synth = importlib.import_module("openfx-opencv_b200.synth")
    seq = importlib.import_module("openfx-opencv_b200.sequence")
    ctx = pkg.Context(local)
    par = pkg.FbParams()
    W, H, P = args.width, args.height, args.pairs
    # this rank's shard of the (world*P+1)-frame sequence: contiguous block + one halo frame
    first, count = seq.shard_range(world * P, world, rank)
    base = synth.gray(synth.texture(H, W, seed=2000))
    frames = [synth.shift_bilinear(base, 2.5 * f, -1.5 * f) for f in range(first, first + count + 1)]
    d_frames = ctx.to_device(np.stack(frames))          # the rank's clip block, contiguous in HBM
    d_flows = ctx.alloc(W * H * 8 * count)
    h_frames = [ctx.pinned_array((H, W), np.uint8) for _ in frames]
    for a, b in zip(h_frames, frames):
        a[...] = b
    h_flows = [ctx.pinned_array((H, W, 2), np.float32) for _ in range(count)]
    ext = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one step = one pass over the rank's clip block through the sequence entry points: every frame is blurred and
    # expanded ONCE per pass (frame t+1 of pair t is frame t of pair t+1), nothing is carried from step to step
    def step_resident():
        ctx.farneback_sequence_dev(d_frames.ptr, W, H, count + 1, d_flows.ptr, par)

    def step_e2e():
        ctx.farneback_sequence(h_frames, par, out=h_flows)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(steps):
            fn()
        e1.record(ext)
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        step_resident()
    ctx.synchronize()
    ctx.kernel_time_ms(0)
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launch_count()
    ctx.timing(True)
    ms = timed(step_resident, args.steps)
    ctx.timing(False)
    launches = ctx.launch_count() - l0
    n_iter_situ, iter_ms_situ = ctx.kernel_time_ms(0)   # with the other lane's kernels sharing the GPU
    clocks = sampler.stop() if sampler else None
    # the same K steps without the per-kernel events, to show what the instrumentation costs
    ms_plain = timed(step_resident, args.steps)
    # the dominant kernel on its own: the same steps with ONE pair in flight, so that no other kernel shares the SMs
    ctx.farneback_set_lanes(1)
    step_resident()
    ctx.synchronize()
    ctx.kernel_time_ms(0)
    ctx.timing(True)
    ms_one_lane = timed(step_resident, args.steps)
    ctx.timing(False)
    n_iter, iter_ms = ctx.kernel_time_ms(0)
    ctx.farneback_set_lanes(0)   # back to the default (by frame size: two pairs in flight at 4K)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    lt = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    if rank == 0:
        pairs = world * count * args.steps
        value = pairs / (ms * 1e-3)
        peak, peak_src = hbm_peak()
        launch_bytes = pkg.farneback_iter_bytes(W, H, par)
        achieved = launch_bytes * n_iter / (iter_ms * 1e-3) / 1e9 if iter_ms > 0 else 0.0
        traffic, traffic_src = ncu_traffic(W, H)
        alg = pkg.farneback_algorithmic_bytes(W, H, par)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "farneback_4k" if (W, H) == (W4K, H4K) else "farneback_%dx%d" % (W, H), "width": W, "height": H,
                       "levels": par.levels, "iterations": par.iterations, "poly_n": par.poly_n, "poly_sigma": par.poly_sigma,
                       "winsize": par.winsize, "pairs_per_step": world * count, "sharding": "contiguous frame blocks, one per GPU, 1-frame halo",
                       "call": "ofxcv_farneback_sequence_u8 (one clip block of pairs_per_step/n_gpus + 1 frames per step; each frame's pyramid built once per step)",
                       "l2": "per-pair working set (%.0f MB of M/R/flow planes) exceeds the 126 MB L2; %d distinct pairs rotate" % (
                           W * H * 68 / 1e6, count),
                       "algorithmic_gb_per_pair": alg / 1e9},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "fb_band3<ITER> at %dx%d (full-resolution Farneback iteration launches)" % (W, H), "launches": n_iter,
                         "us_per_launch": 1e3 * iter_ms / max(n_iter, 1), "algorithmic_bytes_per_launch": launch_bytes,
                         "timed": "CUDA events around every such launch over %d steps with one pair in flight (%.1f frames/s in that pass); "
                                  "in the headline pass two pairs are in flight and the same launches share the SMs with the other "
                                  "pair's kernels: %.1f us per launch in situ" % (
                                      args.steps, world * count * args.steps / (ms_one_lane * 1e-3), 1e3 * iter_ms_situ / max(n_iter_situ, 1)),
                         "in_situ_us_per_launch": 1e3 * iter_ms_situ / max(n_iter_situ, 1),
                         "traffic_source": traffic_src, "peak_source": peak_src,
                         "whole_pair_effective_gbs": alg * pairs / world / (ms * 1e-3) / 1e9,
                         "whole_pair_frac": alg * pairs / world / (ms * 1e-3) / 1e9 / peak},
            "e2e": {"value": world * count * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": W * H * (count + 1),
                    "d2h_bytes_per_step": 8 * W * H * count,
                    "api": "ofxcv_farneback_sequence_u8_host, page-locked host frames, upload/compute/download on three streams"},
            "gpu_launches": int(lt.item()),
            "clocks": clocks,
            "value_without_kernel_events": pairs / (ms_plain * 1e-3),
        }
        if world == 1 and not args.no_cpu:
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0",
                                      "--ref-workers", str(args.ref_workers)], capture_output=True, text=True, timeout=600)
                ref = json.loads(out.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = ref["cpu_baseline"]
            except Exception as e:  # the CPU leg must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
        if world == 1 and not args.no_plugins:
            try:
                line["plugins"] = bench_plugins(pkg, synth, ctx, W, H, not args.no_cpu)
            except Exception as e:
                line["plugins"] = {"error": repr(e)}
        args.out.emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


def bench_plugins(pkg, synth, ctx, W, H, with_cpu):
    """The other two plugin bodies at the bench resolution: frames resident in HBM, wall-clock around a synchronise
    (these bodies are many launches each); CPU = cv2 on one host core (they are single-threaded algorithms)."""
    import numpy as np
    out = {}
    img = synth.texture(H, W, seed=4)
    mask = synth.iid_mask(H, W, 1000, 0.10)
    d_img, d_mask, d_out = ctx.to_device(img), ctx.to_device(mask), ctx.alloc(W * H * 3)
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        cv2 = None

    def gpu_time(fn, n):
        fn(); ctx.synchronize()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        ctx.synchronize()
        return (time.perf_counter() - t) / n

    for name, method in (("inpaint_ns", pkg.INPAINT_NS), ("inpaint_telea", pkg.INPAINT_TELEA)):
        dt = gpu_time(lambda: ctx.inpaint_dev(d_img.ptr, 3, d_mask.ptr, d_out.ptr, W, H, 3.0, method), 3)
        ent = {"value": 1 / dt, "unit": "frames/s", "ms_per_frame": dt * 1e3, "workload": "%dx%d RGB8, 10%% iid mask, radius 3" % (W, H),
               "algorithmic_gbs": 7.0 * W * H / dt / 1e9, "bound": "latency (FMM order), not HBM"}
        if cv2 is not None and with_cpu:
            t = time.perf_counter()
            ref = cv2.inpaint(img, mask, 3.0, cv2.INPAINT_NS if method == pkg.INPAINT_NS else cv2.INPAINT_TELEA)
            ent["cpu_frames_per_s"] = 1 / (time.perf_counter() - t)
            got = d_out.download((H, W, 3), np.uint8)
            ent["bytes_differing_from_cv2"] = int((got != ref).sum())
        out[name] = ent
    # a clip of inpaint frames (BASELINE config 4 is a 300-frame sequence): the fill stage is bound by a dependency chain,
    # one frame leaves the GPU mostly idle, so the clip entry point keeps 8 frames in flight (a worker sub-context and
    # host thread each, the persistent fill CTAs split between them)
    K, nper = 8, 4
    mbufs = [ctx.to_device(synth.iid_mask(H, W, 1000 + k, 0.10)) for k in range(K)]
    obufs = [ctx.alloc(W * H * 3) for _ in range(K)]
    imgs = [d_img.ptr] * (K * nper)
    msk = [mbufs[f % K].ptr for f in range(K * nper)]
    outs = [obufs[f % K].ptr for f in range(K * nper)]      # frames f and f+K belong to the same worker: sequential
    ctx.inpaint_sequence_dev(imgs[:K], 3, msk[:K], outs[:K], W, H, 3.0, pkg.INPAINT_NS, K)
    t = time.perf_counter()
    ctx.inpaint_sequence_dev(imgs, 3, msk, outs, W, H, 3.0, pkg.INPAINT_NS, K)
    dt = time.perf_counter() - t
    out["inpaint_ns_8frames"] = {"value": K * nper / dt, "unit": "frames/s", "frames_in_flight": K,
                                 "workload": "%dx%d RGB8, 10%% iid masks, radius 3, %d frames through ofxcv_inpaint_sequence_u8" % (W, H, K * nper),
                                 "algorithmic_gbs": 7.0 * W * H * K * nper / dt / 1e9, "bound": "latency (FMM order), not HBM"}
    for b_ in mbufs + obufs:
        b_.free()
    # Dual TV-L1, the VectorGenerator plugin's second method (default parameters; parity of the method is unpinned)
    base = synth.gray(synth.texture(H, W, seed=2000))
    nxt = synth.shift_bilinear(base, 2.5, -1.5)
    d_a, d_b, d_f = ctx.to_device(base), ctx.to_device(nxt), ctx.alloc(W * H * 8)
    L = pkg.lib()
    dt = gpu_time(lambda: ctx.tvl1_dev(d_a.ptr, d_b.ptr, W, H, d_f.ptr), 3)
    ent = {"value": 1 / dt, "unit": "pairs/s", "ms_per_pair": dt * 1e3, "inner_iterations_run": int(L.ofxcv_tvl1_iterations_run(ctx.h)),
           "workload": "%dx%d gray8 pair, plugin defaults (5 scales, 5 warps, 10 x 15 iterations, epsilon 0.01)" % (W, H)}
    # the iteration kernel alone: every full-resolution launch of a run that cannot stop early
    ctx.timing(True)
    ctx.tvl1_dev(d_a.ptr, d_b.ptr, W, H, d_f.ptr, pkg.Tvl1Params(epsilon=0.0, warps=1, outer_iterations=2, nscales=1))
    ctx.synchronize()
    n_it, ms_it = ctx.kernel_time_ms(1)
    ctx.timing(False)
    if n_it:
        us = ms_it * 1e3 / n_it
        ent["iter_kernel"] = {"name": "tv_iter", "us_per_launch": us, "launches_timed": n_it,
                              "algorithmic_gbs": L.ofxcv_tvl1_iter_bytes(W, H) / us / 1e3, "bytes_per_px": 64}
    if with_cpu:
        import oracle  # CPU port of the same method, timed on a 1/36-area sample (it is a scalar C loop)
        sw, sh = W // 6, H // 6
        sb = synth.gray(synth.texture(sh, sw, seed=2000))
        sn = synth.shift_bilinear(sb, 2.5, -1.5)
        t = time.perf_counter()
        oracle.tvl1(sb, sn)
        ent["cpu_port"] = {"pairs_per_s": 1 / (time.perf_counter() - t), "sample": "%dx%d (1/36 of the area), 1 core" % (sw, sh)}
    out["tvl1"] = ent
    d_a.free(); d_b.free(); d_f.free()
    mk = synth.seed_markers(H, W, 256, 5)
    for nf in (1, 512):
        d_rgbs, d_mks = ctx.alloc(W * H * 3 * nf), ctx.alloc(W * H * 4 * nf)
        L = pkg.lib()
        L.ofxcv_upload(ctx.h, None, d_rgbs.ptr, img.ctypes.data, W * H * 3)
        L.ofxcv_upload(ctx.h, None, d_mks.ptr, mk.ctypes.data, W * H * 4)
        for f in range(1, nf):   # the same frame nf times: replicated on the device
            L.ofxcv_device_copy(ctx.h, None, d_rgbs.ptr + f * W * H * 3, d_rgbs.ptr, W * H * 3)
            L.ofxcv_device_copy(ctx.h, None, d_mks.ptr + f * W * H * 4, d_mks.ptr, W * H * 4)
        ctx.synchronize()
        t = time.perf_counter()
        ctx.watershed_dev(d_rgbs.ptr, d_mks.ptr, W, H, nf)
        ctx.synchronize()
        dt = time.perf_counter() - t
        ent = {"value": nf / dt, "unit": "frames/s", "frames_in_flight": nf, "ms_total": dt * 1e3, "workload": "%dx%d RGB8, 256 seeds" % (W, H),
               "algorithmic_gbs": 11.0 * W * H * nf / dt / 1e9, "bound": "latency (ordered flood), not HBM"}
        if nf == 1 and cv2 is not None and with_cpu:
            m = mk.copy()
            t = time.perf_counter()
            cv2.watershed(img, m)
            ent["cpu_frames_per_s"] = 1 / (time.perf_counter() - t)
            got = np.empty((H, W), np.int32)
            L.ofxcv_download(ctx.h, None, got.ctypes.data, d_mks.ptr, W * H * 4)
            ctx.synchronize()
            ent["labels_differing_from_cv2"] = int((got != m).sum())
        out["watershed_%dframes" % nf] = ent
        d_rgbs.free(); d_mks.free()
    return out


class OneLineStdout:
    """Everything libraries write to fd 1 while the bench runs (NCCL prints its version there) goes to stderr; the JSON
    line is written to the real stdout at the end, so that stdout carries exactly one line."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.real, (text.rstrip("\n") + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=W4K)
    ap.add_argument("--height", type=int, default=H4K)
    ap.add_argument("--pairs", type=int, default=16, help="frame pairs per GPU per step (one clip block of pairs+1 frames)")
    ap.add_argument("--ref-workers", type=int, default=64, help="cap on CPU worker processes of the reference arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-plugins", action="store_true", help="skip the inpaint / watershed lines")
    args = ap.parse_args()
    args.out = OneLineStdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
