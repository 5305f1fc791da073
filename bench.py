#!/usr/bin/env python
"""bench.py — frames/s of the filter bodies behind openfx-opencv's render actions on N B200s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload NAME]      # our arm (one process per GPU; torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W [--workload NAME]   # the reference's OpenCV CPU path, rank 0 only

Workloads (BASELINE.json configs; the default is the headline the metric is quoted on):
    farneback_4k        3840x2160 Farneback, default plugin parameters (levels 3, winsize 3, 15 iterations, polyN 5, sigma 1.1)
    farneback_1080p     config 2: 1920x1080, same parameters
    farneback_8k_l5     config 5: 7680x4320, levels 5  (--clip-frames 1000 = the 1000-frame sequence, strong scaling)
    inpaint_ns_4k       config 4: Navier-Stokes inpaint, 3840x2160 RGB8, 10 % mask, radius 3 (--clip-frames 300 = the 300-frame sequence)
    inpaint_telea_vga   config 1: Telea inpaint, 640x480, 5 % mask
    watershed_4k        config 3: cv::watershed semantics, 3840x2160, 256 seeds, ONE frame per render (what the segment plugin does)

A "step" = one pass of the hot path over this rank's block of the synthetic sequence.  Sharding: contiguous blocks of
the sequence, one per GPU, one halo frame for flow, no data-path collective (weak scaling: fixed block per GPU; with
--clip-frames the sequence is fixed and split: strong scaling).  Every rank's block holds the same synthetic frames (the
sequence is periodic with the block length), so the per-output checksums of all ranks must be identical and equal to the
1-GPU run's: they are gathered and compared (`checksums`), which is the multi-GPU identity check of SURVEY.md section 4.
  value    = outputs per second, whole job, inputs resident in HBM, CUDA-event timed on the library's stream, max over ranks,
             barrier + synchronize on both sides;
  e2e      = the same through the host-buffer C-ABI call with page-locked host buffers: H2D of every input and D2H of every
             output inside the timed region, every step;
  e2e_plugin (farneback_4k, N=1) = the same through the drop-in .ofx bundle itself: the mini-host drives
             kOfxImageEffectActionRender of VectorGenerator.ofx over a 4K clip of host float RGBA images (a default render is
             TWO pairs: forward and backward flow), plus the inpaint / segment bundles on RGBA8 frames.  Measured first, in
             a process of its own (idle GPU and host cores: one render at a time is a latency measurement); the boxes are
             shared, so every pass over the clip is listed and the fastest pass's median is the figure;
  roofline = the dominant kernel: algorithmic bytes per launch / its average CUDA-event duration inside a timed pass,
             against the measured HBM copy peak in MEASURED_PEAKS.json (fallback 6650 GB/s); `traffic` = dram bytes
             read+written per launch from the newest committed `ncu --set full` summary (profiles/*_ncu_fb_band3.json);
  parity   = this run's output 0 against the reference's OpenCV call (cv2; the C oracle when cv2 is absent) on the SAME frames;
  cpu_baseline = the reference arm run once on this box's host cores on a bounded sample (rank 0, N=1 only).
"""
import argparse
import hashlib
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

UNIT = "frames/s"
WORKLOADS = {
    "farneback_4k": dict(kind="flow", W=3840, H=2160, levels=3, block=16, chunk=16, ref_per_worker=1,
                         metric="4K frames/sec, Farneback optical flow (VectorGenerator plugin body), frames resident in HBM"),
    "farneback_1080p": dict(kind="flow", W=1920, H=1080, levels=3, block=32, chunk=32, ref_per_worker=2,
                            metric="1080p frames/sec, Farneback optical flow (VectorGenerator plugin body), frames resident in HBM"),
    "farneback_8k_l5": dict(kind="flow", W=7680, H=4320, levels=5, block=8, chunk=4, ref_per_worker=1,
                            metric="8K frames/sec, Farneback optical flow with 5 pyramid levels (VectorGenerator plugin body), frames resident in HBM"),
    "inpaint_ns_4k": dict(kind="inpaint", W=3840, H=2160, method="ns", frac=0.10, block=32, ref_per_worker=1,
                          metric="4K frames/sec, Navier-Stokes inpaint (inpaint plugin body), frames resident in HBM"),
    "inpaint_telea_vga": dict(kind="inpaint", W=640, H=480, method="telea", frac=0.05, block=64, ref_per_worker=16,
                              metric="640x480 frames/sec, Telea inpaint (inpaint plugin body), frames resident in HBM"),
    "watershed_4k": dict(kind="watershed", W=3840, H=2160, seeds=256, block=1, ref_per_worker=4,
                         metric="4K frames/sec, marker watershed (segment plugin body), one frame per render, frame resident in HBM"),
}


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


# ------------------------------------------------------------------------------------------------ synthetic inputs
def flow_frames(synth, W, H, n):
    """Frame t of every rank's block: Texture(seed 2000) translated by t * (2.5, -1.5) px."""
    base = synth.gray(synth.texture(H, W, seed=2000))
    return [base if t == 0 else synth.shift_bilinear(base, 2.5 * t, -1.5 * t) for t in range(n)]


def inpaint_inputs(synth, W, H, frac, n):
    img = synth.texture(H, W, seed=4)
    return img, [synth.iid_mask(H, W, 1000 + t, frac) for t in range(n)]


def watershed_inputs(synth, W, H, seeds):
    return synth.texture(H, W, seed=4), synth.seed_markers(H, W, seeds, 5)


def digest(keys):
    return hashlib.sha1(",".join("%016x" % k for k in keys).encode()).hexdigest()[:16]


# ------------------------------------------------------------------------------------------------ reference arm
def _ref_worker(args):
    """`per` units of a workload through the reference's CPU body in a worker process (OpenCV runs these single-threaded)."""
    name, kind, seed, per = args
    import numpy as np
    wl = WORKLOADS[name]
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        cv2 = None
    import oracle
    synth = importlib.import_module("openfx-opencv_b200.synth")
    warm = seed < 0
    W, H = (256, 256) if warm else (wl["W"], wl["H"])
    if wl["kind"] == "flow":
        fr = flow_frames(synth, W, H, 2)
        t0 = time.perf_counter()
        for _ in range(1 if warm else per):
            if cv2 is not None:
                cv2.calcOpticalFlowFarneback(fr[0], fr[1], None, 0.5, wl["levels"], 3, 15, 5, 1.1, 0)
            else:
                oracle.farneback(fr[0], fr[1], levels=wl["levels"])
    elif wl["kind"] == "inpaint":
        img, masks = inpaint_inputs(synth, W, H, wl["frac"], 1)
        t0 = time.perf_counter()
        for _ in range(1 if warm else per):
            if cv2 is not None:
                cv2.inpaint(img, masks[0], 3.0, cv2.INPAINT_NS if wl["method"] == "ns" else cv2.INPAINT_TELEA)
            else:
                oracle.inpaint(img, masks[0], 3, oracle.INPAINT_NS if wl["method"] == "ns" else oracle.INPAINT_TELEA)
    else:
        img, mk = watershed_inputs(synth, W, H, 4 if warm else wl["seeds"])
        t0 = time.perf_counter()
        for _ in range(1 if warm else per):
            if cv2 is not None:
                cv2.watershed(img, mk.copy())
            else:
                oracle.watershed(img, mk)
    return time.perf_counter() - t0


def reference_call(wl):
    try:
        import cv2
        v = "OpenCV %s" % cv2.__version__
        kind = "reference"
    except Exception:
        v, kind = "C restatement in oracle/", "port"
    name = {"flow": "calcOpticalFlowFarneback", "inpaint": "inpaint", "watershed": "watershed"}[wl["kind"]]
    return kind, ("cv2.%s (%s)" % (name, v)) if kind == "reference" else ("oracle %s (%s)" % (name, v))


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    wl = WORKLOADS[args.workload]
    kind, what = reference_call(wl)
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, args.ref_workers))
    per = wl["ref_per_worker"]
    ctx = mp.get_context("spawn")
    with ctx.Pool(workers) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_worker, [(args.workload, kind, -1, 1)] * workers)
        times = []
        for i in range(args.steps):
            t0 = time.perf_counter()
            pool.map(_ref_worker, [(args.workload, kind, 1000 * i + k, per) for k in range(workers)])
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = workers * per * args.steps / total
    line = {
        "impl": "reference", "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if wl["kind"] == "flow" else "u8" if wl["kind"] == "inpaint" else "i32", "data": "synthetic",
        "config": workload_config(args.workload, wl, workers * per),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind,
                         "sample": "%d worker processes x %d unit(s) %dx%d per step, %s, 1 thread each (host has %d cores); warm-up steps run a 256x256 case per worker" % (
                             workers, per, wl["W"], wl["H"], what, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    args.out.emit(json.dumps(line))
    return 0


def workload_config(name, wl, per_step):
    c = {"workload": name, "width": wl["W"], "height": wl["H"]}
    if wl["kind"] == "flow":
        c.update({"levels": wl["levels"], "iterations": 15, "poly_n": 5, "poly_sigma": 1.1, "winsize": 3, "pairs_per_step": per_step})
    elif wl["kind"] == "inpaint":
        c.update({"method": wl["method"], "radius": 3, "mask_fraction": wl["frac"], "frames_per_step": per_step})
    else:
        c.update({"seeds": wl["seeds"], "frames_per_step": per_step})
    return c


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


class Rig:
    """What every workload needs: the rank's context, distributed plumbing, the timed-region protocol."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — this framework has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            import datetime
            # a collective that does not complete within minutes is a bug of this script: fail fast instead of hanging the box
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local), timeout=datetime.timedelta(seconds=300))
        self.pkg = importlib.import_module("openfx-opencv_b200")
        self.synth = importlib.import_module("openfx-opencv_b200.synth")
        self.seq = importlib.import_module("openfx-opencv_b200.sequence")
        self.ctx = self.pkg.Context(self.local)
        self.ext = torch.cuda.ExternalStream(self.ctx.stream(), device=torch.device("cuda", self.local))
        self.args = args

    def barrier(self):
        self.ctx.synchronize()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, blocking=False):
        """K steps bracketed by barrier + synchronize, CUDA events on the library's stream, MAX over ranks (ms).
        blocking=True: fn synchronises internally (round loops with host read-backs): the events still bracket it."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.ext)
        for _ in range(steps):
            fn()
        e1.record(self.ext)
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        self.barrier()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def total(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather(self, obj):
        return self.seq.gather_results(obj)

    def share(self, count_total):
        """(first, count) of this rank's block: a fixed block per GPU (weak) or a split of --clip-frames outputs (strong)."""
        if count_total is None:
            return None
        return self.seq.shard_range(count_total, self.world, self.rank)

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()
        self.ctx.close()


def checksum_report(rig, keys):
    """Per-output content keys of this rank's block, gathered: every rank holds the same synthetic block, so every rank's
    list must be a prefix of the longest one — and the digest must not depend on the number of GPUs."""
    allk = rig.gather(keys)
    longest = max(allk, key=len)
    same = all(k == longest[:len(k)] for k in allk)
    return {"identical_across_ranks": bool(same), "ranks": len(allk), "outputs_compared": min(len(k) for k in allk),
            "digest_first_block": digest(longest), "key_output_0": "%016x" % longest[0] if longest else None, "how": "ofxcv_content_key_u8 over the bytes of every output of the rank's first call"}


def cpu_baseline_leg(args, line):
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "1",
                              "--warmup", "0", "--ref-workers", str(args.ref_workers)], capture_output=True, text=True, timeout=900)
        ref = json.loads(out.stdout.strip().splitlines()[-1])
        line["cpu_baseline"] = ref["cpu_baseline"]
    except Exception as e:  # the CPU leg must never take the GPU number down with it
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}


# ---- dense optical flow ---------------------------------------------------------------------------------------
def plugin_leg_first(args):
    """e2e_plugin: single renders through the .ofx bundles, measured BEFORE this process creates its own context, in a
    process of its own (an idle GPU and idle host cores: it is a latency measurement of one render at a time, and the
    host-side row copies of the glue must not share cores with this process's thread pools)."""
    if int(os.environ.get("WORLD_SIZE", "1")) != 1 or args.workload != "farneback_4k" or args.no_plugins:
        return None
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--plugin-leg"], capture_output=True, text=True, timeout=900)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {"error": repr(e)}


def run_flow(args, wl):
    import numpy as np
    e2e_plugin = plugin_leg_first(args)
    rig = Rig(args)
    pkg, synth, ctx, rank, world = rig.pkg, rig.synth, rig.ctx, rig.rank, rig.world
    W, H = wl["W"], wl["H"]
    par = pkg.FbParams(levels=wl["levels"])
    strong = args.clip_frames > 0
    count = rig.share(args.clip_frames - 1)[1] if strong else (args.pairs or wl["block"])
    chunk = min(count, args.chunk or wl["chunk"])          # pairs per clip call (bounded by the flow buffer in HBM)
    calls = [min(chunk, count - lo) for lo in range(0, count, chunk)]
    frames = flow_frames(synth, W, H, chunk + 1)
    d_frames = ctx.to_device(np.stack(frames))            # the rank's clip block, contiguous in HBM
    d_flows = ctx.alloc(W * H * 8 * chunk)
    h_frames = [ctx.pinned_array((H, W), np.uint8) for _ in frames]
    for a, b in zip(h_frames, frames):
        a[...] = b
    h_flows = [ctx.pinned_array((H, W, 2), np.float32) for _ in range(chunk)]

    # one step = one pass over the rank's block through the clip entry points: every frame of a call is blurred and
    # expanded ONCE (frame t+1 of pair t is frame t of pair t+1), nothing is carried from call to call or step to step
    def step_resident():
        for n in calls:
            ctx.farneback_sequence_dev(d_frames.ptr, W, H, n + 1, d_flows.ptr, par)

    def step_e2e():
        for n in calls:
            ctx.farneback_sequence(h_frames[:n + 1], par, out=h_flows[:n])

    for _ in range(max(args.warmup, 3)):
        step_resident()
    ctx.synchronize()
    ctx.kernel_time_ms(0)
    sampler = ClockSampler(rig.local) if rank == 0 else None
    l0 = ctx.launch_count()
    ctx.timing(True)
    ms = rig.timed(step_resident, args.steps)
    ctx.timing(False)
    launches = ctx.launch_count() - l0
    n_iter_situ, iter_ms_situ = ctx.kernel_time_ms(0)   # with the other lanes' kernels sharing the GPU
    clocks = sampler.stop() if sampler else None
    # every call of a step writes the same block of flow fields (periodic sequence), so after the step slot i holds pair i of
    # the block whichever call wrote it last: the keys cover one full call and do not depend on how many ranks share the clip
    keys = [ctx.content_key(d_flows.ptr + i * W * H * 8, W * 8, H) for i in range(max(calls))] if calls else []
    flow0 = d_flows.download((1, H, W, 2), np.float32)[0] if (rank == 0 and world == 1 and calls[-1] >= 1) else None
    # the same K steps without the per-kernel events, to show what the instrumentation costs
    ms_plain = rig.timed(step_resident, args.steps)
    # the dominant kernel on its own: the same steps with ONE pair in flight, so that no other kernel shares the SMs
    ctx.farneback_set_lanes(1)
    step_resident()
    ctx.synchronize()
    ctx.kernel_time_ms(0)
    ctx.timing(True)
    ms_one_lane = rig.timed(step_resident, max(1, args.steps // 2))
    ctx.timing(False)
    n_iter, iter_ms = ctx.kernel_time_ms(0)
    ctx.farneback_set_lanes(0)   # back to the default (by frame size: two pairs in flight at 4K)
    for _ in range(2):
        step_e2e()
    ms_e2e = rig.timed(step_e2e, args.steps)

    # the ceiling the host fabric puts on e2e: the same H2D and D2H bytes, same page-locked buffers, NO compute, uploads on the
    # library's stream and downloads on its staging stream (full duplex), all ranks at once
    L = pkg.lib()
    aux = L.ofxcv_aux_stream(ctx.h)

    def step_copies():
        for n in calls:
            for t in range(n + 1):
                L.ofxcv_upload(ctx.h, None, d_frames.ptr + t * W * H, h_frames[t].ctypes.data, W * H)
            for t in range(n):
                L.ofxcv_download(ctx.h, aux, h_flows[t].ctypes.data, d_flows.ptr + t * W * H * 8, W * H * 8)
        L.ofxcv_stream_wait(ctx.h, None, aux)     # the timing events sit on the main stream

    step_copies()
    ms_copy = rig.timed(step_copies, args.steps)
    # every collective happens here, on every rank; below this point only rank 0 works
    count_all = int(rig.total(count))
    pairs_all = count_all * args.steps
    launches_all = rig.total(launches)
    h2d_all = int(rig.total(sum(W * H * (n + 1) for n in calls)))
    d2h_all = int(rig.total(sum(8 * W * H * n for n in calls)))
    sums = checksum_report(rig, keys)
    if rank == 0:
        value = pairs_all / (ms * 1e-3)
        peak, peak_src = hbm_peak()
        launch_bytes = pkg.farneback_iter_bytes(W, H, par)
        achieved = launch_bytes * n_iter / (iter_ms * 1e-3) / 1e9 if iter_ms > 0 else 0.0
        traffic, traffic_src = ncu_traffic(W, H)
        # byte model: a pair = 2 frame pyramids + one solve; a clip call of n pairs builds n+1 pyramids
        alg = lambda it: pkg.farneback_algorithmic_bytes(W, H, pkg.FbParams(levels=wl["levels"], iterations=it))
        a1, a2 = alg(1), alg(2)
        sum_n = (a2 - a1) / 88.0
        frame_bytes = ((a1 - 88.0 * sum_n) - 10.0 * sum_n) / 2.0
        solve_bytes = (88.0 * par.iterations + 10.0) * sum_n
        pair_model = alg(par.iterations)
        step_bytes = sum((n + 1) * frame_bytes + n * solve_bytes for n in calls)     # what this rank's step really has to move
        line = {
            "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic: Texture(seed 2000) translated by t*(2.5,-1.5) px; every rank's block holds frames t = 0..%d (periodic sequence)" % chunk,
            "config": dict(workload_config(args.workload, wl, count_all), **{
                "sharding": "contiguous frame blocks, one per GPU, 1-frame halo" + (" (clip of %d frames split over the ranks)" % args.clip_frames if strong else ""),
                "call": "ofxcv_farneback_sequence_u8, %s pair(s) + 1 frames per call, %d call(s) per step and GPU; each frame's pyramid built once per call" % (
                    "/".join(str(n) for n in sorted(set(calls), reverse=True)), len(calls)),
                "l2": "per-pair working set (%.0f MB of M/R/flow planes) exceeds the 126 MB L2; %d distinct pairs rotate" % (W * H * 68 / 1e6, chunk),
                "algorithmic_gb_per_pair": pair_model / 1e9}),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "fb_band3<ITER> at %dx%d (full-resolution Farneback iteration launches)" % (W, H), "launches": n_iter,
                         "us_per_launch": 1e3 * iter_ms / max(n_iter, 1), "algorithmic_bytes_per_launch": launch_bytes,
                         "timed": "CUDA events around every such launch over %d steps with one pair in flight (%.1f frames/s in that pass); "
                                  "in the headline pass several pairs are in flight and the same launches share the SMs with the other "
                                  "pairs' kernels: %.1f us per launch in situ" % (
                                      max(1, args.steps // 2), pairs_all / args.steps * max(1, args.steps // 2) / (ms_one_lane * 1e-3),
                                      1e3 * iter_ms_situ / max(n_iter_situ, 1)),
                         "in_situ_us_per_launch": 1e3 * iter_ms_situ / max(n_iter_situ, 1),
                         "traffic_source": traffic_src, "peak_source": peak_src,
                         "whole_pair_effective_gbs": step_bytes * args.steps / (ms * 1e-3) / 1e9,
                         "whole_pair_frac": step_bytes * args.steps / (ms * 1e-3) / 1e9 / peak,
                         "whole_pair_model": "per GPU: sum over calls of (n+1) frame pyramids x %.3f GB + n solves x %.3f GB (a stand-alone pair = %.3f GB)" % (
                             frame_bytes / 1e9, solve_bytes / 1e9, pair_model / 1e9)},
            "e2e": {"value": pairs_all / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
                    "api": "ofxcv_farneback_sequence_u8_host, page-locked host frames, upload/compute/download on three streams",
                    "copy_ceiling": {"value": pairs_all / (ms_copy * 1e-3), "unit": UNIT,
                                     "gb_per_s_per_gpu": (h2d_all + d2h_all) / max(world, 1) * args.steps / (ms_copy * 1e-3) / 1e9,
                                     "what": "the same H2D + D2H bytes from / to the same page-locked buffers with no compute, both directions at once, all ranks at once"},
                    "frac_of_copy_ceiling": (pairs_all / (ms_e2e * 1e-3)) / (pairs_all / (ms_copy * 1e-3))},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "checksums": sums,
            "value_without_kernel_events": pairs_all / (ms_plain * 1e-3),
        }
        if world == 1 and flow0 is not None and not args.no_parity:
            line["parity"] = flow_parity(frames[0], frames[1], flow0, wl["levels"])
        if world == 1 and not args.no_cpu:
            cpu_baseline_leg(args, line)
        if world == 1 and args.workload == "farneback_4k" and not args.no_plugins:
            line["e2e_plugin"] = e2e_plugin
            try:
                line["plugins"] = bench_plugins(pkg, synth, ctx, W, H, not args.no_cpu)
            except Exception as e:
                line["plugins"] = {"error": repr(e)}
        args.out.emit(json.dumps(line))
    rig.finish()
    return 0


def flow_parity(prev, nxt, got, levels):
    """This run's flow field 0 against the reference's OpenCV call on the same two frames (SURVEY.md 8c metrics)."""
    import numpy as np
    try:
        import cv2
        cv2.setNumThreads(max(1, (os.cpu_count() or 1)))
        ref = cv2.calcOpticalFlowFarneback(prev, nxt, None, 0.5, levels, 3, 15, 5, 1.1, 0)
        against = "cv2.calcOpticalFlowFarneback (OpenCV %s)" % cv2.__version__
    except Exception:
        import oracle
        ref = oracle.farneback(prev, nxt, levels=levels)
        against = "oracle/farneback.c"
    d = np.abs(got - ref).max(axis=2)
    return {"against": against, "mean_abs": float(d.mean()), "frac_gt_1e-2": float((d > 1e-2).mean()), "frac_gt_1": float((d > 1).mean()),
            "tolerance": "mean <= 1e-3, frac(>1e-2) <= 1e-3, frac(>1) <= 2e-4",
            "ok": bool(d.mean() <= 1e-3 and (d > 1e-2).mean() <= 1e-3 and (d > 1).mean() <= 2e-4)}


# ---- inpaint --------------------------------------------------------------------------------------------------
def run_inpaint(args, wl):
    import numpy as np
    rig = Rig(args)
    pkg, synth, ctx, rank, world = rig.pkg, rig.synth, rig.ctx, rig.rank, rig.world
    W, H = wl["W"], wl["H"]
    method = pkg.INPAINT_NS if wl["method"] == "ns" else pkg.INPAINT_TELEA
    strong = args.clip_frames > 0
    count = rig.share(args.clip_frames)[1] if strong else (args.pairs or wl["block"])
    K = 8                                                  # frames in flight (worker sub-contexts of the clip entry point)
    nd = min(count, 2 * K)                                 # distinct masks / output buffers; frame f uses slot f % nd
    img, masks = inpaint_inputs(synth, W, H, wl["frac"], nd)
    d_img = ctx.to_device(img)
    d_masks = [ctx.to_device(m) for m in masks]
    d_outs = [ctx.alloc(W * H * 3) for _ in range(nd)]
    imgs = [d_img.ptr] * count
    msk = [d_masks[f % nd].ptr for f in range(count)]
    outs = [d_outs[f % nd].ptr for f in range(count)]
    h_img = ctx.pinned_array((H, W, 3), np.uint8); h_img[...] = img
    h_masks = []
    for m in masks:
        a = ctx.pinned_array((H, W), np.uint8); a[...] = m; h_masks.append(a)

    def step_resident():
        ctx.inpaint_sequence_dev(imgs, 3, msk, outs, W, H, 3.0, method, K)

    def step_e2e():
        ctx.inpaint_sequence([h_img] * count, [h_masks[f % nd] for f in range(count)], 3.0, method, K)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    ctx.synchronize()
    sampler = ClockSampler(rig.local) if rank == 0 else None
    l0 = ctx.launch_count()
    ms = rig.timed(step_resident, args.steps, blocking=True)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    keys = [ctx.content_key(d_outs[f].ptr, W * 3, H) for f in range(nd)]
    out0 = d_outs[0].download((H, W, 3), np.uint8) if rank == 0 and world == 1 else None
    # one frame at a time: what ONE render of the inpaint plugin costs; its fill kernel is the dominant kernel
    ctx.kernel_time_ms(1)
    ctx.timing(True)
    t = time.perf_counter()
    nsingle = 4
    for f in range(nsingle):
        ctx.inpaint_dev(d_img.ptr, 3, d_masks[f % nd].ptr, d_outs[f % nd].ptr, W, H, 3.0, method)
    ctx.synchronize()
    single_ms = (time.perf_counter() - t) * 1e3 / nsingle
    ctx.timing(False)
    n_fill, fill_ms = ctx.kernel_time_ms(1)
    step_e2e()
    ms_e2e = rig.timed(step_e2e, args.steps, blocking=True)
    count_all = int(rig.total(count))   # every collective on every rank; below only rank 0 works
    frames_all = count_all * args.steps
    launches_all = rig.total(launches)
    sums = checksum_report(rig, keys)
    if rank == 0:
        peak, peak_src = hbm_peak()
        px_bytes = 7.0 * W * H                                # RGB8 in, mask, RGB8 out
        achieved = px_bytes * n_fill / (fill_ms * 1e-3) / 1e9 if fill_ms > 0 else 0.0
        line = {
            "metric": wl["metric"], "value": frames_all / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "u8",
            "data": "synthetic: Texture(seed 4) RGB8, iid masks (seeds 1000+t), every rank's block holds the same %d distinct frames" % nd,
            "config": dict(workload_config(args.workload, wl, count_all), **{
                "sharding": "contiguous frame blocks, one per GPU, no halo" + (" (clip of %d frames split over the ranks)" % args.clip_frames if strong else ""),
                "call": "ofxcv_inpaint_sequence_u8, %d frames per call, %d in flight" % (count, K),
                "l2": "a 4K frame's working set (T map, flags, colours: ~100 MB per frame in flight) exceeds the 126 MB L2 with 8 frames in flight" if W * H > 4e6
                      else "small frames: the working set of 8 frames in flight fits L2; the path is bound by the FMM dependency chain, not by memory"}),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "ip_fill_inc (one frame at a time: %d launches timed with CUDA events)" % n_fill, "launches": n_fill,
                         "us_per_launch": 1e3 * fill_ms / max(n_fill, 1), "algorithmic_bytes_per_launch": px_bytes, "peak_source": peak_src,
                         "note": "the fill is bound by the dependency chain of the fast-marching order (thousands of steps deep), not by HBM: "
                                 "the fraction is reported because the contract asks for it, it is not a target"},
            "single_frame": {"ms": single_ms, "frames_per_s": 1e3 / single_ms, "call": "ofxcv_inpaint_u8 (what one render of the plugin runs)"},
            "e2e": {"value": frames_all / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": count_all * W * H * 4,
                    "d2h_bytes_per_step": count_all * W * H * 3,
                    "api": "ofxcv_inpaint_sequence_u8_host, page-locked host frames, uploads / downloads overlap the other frames' compute"},
            "gpu_launches": int(launches_all), "clocks": clocks, "checksums": sums,
        }
        if world == 1 and out0 is not None and not args.no_parity:
            try:
                import cv2
                cv2.setNumThreads(1)
                t = time.perf_counter()
                ref = cv2.inpaint(img, masks[0], 3.0, cv2.INPAINT_NS if wl["method"] == "ns" else cv2.INPAINT_TELEA)
                line["parity"] = {"against": "cv2.inpaint (OpenCV %s)" % cv2.__version__, "bytes_differing": int((out0 != ref).sum()),
                                  "ok": bool((out0 == ref).all()), "cv2_ms_this_frame": (time.perf_counter() - t) * 1e3}
            except Exception as e:
                line["parity"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu:
            cpu_baseline_leg(args, line)
        args.out.emit(json.dumps(line))
    rig.finish()
    return 0


# ---- watershed ------------------------------------------------------------------------------------------------
def run_watershed(args, wl):
    import numpy as np
    rig = Rig(args)
    pkg, synth, ctx, rank, world = rig.pkg, rig.synth, rig.ctx, rig.rank, rig.world
    W, H = wl["W"], wl["H"]
    count = args.pairs or wl["block"]
    img, mk = watershed_inputs(synth, W, H, wl["seeds"])
    d_rgb = ctx.to_device(np.stack([img] * count))
    d_mk0 = ctx.to_device(np.stack([mk] * count))          # pristine markers (the flood works in place)
    d_mk = ctx.alloc(W * H * 4 * count)
    L = pkg.lib()
    h_rgb = ctx.pinned_array((H, W, 3), np.uint8); h_rgb[...] = img
    h_mk = ctx.pinned_array((H, W), np.int32)

    def step_resident():
        L.ofxcv_device_copy(ctx.h, None, d_mk.ptr, d_mk0.ptr, W * H * 4 * count)
        ctx.watershed_dev(d_rgb.ptr, d_mk.ptr, W, H, count)

    def step_e2e():
        for _ in range(count):
            h_mk[...] = mk
            st = L.ofxcv_watershed_u8c3_host(ctx.h, h_rgb.ctypes.data, W * 3, h_mk.ctypes.data, W * 4, W, H)
            assert st == 0, st

    for _ in range(max(args.warmup, 3)):
        step_resident()
    ctx.synchronize()
    sampler = ClockSampler(rig.local) if rank == 0 else None
    l0 = ctx.launch_count()
    ms = rig.timed(step_resident, args.steps, blocking=True)
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    stats = (pkg.C.c_int64 * 4)()
    L.ofxcv_watershed_last_stats(ctx.h, stats)
    keys = [ctx.content_key(d_mk.ptr + f * W * H * 4, W * 4, H) for f in range(count)]
    lab0 = d_mk.download((count, H, W), np.int32)[0] if rank == 0 and world == 1 else None
    step_e2e()
    ms_e2e = rig.timed(step_e2e, args.steps, blocking=True)
    count_all = int(rig.total(count))   # every collective on every rank; below only rank 0 works
    frames_all = count_all * args.steps
    launches_all = rig.total(launches)
    sums = checksum_report(rig, keys)
    if rank == 0:
        peak, peak_src = hbm_peak()
        px_bytes = 11.0 * W * H * count                        # RGB8 in, int32 markers in + out
        achieved = px_bytes * args.steps / (ms * 1e-3) / 1e9
        line = {
            "metric": wl["metric"], "value": frames_all / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "i32", "data": "synthetic: Texture(seed 4) RGB8, %d 5x5 seed squares (seed 5), the same frame on every rank" % wl["seeds"],
            "config": dict(workload_config(args.workload, wl, count_all), **{
                "sharding": "one frame per GPU per step (frames of a sequence are independent)",
                "call": "ofxcv_watershed_u8c3_batch with %d frame(s): the exact intra-frame parallel flood (watershed_par.cu), %d rounds / %d passes for this frame" % (
                    count, stats[2], stats[3]),
                "l2": "claim words (66 MB) + colour plane (33 MB) + record pool exceed the 126 MB L2; the pristine markers are copied in again every step"}),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "whole flood (wsp_run / wsp_validate / wsp_commit rounds)", "launches": int(launches / max(args.steps, 1)),
                         "algorithmic_bytes_per_launch": px_bytes, "peak_source": peak_src,
                         "note": "an ordered flood is bound by the length of its dependency chains (the longest sub-flood of every round), "
                                 "not by HBM: the fraction is reported because the contract asks for it, it is not a target"},
            "pops_per_frame": int(stats[0] // max(count, 1)),
            "e2e": {"value": frames_all / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": count_all * W * H * 7,
                    "d2h_bytes_per_step": count_all * W * H * 4, "api": "ofxcv_watershed_u8c3_host, page-locked host frame and markers"},
            "gpu_launches": int(launches_all), "clocks": clocks, "checksums": sums,
        }
        if world == 1 and lab0 is not None and not args.no_parity:
            try:
                import cv2
                cv2.setNumThreads(1)
                m = mk.copy()
                t = time.perf_counter()
                cv2.watershed(img, m)
                line["parity"] = {"against": "cv2.watershed (OpenCV %s)" % cv2.__version__, "labels_differing": int((lab0 != m).sum()),
                                  "ok": bool((lab0 == m).all()), "cv2_ms_this_frame": (time.perf_counter() - t) * 1e3}
            except Exception as e:
                line["parity"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu:
            cpu_baseline_leg(args, line)
        args.out.emit(json.dumps(line))
    rig.finish()
    return 0


# ---- through the .ofx bundles (mini-host) -----------------------------------------------------------------------
def bench_plugin_boundary(pkg, synth, ctx, W, H):
    """kOfxImageEffectActionRender of the drop-in bundles, driven by the mini-host the way an OFX host drives them
    (/root/reference/openfx/HostSupport/src/ofxhImageEffect.cpp:1361-1436): host-memory clips in pageable memory."""
    import numpy as np
    mh = importlib.import_module("openfx-opencv_b200.minihost")
    out = {}
    base = synth.gray(synth.texture(H, W, seed=2000))
    lin = (np.arange(256, dtype=np.float32) / 255.0) ** 2.2      # any monotone byte -> linear float ramp will do here
    nfr = 8
    frames = {}
    for t in range(nfr):
        g = base if t == 0 else synth.shift_bilinear(base, 2.5 * t, -1.5 * t)
        f = np.empty((H, W, 4), np.float32)
        f[..., 0] = f[..., 1] = f[..., 2] = lin[g]
        f[..., 3] = 1.0
        frames[t] = f
    dst = np.zeros((H, W, 4), np.float32)

    def render_clip(labelled, cuda):
        mh.Plugin.provide_unique_identifiers(labelled)
        p = mh.Plugin("VectorGenerator")
        assert p.create_instance() == 0
        keep = []
        if cuda:
            dout = ctx.alloc(W * H * 16)
            for t, f in frames.items():
                d = ctx.to_device(f)
                keep.append(d)
                p.set_device_image("Source", t, d.ptr, W, H, 4, np.float32)
        else:
            for t, f in frames.items():
                p.set_image("Source", t, f)
        times = []
        for t in range(1, nfr - 1):
            p.clear_images("Output")
            if cuda:
                p.set_device_image("Output", t, dout.ptr, W, H, 4, np.float32)
            else:
                p.set_image("Output", t, dst)
            t0 = time.perf_counter()
            st = p.render(t, (0, 0, W, H), cuda_enabled=1 if cuda else -1)
            times.append((time.perf_counter() - t0) * 1e3)
            assert st == 0, st
        p.close()
        mh.Plugin.provide_unique_identifiers(True)
        steady = times[1:]
        return {"ms_per_render": statistics.median(steady), "first_render_ms": times[0], "renders_timed": len(steady),
                "pairs_per_s": 2e3 / statistics.median(steady)}

    def best_of(n, labelled, cuda):
        # the box is shared with other tenants' host load: a whole pass over the clip can come out 2x slower than the next
        # one with identical code (measured); every pass is listed, the fastest pass's median is the figure
        passes = [render_clip(labelled, cuda) for _ in range(n)]
        best = min(passes, key=lambda r: r["ms_per_render"])
        best["passes_ms_per_render"] = [round(r["ms_per_render"], 3) for r in passes]
        return best

    a = best_of(5, True, False)
    a.update({"h2d_bytes_per_render": W * H * 16, "d2h_bytes_per_render": W * H * 16,
              "what": "VectorGenerator.ofx, %dx%d float RGBA clips in pageable host memory, default parameters (forward AND backward flow = two "
                      "pairs per render), images labelled by the host (kOfxImagePropUniqueIdentifier): the staged gray frames of the "
                      "previous render are reused, one new frame is uploaded per render" % (W, H)})
    out["vectorgenerator_host_clips"] = a
    b = render_clip(False, False)
    b.update({"h2d_bytes_per_render": 3 * W * H * 16, "d2h_bytes_per_render": W * H * 16,
              "what": "the same with a host that does not label its images: all three frames are uploaded and converted every render"})
    out["vectorgenerator_host_clips_unlabelled"] = b
    c = best_of(3, True, True)
    c.update({"h2d_bytes_per_render": 0, "d2h_bytes_per_render": 0,
              "what": "the same with kOfxImageEffectPropCudaEnabled clips (device pointers): no staging copies"})
    out["vectorgenerator_cuda_clips"] = c
    rgb = synth.texture(H, W, seed=4)
    mask = synth.iid_mask(H, W, 1000, 0.10)
    rgba = np.dstack([np.maximum(rgb, 1), np.full((H, W), 255, np.uint8)]).astype(np.uint8)
    rgba[mask != 0, :3] = 0
    for name in ("inpaint", "segment"):
        q = mh.Plugin(name)
        assert q.create_instance() == 0
        o8 = np.zeros((H, W, 4), np.uint8)
        q.set_image("Source", 0, rgba)
        q.set_image("Output", 0, o8)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            st = q.render(0, (0, 0, W, H))
            ts.append((time.perf_counter() - t0) * 1e3)
            assert st == 0, st
        q.close()
        out[name + "_host_clips"] = {"ms_per_render": statistics.median(ts[1:]), "first_render_ms": ts[0], "frames_per_s": 1e3 / statistics.median(ts[1:]),
                                     "h2d_bytes_per_render": W * H * 4, "d2h_bytes_per_render": W * H * 4,
                                     "what": "%s.ofx, %dx%d RGBA8 clips in pageable host memory, default parameters" % (name, W, H)}
    return out


def bench_plugins(pkg, synth, ctx, W, H, with_cpu):
    """The other plugin bodies at the bench resolution (frames resident in HBM, wall-clock around a synchronise: these
    bodies are many launches each); CPU = cv2 on one host core (they are single-threaded algorithms).  Their own
    bench lines: --workload inpaint_ns_4k / inpaint_telea_vga / watershed_4k."""
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
    # config 1: Telea at 640x480 with a 5 % mask, one frame at a time
    i1, m1 = inpaint_inputs(synth, 640, 480, 0.05, 1)
    a1, b1, o1 = ctx.to_device(i1), ctx.to_device(m1[0]), ctx.alloc(640 * 480 * 3)
    dt = gpu_time(lambda: ctx.inpaint_dev(a1.ptr, 3, b1.ptr, o1.ptr, 640, 480, 3.0, pkg.INPAINT_TELEA), 10)
    ent = {"value": 1 / dt, "unit": "frames/s", "ms_per_frame": dt * 1e3, "workload": "BASELINE config 1: 640x480 RGB8, 5% iid mask, radius 3, Telea"}
    if cv2 is not None and with_cpu:
        t = time.perf_counter()
        ref = cv2.inpaint(i1, m1[0], 3.0, cv2.INPAINT_TELEA)
        ent["cpu_frames_per_s"] = 1 / (time.perf_counter() - t)
        ent["bytes_differing_from_cv2"] = int((o1.download((480, 640, 3), np.uint8) != ref).sum())
    out["inpaint_telea_vga"] = ent
    # Dual TV-L1, the VectorGenerator plugin's second method (default parameters; parity of the method is unpinned)
    base = synth.gray(synth.texture(H, W, seed=2000))
    nxt = synth.shift_bilinear(base, 2.5, -1.5)
    d_a, d_b, d_f = ctx.to_device(base), ctx.to_device(nxt), ctx.alloc(W * H * 8)
    L = pkg.lib()
    dt = gpu_time(lambda: ctx.tvl1_dev(d_a.ptr, d_b.ptr, W, H, d_f.ptr), 3)
    ent = {"value": 1 / dt, "unit": "pairs/s", "ms_per_pair": dt * 1e3, "inner_iterations_run": int(L.ofxcv_tvl1_iterations_run(ctx.h)),
           "workload": "%dx%d gray8 pair, plugin defaults (5 scales, 5 warps, 10 x 15 iterations, epsilon 0.01)" % (W, H)}
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
    img_w, mk = watershed_inputs(synth, W, H, 256)
    for nf in (1, 512):
        d_rgbs, d_mks = ctx.alloc(W * H * 3 * nf), ctx.alloc(W * H * 4 * nf)
        L.ofxcv_upload(ctx.h, None, d_rgbs.ptr, img_w.ctypes.data, W * H * 3)
        L.ofxcv_upload(ctx.h, None, d_mks.ptr, mk.ctypes.data, W * H * 4)
        for f in range(1, nf):   # the same frame nf times: replicated on the device
            L.ofxcv_device_copy(ctx.h, None, d_rgbs.ptr + f * W * H * 3, d_rgbs.ptr, W * H * 3)
            L.ofxcv_device_copy(ctx.h, None, d_mks.ptr + f * W * H * 4, d_mks.ptr, W * H * 4)
        ctx.synchronize()
        if nf == 1:   # warm the workspaces of the parallel flood (a first call allocates them)
            ctx.watershed_dev(d_rgbs.ptr, d_mks.ptr, W, H, nf)
            L.ofxcv_upload(ctx.h, None, d_mks.ptr, mk.ctypes.data, W * H * 4)
            ctx.synchronize()
        t = time.perf_counter()
        ctx.watershed_dev(d_rgbs.ptr, d_mks.ptr, W, H, nf)
        ctx.synchronize()
        dt = time.perf_counter() - t
        ent = {"value": nf / dt, "unit": "frames/s", "frames_in_flight": nf, "ms_total": dt * 1e3, "workload": "%dx%d RGB8, 256 seeds" % (W, H),
               "algorithmic_gbs": 11.0 * W * H * nf / dt / 1e9,
               "path": "exact intra-frame parallel flood (watershed_par.cu)" if nf < 16 else "one thread per frame, all frames in flight",
               "bound": "dependency chains of the ordered flood, not HBM"}
        if nf == 1 and cv2 is not None and with_cpu:
            m = mk.copy()
            t = time.perf_counter()
            cv2.watershed(img_w, m)
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
    ap.add_argument("--workload", default="farneback_4k", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="outputs (flow pairs / frames) per GPU per step; 0 = the workload's block")
    ap.add_argument("--chunk", type=int, default=0, help="flow pairs per clip call; 0 = the workload's default")
    ap.add_argument("--clip-frames", type=int, default=0, help="fixed sequence of this many frames split over the GPUs (strong scaling)")
    ap.add_argument("--ref-workers", type=int, default=64, help="cap on CPU worker processes of the reference arm")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-plugins", action="store_true", help="skip the plugin-boundary and other-plugin entries")
    ap.add_argument("--no-parity", action="store_true", help="skip the comparison of output 0 with the reference's OpenCV call")
    ap.add_argument("--plugin-leg", action="store_true", help=argparse.SUPPRESS)   # internal: the e2e_plugin entry, run as a subprocess
    args = ap.parse_args()
    args.out = OneLineStdout()
    if args.plugin_leg:
        pkg = importlib.import_module("openfx-opencv_b200")
        synth = importlib.import_module("openfx-opencv_b200.synth")
        ctx = pkg.Context(int(os.environ.get("LOCAL_RANK", "0")))
        args.out.emit(json.dumps(bench_plugin_boundary(pkg, synth, ctx, 3840, 2160)))
        return 0
    if args.impl == "reference":
        return run_reference(args)
    wl = WORKLOADS[args.workload]
    return {"flow": run_flow, "inpaint": run_inpaint, "watershed": run_watershed}[wl["kind"]](args, wl)


if __name__ == "__main__":
    sys.exit(main())
