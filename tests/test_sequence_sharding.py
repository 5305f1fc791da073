"""Host-side logic of the multi-GPU sequence driver, exercised with world_size-2 gloo on CPU."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_everything():
    seq = importlib.import_module("openfx-opencv_b200.sequence")
    for n in (0, 1, 7, 8, 300, 1000):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                first, count = seq.shard_range(n, world, r)
                got += list(range(first, first + count))
                assert count in (n // world, n // world + 1)
            assert got == list(range(n))
    assert seq.frames_needed(4, 4, 20) == [4, 5, 6, 7, 8]            # forward flow: +1 halo frame
    assert seq.frames_needed(4, 4, 20, backward=True) == [3, 4, 5, 6, 7, 8]
    assert seq.frames_needed(16, 4, 20) == [16, 17, 18, 19]          # clamped at the sequence end
    with pytest.raises(ValueError):
        seq.shard_range(4, 2, 2)


def _worker(rank, world, port, n_units, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq = importlib.import_module("openfx-opencv_b200.sequence")

    def unit(i):  # stands for "render output frame i and checksum it"
        rng = np.random.default_rng(i)
        return (rank, seq.checksum64(rng.integers(0, 255, (16, 16), dtype=np.uint8)))

    res = seq.run_sharded(n_units, unit)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)   # the bench's max-over-ranks timing reduction
    q.put((rank, res, float(t.item())))
    dist.destroy_process_group()


def test_run_sharded_gloo_world2():
    seq = importlib.import_module("openfx-opencv_b200.sequence")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_units, world, port = 7, 2, 29517 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [seq.checksum64(np.random.default_rng(i).integers(0, 255, (16, 16), dtype=np.uint8)) for i in range(n_units)]
    for rank, res, tmax in outs:
        assert [c for _, c in res] == expect          # every rank sees every unit's result, in order
        owners = [o for o, _ in res]
        assert owners == [0, 0, 0, 0, 1, 1, 1]        # contiguous blocks
        assert tmax == 2.0


class _FakeCtx:
    """Stands in for the CUDA context: 'flow' of a pair = difference of the two frames (host logic under test is the
    sharding, the halo and the gather)."""
    def farneback_sequence(self, frames, params=None):
        return [np.stack([b.astype(np.float32) - a, a.astype(np.float32)], axis=-1) for a, b in zip(frames[:-1], frames[1:])]


class _FakeWatershedCtx:
    def __init__(self):
        self.batches = []

    def watershed_sequence(self, rgbs, markers):
        self.batches.append(len(rgbs))
        return [np.where(m > 0, m, -1).astype(np.int32) for m in markers]


def test_watershed_clip_batches_and_order():
    """Host logic of the watershed clip driver on one rank: contiguous block, frames_in_flight-sized batches."""
    seq = importlib.import_module("openfx-opencv_b200.sequence")
    fake = _FakeWatershedCtx()
    frames = [(np.zeros((3, 4, 3), np.uint8), np.full((3, 4), t % 3, np.int32)) for t in range(10)]
    first, labs, sums = seq.watershed_clip(fake, lambda t: frames[t], 10, world=2, rank=1, frames_in_flight=2)
    assert first == 5 and fake.batches == [2, 2, 1] and len(labs) == 5
    assert sums[:5] == [None] * 5   # no process group here: the other rank's checksums are not gathered
    assert sums[5:] == [seq.checksum64(np.where(frames[t][1] > 0, frames[t][1], -1).astype(np.int32)) for t in range(5, 10)]


class _FakeInpaintCtx:
    def inpaint_sequence(self, imgs, masks, radius, method, frames_in_flight=0):
        return [np.where(m[..., None] != 0, 255 - a, a).astype(np.uint8) for a, m in zip(imgs, masks)]


def _inpaint_clip_worker(rank, world, port, n_frames, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq = importlib.import_module("openfx-opencv_b200.sequence")
    loaded = []

    def load(t):
        loaded.append(t)
        return np.full((4, 6, 3), 10 * t, np.uint8), (np.arange(24).reshape(4, 6) % (t + 2) == 0).astype(np.uint8)

    first, outs, sums = seq.inpaint_clip(_FakeInpaintCtx(), load, n_frames, 3.0, 0)
    q.put((rank, first, loaded, len(outs), sums))
    dist.destroy_process_group()


def test_inpaint_clip_gloo_world2():
    seq = importlib.import_module("openfx-opencv_b200.sequence")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames, world, port = 5, 2, 29717 + os.getpid() % 1000
    procs = [ctx.Process(target=_inpaint_clip_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fake = _FakeInpaintCtx()
    expect = []
    for t in range(n_frames):
        a, m = np.full((4, 6, 3), 10 * t, np.uint8), (np.arange(24).reshape(4, 6) % (t + 2) == 0).astype(np.uint8)
        expect.append(seq.checksum64(fake.inpaint_sequence([a], [m], 3.0, 0)[0]))
    (r0, f0, l0, n0, s0), (r1, f1, l1, n1, s1) = outs
    assert (f0, n0, l0) == (0, 3, [0, 1, 2]) and (f1, n1, l1) == (3, 2, [3, 4])   # contiguous blocks, no halo
    assert s0 == expect and s1 == expect


def _clip_worker(rank, world, port, n_frames, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq = importlib.import_module("openfx-opencv_b200.sequence")
    loaded = []

    def load(t):
        loaded.append(t)
        return np.full((4, 6), 3 * t, np.uint8)

    first, flows, sums = seq.flow_clip(_FakeCtx(), load, n_frames)
    q.put((rank, first, loaded, len(flows), sums))
    dist.destroy_process_group()


def test_flow_clip_gloo_world2():
    seq = importlib.import_module("openfx-opencv_b200.sequence")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames, world, port = 8, 2, 29617 + os.getpid() % 1000
    procs = [ctx.Process(target=_clip_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [seq.checksum64(np.stack([np.full((4, 6), 3.0, np.float32), np.full((4, 6), 3.0 * t, np.float32)], axis=-1)) for t in range(n_frames - 1)]
    (r0, f0, l0, n0, s0), (r1, f1, l1, n1, s1) = outs
    assert (f0, n0, l0) == (0, 4, [0, 1, 2, 3, 4])          # 7 pairs: rank 0 owns 4 of them + the halo frame
    assert (f1, n1, l1) == (4, 3, [4, 5, 6, 7])
    assert s0 == expect and s1 == expect                    # every rank sees every checksum, in clip order
