"""Frame-sequence sharding for the multi-GPU driver.

Frames of a sequence are independent units (inpaint, watershed: one frame; flow: one PAIR), so the sequence is cut
into contiguous blocks, one per rank, with no data-path collective.  Flow needs frame t+1 for output t, so a rank
that owns outputs [first, first+count) stages frames [first, first+count] — a one-frame halo
(VectorGeneratorPlugin::getFramesNeeded, /root/reference/VectorGenerator/VectorGenerator.cpp:675-695).
torch.distributed is used only for the start/stop barrier and the gather of per-rank results.
"""
import numpy as np


def shard_range(n_units, world, rank):
    """(first, count) of the contiguous block of `n_units` outputs owned by `rank` (blocks differ by at most 1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n_units, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def frames_needed(first, count, n_frames, forward=True, backward=False):
    """Indices of the input frames a rank must stage for outputs [first, first+count): its block plus the halo."""
    lo = first - (1 if backward else 0)
    hi = first + count - 1 + (1 if forward else 0)
    return list(range(max(lo, 0), min(hi, n_frames - 1) + 1))


def checksum64(arr):
    """Order-sensitive 64-bit checksum of an output frame (gathered instead of the frames themselves)."""
    a = np.ascontiguousarray(arr).view(np.uint8).ravel().astype(np.uint64)
    idx = np.arange(1, a.size + 1, dtype=np.uint64)
    return int((a * (idx * np.uint64(0x9E3779B97F4A7C15) | np.uint64(1))).sum(dtype=np.uint64))


def gather_results(local, group=None):
    """all_gather of small per-rank python results (frame ranges, checksums, timings) to every rank."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return [local]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    return out


def run_sharded(n_units, process_unit, world=None, rank=None):
    """Run process_unit(i) -> small result for every unit of this rank's block, barrier on both sides, and return
    the results of ALL ranks in unit order (on every rank)."""
    import torch.distributed as dist
    inited = dist.is_available() and dist.is_initialized()
    world = world if world is not None else (dist.get_world_size() if inited else 1)
    rank = rank if rank is not None else (dist.get_rank() if inited else 0)
    first, count = shard_range(n_units, world, rank)
    if inited:
        dist.barrier()
    mine = [(i, process_unit(i)) for i in range(first, first + count)]
    if inited:
        dist.barrier()
    merged = {}
    for part in gather_results(mine):
        merged.update(dict(part))
    return [merged[i] for i in range(n_units)]


def flow_clip(ctx, load_frame, n_frames, params=None, world=None, rank=None, keep=True, method="farneback"):
    """Forward flow t -> t+1 of a whole clip, frame-sharded over the ranks of the default process group.
    method = "farneback" (one clip call, pyramids shared between pairs) or "tvl1" (the plugin's second method, pair by
    pair: it has no per-frame state to share).

    load_frame(t) -> HxW uint8 (host) is called only for the frames this rank needs (its contiguous block of outputs
    plus the one-frame halo).  The rank's block goes through ONE clip call of the C ABI (ofxcv_farneback_sequence_u8_host:
    one pyramid per frame, two pairs in flight, copies overlapped).  Returns (first, flows or None, checksums) where
    `checksums` are the 64-bit checksums of ALL n_frames-1 flow fields in clip order, gathered from every rank (the
    only collective on this path, besides the barriers)."""
    import torch.distributed as dist
    inited = dist.is_available() and dist.is_initialized()
    world = world if world is not None else (dist.get_world_size() if inited else 1)
    rank = rank if rank is not None else (dist.get_rank() if inited else 0)
    first, count = shard_range(n_frames - 1, world, rank)
    flows = []
    if count > 0:
        frames = [np.ascontiguousarray(load_frame(t), np.uint8) for t in frames_needed(first, count, n_frames)]
        if inited:
            dist.barrier()
        if method == "farneback":
            flows = ctx.farneback_sequence(frames, params)
        elif method == "tvl1":
            flows = [ctx.tvl1(a, b, params)[0] for a, b in zip(frames[:-1], frames[1:])]
        else:
            raise ValueError("method must be 'farneback' or 'tvl1'")
    elif inited:
        dist.barrier()
    mine = [(first + i, checksum64(f)) for i, f in enumerate(flows)]
    merged = {}
    for part in gather_results(mine):
        merged.update(dict(part))
    return first, (flows if keep else None), [merged[i] for i in range(n_frames - 1)]


def inpaint_clip(ctx, load_frame, n_frames, radius, method, world=None, rank=None, keep=True, frames_in_flight=0):
    """Inpaint a whole clip, frame-sharded over the ranks of the default process group (BASELINE.json config 4).

    load_frame(t) -> (HxWx3 or HxW uint8 image, HxW uint8 mask) is called only for this rank's contiguous block; the
    block goes through ONE clip call of the C ABI (ofxcv_inpaint_sequence_u8_host: several frames in flight).  Returns
    (first, frames or None, checksums of ALL n_frames outputs in clip order, gathered from every rank)."""
    import torch.distributed as dist
    inited = dist.is_available() and dist.is_initialized()
    world = world if world is not None else (dist.get_world_size() if inited else 1)
    rank = rank if rank is not None else (dist.get_rank() if inited else 0)
    first, count = shard_range(n_frames, world, rank)
    pairs = [load_frame(t) for t in range(first, first + count)]
    if inited:
        dist.barrier()
    outs = ctx.inpaint_sequence([p[0] for p in pairs], [p[1] for p in pairs], radius, method, frames_in_flight) if count else []
    mine = [(first + i, checksum64(o)) for i, o in enumerate(outs)]
    merged = {}
    for part in gather_results(mine):
        merged.update(dict(part))
    return first, (outs if keep else None), [merged[i] for i in range(n_frames)]


def watershed_clip(ctx, load_frame, n_frames, world=None, rank=None, keep=True, frames_in_flight=128):
    """Watershed of a whole clip, frame-sharded over the ranks of the default process group.

    load_frame(t) -> (HxWx3 uint8 image, HxW int32 markers), called only for this rank's contiguous block; the block is
    flooded `frames_in_flight` frames at a time (ofxcv_watershed_u8c3_batch — the exact flood is sequential per frame, a
    clip's throughput comes from frames in flight).  Returns (first, label maps or None, checksums of ALL frames)."""
    import torch.distributed as dist
    inited = dist.is_available() and dist.is_initialized()
    world = world if world is not None else (dist.get_world_size() if inited else 1)
    rank = rank if rank is not None else (dist.get_rank() if inited else 0)
    first, count = shard_range(n_frames, world, rank)
    if inited:
        dist.barrier()
    outs, sums = [], []
    for lo in range(first, first + count, max(1, frames_in_flight)):
        pairs = [load_frame(t) for t in range(lo, min(lo + max(1, frames_in_flight), first + count))]
        labs = ctx.watershed_sequence([p[0] for p in pairs], [p[1] for p in pairs])
        sums += [(lo + i, checksum64(l)) for i, l in enumerate(labs)]
        if keep:
            outs += labs
    merged = {}
    for part in gather_results(sums):
        merged.update(dict(part))
    return first, (outs if keep else None), [merged.get(i) for i in range(n_frames)]   # None: owned by a rank outside the group
