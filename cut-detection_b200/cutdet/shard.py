"""Time-sharding of one video across the GPUs of a box, and the one exchange step the path has.

Frames are classified independently, so rank r takes the contiguous range ``shard_range(N, r, R)`` and runs the whole
per-frame pipeline on it with no communication.  Only the per-shard RUN TABLES are exchanged -- one NCCL all-gather
of a fixed-capacity packed table (40 bytes per run; carrying float64 sums and lengths, not means, so a run cut by a
shard edge is re-joined exactly) -- after which every rank stitches the shards and runs the global, order-dependent
smoothing pass redundantly (it cannot be sharded: the least-confident-first order of glue_orphans is global,
reference frameID/segmentation.py:103-107).  The reference itself has no multi-GPU path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

# packed row: end, start, length (int64), sum (float64), type (int32), mean (float32) = 40 bytes
ROW_BYTES = 40
HEADER_BYTES = 16           # n_runs (int64), n_frames (int64)


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous time range: ceil(N / R) frames per rank, the last ranks may be short or empty."""
    per = -(-n_frames // world)
    lo = min(n_frames, rank * per)
    return lo, min(n_frames, lo + per)


def pack_columns(columns: dict, n_runs: int, n_frames: int, capacity: int) -> torch.Tensor:
    """Run-table columns (tensors of >= n_runs rows, any device) -> one uint8 buffer of fixed size."""
    dev = columns["end_frames"].device
    buf = torch.zeros(HEADER_BYTES + capacity * ROW_BYTES, dtype=torch.uint8, device=dev)
    buf[:HEADER_BYTES].view(torch.int64).copy_(torch.tensor([n_runs, n_frames], dtype=torch.int64))
    body = buf[HEADER_BYTES:]
    off = 0
    for name, dtype, width in (("end_frames", torch.int64, 8), ("start_frames", torch.int64, 8),
                               ("run_lengths", torch.int64, 8), ("score_sums", torch.float64, 8),
                               ("frame_types", torch.int32, 4), ("score_means", torch.float32, 4)):
        dst = body[off:off + capacity * width].view(dtype)
        dst[:n_runs] = columns[name][:n_runs].to(dtype)
        off += capacity * width
    return buf


def unpack_columns(buf: torch.Tensor, capacity: int):
    """Inverse of pack_columns: (columns dict with `capacity` rows each, n_runs, n_frames)."""
    head = buf[:HEADER_BYTES].view(torch.int64)
    body = buf[HEADER_BYTES:]
    cols, off = {}, 0
    for name, dtype, width in (("end_frames", torch.int64, 8), ("start_frames", torch.int64, 8),
                               ("run_lengths", torch.int64, 8), ("score_sums", torch.float64, 8),
                               ("frame_types", torch.int32, 4), ("score_means", torch.float32, 4)):
        cols[name] = body[off:off + capacity * width].view(dtype)
        off += capacity * width
    return cols, int(head[0].item()), int(head[1].item())


def all_gather_packed(packed: torch.Tensor, group=None) -> list[torch.Tensor]:
    """One all-gather of every rank's packed table (NCCL over NVLink for CUDA tensors, gloo for CPU tensors)."""
    world = dist.get_world_size(group)
    out = torch.empty(world * packed.numel(), dtype=torch.uint8, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return list(out.view(world, -1).unbind(0))


def gather_tables(table, n_frames_local: int, capacity: int, group=None):
    """All ranks' DeviceRunTables -> one DeviceRunTable holding shard i in rows [i*capacity, ...), plus the per-shard
    run counts and global frame offsets as device tensors (the inputs of engine.stitch_shards)."""
    from . import engine

    n = table.count()
    if n > capacity:
        raise RuntimeError(f"shard run table has {n} runs, gather capacity is {capacity}")
    cols = {k: getattr(table, k) for k in ("end_frames", "start_frames", "run_lengths", "score_sums", "frame_types",
                                           "score_means")}
    shards = all_gather_packed(pack_columns(cols, n, n_frames_local, capacity), group)
    world = len(shards)
    big = engine.DeviceRunTable(capacity * world, table.device)
    counts, offsets, total = [], [], 0
    for r, buf in enumerate(shards):
        c, n_runs, n_frames = unpack_columns(buf, capacity)
        for name in cols:
            getattr(big, name)[r * capacity:(r + 1) * capacity] = c[name]
        counts.append(n_runs)
        offsets.append(total)
        total += n_frames
    dev = table.device
    return (big, torch.tensor(counts, dtype=torch.int64, device=dev), torch.tensor(offsets, dtype=torch.int64, device=dev),
            total)


def stitch_all(table, n_frames_local: int, capacity: int, group=None):
    """The exchange step: all-gather the shard tables, join them (K: stitch_shards).  Returns (global table, N)."""
    from . import engine

    big, counts, offsets, total = gather_tables(table, n_frames_local, capacity, group)
    return engine.stitch_shards(big, counts, offsets, capacity), total
