"""Time-sharding of one video across the GPUs of a box, and the one exchange step the path has.

Frames are classified independently, so rank r takes the contiguous range ``shard_range(N, r, R)`` and runs the whole
per-frame pipeline on it with no communication.  Only the per-shard RUN TABLES are exchanged:

    cutdet_shard_pack      this rank's table -> one packed buffer {n_runs, n_frames | 40-byte rows}      (1 launch)
    all_gather_into_tensor the packed buffers of all ranks, rank order = time order                       (NCCL over NVLink)
    cutdet_stitch_packed   reads the gathered buffer directly: counts and frame offsets from the headers, (1 launch)
                           runs that meet at a shard edge with the same type are joined (sums and lengths add)

with no host synchronisation anywhere between ``FramePipeline.finish()`` and the final ``to_te()``.  Every rank then runs
the global, order-dependent smoothing pass redundantly (it cannot be sharded: the least-confident-first order of
glue_orphans is global, reference frameID/segmentation.py:103-107).  The reference itself has no multi-GPU path
(SURVEY.md section 2.1); the contract is "the same table as one process over the concatenated scores"
(segmentation.py:35-60).

The gather capacity is a fixed number of rows per shard (default 4,096 = 160 KB; a full game has a few hundred runs).  A
shard with more runs is reported by the stitch kernel through a negative run count, which ``DeviceRunTable.to_te()`` turns
into ``ShardOverflow``; ``stitch_all_checked`` catches it and repeats the exchange with a capacity that fits.

The same packed format joins the tables of the decode workers' time ranges on ONE GPU (cutdet.decode): ``stitch_local``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi

# packed row: end, start, length (int64), sum (float64), type (int32), mean (float32) = 40 bytes
ROW_BYTES = 40
HEADER_BYTES = 16           # n_runs (int64), n_frames (int64)
DEFAULT_CAPACITY = 4096
ROW_DTYPE = np.dtype([("end", "<i8"), ("start", "<i8"), ("length", "<i8"), ("sum", "<f8"), ("type", "<i4"), ("mean", "<f4")])
assert ROW_DTYPE.itemsize == ROW_BYTES


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous time range: ceil(N / R) frames per rank, the last ranks may be short or empty."""
    per = -(-n_frames // world)
    lo = min(n_frames, rank * per)
    return lo, min(n_frames, lo + per)


def packed_bytes(capacity: int) -> int:
    return HEADER_BYTES + int(capacity) * ROW_BYTES


# ------------------------------------------------------------------------------------------------- host-side packer
# The byte layout of cutdet_shard_pack restated with numpy: used by the CPU (gloo) tests of the exchange and as the checker of
# the kernel in the GPU tests.
def pack_columns(columns: dict, n_runs: int, n_frames: int, capacity: int) -> torch.Tensor:
    """Run-table columns (CPU tensors or arrays of >= n_runs rows) -> one uint8 CPU tensor in the packed layout."""
    buf = np.zeros(packed_bytes(capacity), dtype=np.uint8)
    buf[:HEADER_BYTES].view("<i8")[:] = (n_runs, n_frames)
    rows = buf[HEADER_BYTES:].view(ROW_DTYPE)
    m = min(n_runs, capacity)
    for field, name in (("end", "end_frames"), ("start", "start_frames"), ("length", "run_lengths"), ("sum", "score_sums"),
                        ("type", "frame_types"), ("mean", "score_means")):
        rows[field][:m] = np.asarray(columns[name][:m])
    return torch.from_numpy(buf)


def unpack_columns(buf: torch.Tensor, capacity: int):
    """Inverse of pack_columns on a CPU buffer: (columns dict of numpy arrays with n_runs rows, n_runs, n_frames)."""
    raw = buf.cpu().numpy()
    n_runs, n_frames = (int(v) for v in raw[:HEADER_BYTES].view("<i8"))
    rows = raw[HEADER_BYTES:HEADER_BYTES + capacity * ROW_BYTES].view(ROW_DTYPE)[:min(n_runs, capacity)]
    cols = {"end_frames": rows["end"].copy(), "start_frames": rows["start"].copy(), "run_lengths": rows["length"].copy(),
            "score_sums": rows["sum"].copy(), "frame_types": rows["type"].astype(np.int64), "score_means": rows["mean"].copy()}
    return cols, n_runs, n_frames


# ------------------------------------------------------------------------------------------------- device path
def pack_table(table, n_frames_local: int, capacity: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """K: cutdet_shard_pack.  The table's rows -> a packed uint8 buffer on the table's device (no host sync)."""
    from . import engine

    if out is None:
        out = torch.empty(packed_bytes(capacity), dtype=torch.uint8, device=table.device)
    ts = table.struct()
    _cabi.check(_cabi.lib().cutdet_shard_pack(C.byref(ts), table.n_runs.data_ptr(), int(n_frames_local), int(capacity),
                                              out.data_ptr(), engine._stream()))
    return out


def all_gather_packed(packed: torch.Tensor, group=None) -> torch.Tensor:
    """One all-gather of every rank's packed table (NCCL over NVLink for CUDA tensors, gloo for CPU tensors):
    [world * packed_bytes] in rank order."""
    world = dist.get_world_size(group)
    out = torch.empty(world * packed.numel(), dtype=torch.uint8, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return out


def stitch_packed(gathered: torch.Tensor, n_shards: int, capacity: int, dst_capacity: int | None = None):
    """K: cutdet_stitch_packed.  Returns (global DeviceRunTable, total_frames device tensor); no host sync."""
    from . import engine

    dst = engine.DeviceRunTable(dst_capacity or n_shards * capacity, gathered.device)
    dst.shard_capacity = capacity
    total = torch.zeros(1, dtype=torch.int64, device=gathered.device)
    d = dst.struct()
    _cabi.check(_cabi.lib().cutdet_stitch_packed(gathered.data_ptr(), int(n_shards), int(capacity), C.byref(d),
                                                 dst.n_runs.data_ptr(), total.data_ptr(), engine._stream()))
    return dst, total


def stitch_all(table, n_frames_local: int, capacity: int = DEFAULT_CAPACITY, group=None):
    """The exchange step: pack, all-gather, stitch.  Returns (global table, total_frames device tensor).  A shard that did not
    fit ``capacity`` surfaces as ShardOverflow from the table's ``to_te()`` / ``count()``."""
    gathered = all_gather_packed(pack_table(table, n_frames_local, capacity), group)
    return stitch_packed(gathered, dist.get_world_size(group), capacity)


def stitch_local(tables: list, frame_counts: list, capacity: int = DEFAULT_CAPACITY):
    """Tables of consecutive time ranges that live on ONE device (the decode workers' ranges) -> the table of the whole
    sequence: the same pack + stitch kernels, the all-gather replaced by packing into slices of one buffer."""
    dev = tables[0].device
    buf = torch.empty(len(tables) * packed_bytes(capacity), dtype=torch.uint8, device=dev)
    for i, (t, n) in enumerate(zip(tables, frame_counts)):
        pack_table(t, n, capacity, out=buf[i * packed_bytes(capacity):(i + 1) * packed_bytes(capacity)])
    return stitch_packed(buf, len(tables), capacity)


def finish_checked(exchange, after=None, capacity: int = DEFAULT_CAPACITY):
    """Run ``exchange(capacity) -> (table, total)`` then ``after(table)`` (the smoothing pass) and copy the result to the host;
    if a shard had more runs than the gather capacity, repeat once with a capacity that fits.  Returns (te, total_frames)."""
    from . import engine

    for attempt in range(2):
        table, total = exchange(capacity)
        if after is not None:
            after(table)
        try:
            te = table.to_te()
        except engine.ShardOverflow as e:
            if attempt:
                raise
            capacity = 1 << int(e.needed - 1).bit_length()
            continue
        return te, int(total.item())
    raise AssertionError("unreachable")
