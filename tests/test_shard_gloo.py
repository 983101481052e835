"""The N > 1 host path on CPU: two ranks over gloo, each encodes its contiguous time shard, the packed run tables
are exchanged with ONE all-gather (cutdet.shard), and the gathered shards -- joined by the oracle's stitch, the
checker of the CUDA stitch kernel -- must equal the run table of the whole sequence.  No CUDA, no compute calls
into libcutdet_b200.so: the per-shard tables come from the oracle; what is under test is shard_range, the packed byte
layout (pack_columns / unpack_columns, the numpy restatement of cutdet_shard_pack that the GPU suite checks the kernel
against) and all_gather_packed (reference: none -- the reference is single-process)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import segmentation as oseg


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _sequence(n: int, seed: int):
    rng = np.random.default_rng(seed)
    lab = np.repeat(rng.integers(0, 3, n // 37 + 2), 37)[:n].astype(np.uint8)
    top = rng.uniform(1, 9, n).astype(np.float32)
    return lab, top


def _local_table(lab, top):
    te = oseg.run_table_from_labels(lab, top)
    starts, ends = te["start_frames"], te["end_frames"]
    te["score_sums"] = np.array([np.sum(top[s:e + 1], dtype=np.float64) for s, e in zip(starts, ends)], dtype=np.float64)
    return te


def _worker(rank: int, world: int, port: int, n: int, seed: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cutdet import shard

        lab, top = _sequence(n, seed)
        lo, hi = shard.shard_range(n, rank, world)
        te = _local_table(lab[lo:hi], top[lo:hi]) if hi > lo else None
        n_runs = 0 if te is None else len(te["end_frames"])
        cap = 4096
        cols = te if te is not None else {k: np.zeros(0) for k in ("end_frames", "start_frames", "run_lengths", "score_sums",
                                                                   "frame_types", "score_means")}
        packed = shard.pack_columns(cols, n_runs, hi - lo, cap)
        assert packed.numel() == shard.packed_bytes(cap)
        gathered = shard.all_gather_packed(packed).view(world, -1)
        shards, offsets, total = [], [], 0
        for r in range(world):
            c, n_r, frames_r = shard.unpack_columns(gathered[r], cap)
            assert len(c["end_frames"]) == n_r
            shards.append(c)
            offsets.append(total)
            total += frames_r
        assert total == n
        joined = oseg.stitch_tables(shards, [int(o) for o in offsets])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **joined)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 20_011), (2, 37), (3, 5_000)])
def test_two_rank_gather_and_stitch_equals_whole(tmp_path, world, n):
    seed = 11
    mp.spawn(_worker, args=(world, _free_port(), n, seed, str(tmp_path)), nprocs=world, join=True)
    lab, top = _sequence(n, seed)
    want = oseg.run_table_from_labels(lab, top)
    for rank in range(world):
        got = np.load(os.path.join(tmp_path, f"rank{rank}.npz"))
        for k in ("end_frames", "start_frames", "run_lengths", "frame_types"):
            assert np.array_equal(got[k], want[k]), (rank, k)
        np.testing.assert_allclose(got["score_means"], want["score_means"], rtol=2e-6)


def test_shard_range_partitions_every_frame():
    from cutdet import shard
    for n in (0, 1, 7, 1800, 324_000):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
