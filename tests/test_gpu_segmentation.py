"""K4/K5/K6 on the GPU vs the oracle and the tables recorded from the reference's Segmentation class.
Integer columns and CSV bytes must match exactly; run means to float32 rounding (the GPU sums in float64)."""
import os

import numpy as np
import pytest
import torch

import kat_inputs
from oracle import segmentation as oseg

pytestmark = pytest.mark.gpu

INT_COLS = ("end_frames", "frame_types", "run_lengths", "start_frames")


def _np(te):
    return {k: v.numpy() for k, v in te.items()}


def _same(te, want, means_rtol=2e-6):
    for k in INT_COLS:
        assert np.array_equal(te[k], want[k]), k
        assert te[k].dtype == np.int64
    assert te["score_means"].dtype == np.float32
    assert np.allclose(te["score_means"], want["score_means"], rtol=means_rtol, atol=1e-6)


def _kat(golden_dir, name, stage):
    z = np.load(os.path.join(golden_dir, "segmentation_kat.npz"))
    return {k: z[f"{name}/{stage}/{k}"] for k in INT_COLS + ("score_means",)}


def test_argmax_first_index_wins():
    from cutdet import engine
    s = torch.tensor([[1.0, 1.0, 0.0], [0.0, 2.0, 2.0], [3.0, 3.0, 3.0], [-1.0, -2.0, -0.5]], device="cuda")
    lab, top = engine.argmax(s)
    assert lab.cpu().tolist() == [0, 1, 0, 2] and top.cpu().tolist() == [1.0, 2.0, 3.0, -0.5]
    rng = np.random.default_rng(0)
    big = rng.normal(0, 1, (100003, 3)).astype(np.float32)
    big[::7, 1] = big[::7, 0]                       # exact ties
    lab, top = engine.argmax(torch.from_numpy(big).cuda())
    wt, wl = oseg.max_and_argmax(big)
    assert np.array_equal(lab.cpu().numpy(), wl) and np.array_equal(top.cpu().numpy(), wt)
    lab, top = engine.argmax(torch.from_numpy(rng.normal(0, 1, (1000, 8)).astype(np.float32)).cuda())  # generic C
    assert lab.shape[0] == 1000


@pytest.mark.parametrize("name", sorted(kat_inputs.segmentation_cases().keys()))
def test_reference_cases(name, golden_dir):
    """Segmentation(scores) -> glue_orphans -> combine_adjacent_segments -> CSV, stage by stage."""
    from frameID.segmentation import Segmentation
    scores, k1, kb = kat_inputs.segmentation_cases()[name]
    seg = Segmentation(torch.from_numpy(scores))          # CPU scores, as the reference CLI passes them
    _same(_np(seg.te), _kat(golden_dir, name, "init"))
    n0 = len(seg)
    seg.glue_orphans(k1, kb)
    # the merged means follow the reference's (quirky) update; they start from means equal to rounding
    _same(_np(seg.te), _kat(golden_dir, name, "glued"), means_rtol=1e-5)
    seg.combine_adjacent_segments()
    _same(_np(seg.te), _kat(golden_dir, name, "combined"), means_rtol=1e-5)
    assert len(seg) <= n0
    z = np.load(os.path.join(golden_dir, "segmentation_kat.npz"))
    out = os.path.join("/tmp", f"cutdet_test_{name}.csv")
    seg.write_csv(out)
    assert open(out, "rb").read() == bytes(z[f"{name}/csv"])


def test_smoothing_is_bit_exact_given_the_same_table(golden_dir):
    """Feed K6 the reference's own initial table (exact float32 means): results must be identical bit for bit."""
    from cutdet import engine
    for name, (scores, k1, kb) in kat_inputs.segmentation_cases().items():
        init = {k: torch.from_numpy(v) for k, v in _kat(golden_dir, name, "init").items()}
        t = engine.DeviceRunTable.from_te(init, "cuda")
        t.glue_orphans(k1, kb)
        got = _np(t.to_te())
        want = _kat(golden_dir, name, "glued")
        for k in INT_COLS:
            assert np.array_equal(got[k], want[k]), (name, k)
        assert np.array_equal(got["score_means"], want["score_means"]), name
        t.combine_adjacent()
        got = _np(t.to_te())
        want = _kat(golden_dir, name, "combined")
        for k in INT_COLS:
            assert np.array_equal(got[k], want[k]), (name, k)
        assert np.array_equal(got["score_means"], want["score_means"]), name


def test_lone_orphan_raises_index_error():
    from frameID.segmentation import Segmentation
    seg = Segmentation(torch.from_numpy(kat_inputs.scores_from_runs([(0, 50)], 13)))
    with pytest.raises(IndexError):
        seg.glue_orphans(100, 10)


@pytest.mark.parametrize("seed", range(6))
def test_random_against_oracle(seed):
    """Bigger random tables than the fixtures, oracle-checked (tie-free by construction)."""
    from cutdet import engine
    runs = kat_inputs.random_runs(1000 + seed, 150 + 90 * seed)
    scores = kat_inputs.scores_from_runs(runs, 2000 + seed)
    t0, t1, t2, csv = oseg.segment(scores, 100, 10)
    table = engine.run_table_from_scores(torch.from_numpy(scores).cuda())
    _same(_np(table.to_te()), t0)
    table.glue_orphans(100, 10)
    _same(_np(table.to_te()), t1, means_rtol=1e-5)
    table.combine_adjacent()
    _same(_np(table.to_te()), t2, means_rtol=1e-5)


@pytest.mark.parametrize("chunks", [[1], [5, 1, 2048, 3], [2047, 2049, 4096], [100000], [4050] * 5])
def test_streaming_rle_equals_one_shot(chunks):
    """Feeding the encoder chunk by chunk (any split) gives the table of the whole sequence."""
    from cutdet import engine
    n = sum(chunks)
    rng = np.random.default_rng(n)
    # mix of long runs and noise so that runs straddle chunk and tile edges
    lab = np.repeat(rng.integers(0, 3, n // 37 + 2), 37)[:n].astype(np.uint8)
    noisy = rng.uniform(size=n) < 0.02
    lab[noisy] = rng.integers(0, 3, noisy.sum())
    top = rng.uniform(1, 9, n).astype(np.float32)
    want = oseg.run_table_from_labels(lab, top)
    enc = engine.RunLengthEncoder(n, "cuda")
    dl, dt = torch.from_numpy(lab).cuda(), torch.from_numpy(top).cuda()
    pos = 0
    for c in chunks:
        enc.append(dl[pos:pos + c], dt[pos:pos + c])
        pos += c
    _same(_np(enc.finish().to_te()), want)


def test_full_game_size_properties():
    """BASELINE config 2 size (324,000 frames): properties that need no oracle run."""
    from cutdet import engine
    n = 324_000
    rng = np.random.default_rng(1)
    lab = np.repeat(rng.integers(0, 3, n // 50 + 2), 50)[:n].astype(np.uint8)
    lab[rng.uniform(size=n) < 0.01] = 2
    top = rng.uniform(1, 9, n).astype(np.float32)
    enc = engine.RunLengthEncoder(n, "cuda")
    enc.append(torch.from_numpy(lab).cuda(), torch.from_numpy(top).cuda())
    table = enc.finish()
    te = _np(table.to_te())
    # decode(encode(x)) == x ; lengths partition the frames ; means are length-weighted consistent
    assert np.array_equal(np.repeat(te["frame_types"], te["run_lengths"]).astype(np.uint8), lab)
    assert te["run_lengths"].sum() == n and te["start_frames"][0] == 0 and te["end_frames"][-1] == n - 1
    assert np.all(te["frame_types"][1:] != te["frame_types"][:-1])
    assert abs(float((te["score_means"].astype(np.float64) * te["run_lengths"]).sum()) - float(top.astype(np.float64).sum())) < 1e-2 * n / 1000
    table.glue_orphans(100, 10)
    table.combine_adjacent()
    te = _np(table.to_te())
    # idempotence + postconditions of the two passes
    assert te["run_lengths"].sum() == n and np.all(te["frame_types"][1:] != te["frame_types"][:-1])
    real = te["frame_types"] != 2
    assert np.all(te["run_lengths"][real] >= 100) and np.all(te["run_lengths"][~real] >= 10)
    before = {k: v.copy() for k, v in te.items()}
    table.glue_orphans(100, 10)
    table.combine_adjacent()
    after = _np(table.to_te())
    for k in INT_COLS:
        assert np.array_equal(before[k], after[k])


def test_stitch_shards_equals_whole():
    """Per-shard tables (local frame numbers) joined across shard edges == table of the whole sequence."""
    from cutdet import engine
    n = 30_000
    rng = np.random.default_rng(5)
    lab = np.repeat(rng.integers(0, 3, n // 61 + 2), 61)[:n].astype(np.uint8)
    top = rng.uniform(1, 9, n).astype(np.float32)
    want = oseg.run_table_from_labels(lab, top)
    for n_shards in (1, 2, 3, 8):
        per = -(-n // n_shards)
        cap = per
        big = engine.DeviceRunTable(cap * n_shards, "cuda")
        counts, offsets = [], []
        for r in range(n_shards):
            lo, hi = r * per, min(n, (r + 1) * per)
            enc = engine.RunLengthEncoder(cap, "cuda")
            enc.append(torch.from_numpy(lab[lo:hi]).cuda(), torch.from_numpy(top[lo:hi]).cuda())
            t = enc.finish()
            c = t.count()
            for col in ("end_frames", "start_frames", "run_lengths", "frame_types", "score_means", "score_sums"):
                getattr(big, col)[r * cap:r * cap + c] = getattr(t, col)[:c]
            counts.append(c)
            offsets.append(lo)
        out = engine.stitch_shards(big, torch.tensor(counts, dtype=torch.int64, device="cuda"),
                                   torch.tensor(offsets, dtype=torch.int64, device="cuda"), cap)
        _same(_np(out.to_te()), want)


def _encode(lab, top, cap=None):
    from cutdet import engine
    enc = engine.RunLengthEncoder(cap or max(len(lab), 1), "cuda")
    if len(lab):
        enc.append(torch.from_numpy(lab).cuda(), torch.from_numpy(top).cuda())
    return enc.finish()


def test_packed_exchange_equals_whole():
    """The exchange step as the multi-GPU path runs it, on one device: cutdet_shard_pack per shard (bytes checked against the
    numpy restatement of the layout, cutdet.shard.pack_columns), the gathered buffer read by cutdet_stitch_packed (offsets and
    counts from the headers, on the device) == the table of the whole sequence.  Shards of every kind: empty, one frame, cut
    inside a run and on a run boundary."""
    from cutdet import shard
    n = 30_000
    rng = np.random.default_rng(5)
    lab = np.repeat(rng.integers(0, 3, n // 61 + 2), 61)[:n].astype(np.uint8)
    top = rng.uniform(1, 9, n).astype(np.float32)
    want = oseg.run_table_from_labels(lab, top)
    for edges in ([0, n], [0, 15_000, n], [0, 61 * 100, 61 * 100 + 1, 20_000, 20_000, n], list(range(0, n, 3_750)) + [n]):
        tables, counts = [], []
        for lo, hi in zip(edges, edges[1:]):
            tables.append(_encode(lab[lo:hi], top[lo:hi]))
            counts.append(hi - lo)
        cap = 1024
        for t, c in zip(tables, counts):
            packed = shard.pack_table(t, c, cap).cpu()
            k = t.count()
            cols = {name: getattr(t, name)[:k].cpu().numpy() for name in
                    ("end_frames", "start_frames", "run_lengths", "score_sums", "frame_types", "score_means")}
            ref = shard.pack_columns(cols, k, c, cap)
            used = shard.HEADER_BYTES + k * shard.ROW_BYTES        # rows past the count are not written by the kernel
            assert torch.equal(packed[:used], ref[:used])
        out, total = shard.stitch_local(tables, counts, cap)
        _same(_np(out.to_te()), want)
        assert int(total.item()) == n


def test_exchange_overflow_is_reported_and_retried():
    """A shard with more runs than the gather capacity: the stitch kernel reports it through a negative count (ShardOverflow
    from to_te), and finish_checked repeats the exchange with a capacity that fits."""
    from cutdet import engine, pipeline, shard
    n = 6_000
    lab = (np.arange(n) // 3 % 3).astype(np.uint8)                 # 2,000 runs
    top = np.linspace(1, 9, n).astype(np.float32)
    halves = [_encode(lab[:3000], top[:3000]), _encode(lab[3000:], top[3000:])]
    out, _ = shard.stitch_local(halves, [3000, 3000], 64)
    with pytest.raises(engine.ShardOverflow) as e:
        out.to_te()
    assert e.value.needed == 1000
    te, total = shard.finish_checked(lambda cap: shard.stitch_local(halves, [3000, 3000], cap), capacity=64)
    _same(_np(te), oseg.run_table_from_labels(lab, top))
    assert total == n


def test_deferred_status_and_independent_workspaces():
    """pipeline.smooth queues K6 without a host synchronisation: the lone-orphan IndexError of the reference surfaces at the
    next to_te().  Two tables smoothed concurrently on two streams own their scratch and do not disturb each other."""
    from cutdet import engine, pipeline
    lone = engine.run_table_from_scores(torch.from_numpy(kat_inputs.scores_from_runs([(0, 50)], 13)).cuda())
    pipeline.smooth(lone, 100, 10)
    with pytest.raises(IndexError):
        lone.to_te()
    cases = []
    for seed in (0, 1):
        scores = kat_inputs.scores_from_runs(kat_inputs.random_runs(50 + seed, 1500), 60 + seed)
        cases.append((engine.run_table_from_scores(torch.from_numpy(scores).cuda()), oseg.segment(scores, 100, 10)[2]))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for _ in range(3):
        for (table, _), st in zip(cases, streams):
            with torch.cuda.stream(st):
                pipeline.smooth(table, 100, 10)
    torch.cuda.synchronize()
    for table, want in cases:
        _same(_np(table.to_te()), want, means_rtol=1e-5)


def test_nan_mean_orphans_sort_last():
    """torch.argsort places NaN after every number: an orphan whose mean is NaN is still glued, after the others
    (oracle.segmentation.glue_orphans restates that order)."""
    from cutdet import engine
    runs = [(0, 300), (1, 20), (0, 150), (2, 4), (1, 200), (0, 30), (1, 400)]
    scores = kat_inputs.scores_from_runs(runs, 3)
    te0 = oseg.run_table(scores)
    te0["score_means"] = te0["score_means"].copy()
    te0["score_means"][1] = np.nan
    want = oseg.combine_adjacent(oseg.glue_orphans(te0, 100, 10))
    t = engine.DeviceRunTable.from_te({k: torch.from_numpy(np.asarray(v)) for k, v in te0.items()}, "cuda")
    t.glue_orphans(100, 10)
    t.combine_adjacent()
    got = _np(t.to_te())
    for k in INT_COLS:
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(np.isnan(got["score_means"]), np.isnan(want["score_means"]))


def _nccl_worker(rank, world, port, n, seed, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from cutdet import engine, pipeline, shard
        rng = np.random.default_rng(seed)
        lab = np.repeat(rng.integers(0, 3, n // 53 + 2), 53)[:n].astype(np.uint8)
        lab[rng.uniform(size=n) < 0.01] = 2
        top = rng.uniform(1, 9, n).astype(np.float32)
        lo, hi = shard.shard_range(n, rank, world)
        enc = engine.RunLengthEncoder(max(hi - lo, 1), f"cuda:{rank}")
        if hi > lo:
            enc.append(torch.from_numpy(lab[lo:hi]).cuda(), torch.from_numpy(top[lo:hi]).cuda())
        local = enc.finish()
        # 1,024 rows do not hold this shard's ~2,600 runs: the stitch kernel reports it and finish_checked repeats the exchange
        # (pack, NCCL all-gather, stitch) with a capacity that fits -- on every rank alike, so the collectives stay matched
        raw, total = shard.finish_checked(lambda cap: shard.stitch_all(local, hi - lo, cap), capacity=1024)
        te, _ = shard.finish_checked(lambda cap: shard.stitch_all(local, hi - lo, cap), after=lambda t: pipeline.smooth(t, 100, 10),
                                     capacity=1024)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), total=total,
                 **{"raw_" + k: v.numpy() for k, v in raw.items()}, **{k: v.numpy() for k, v in te.items()})
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_nccl_exchange_on_real_ranks(tmp_path, world):
    """Two processes, one GPU each, NCCL: every rank's stitched and smoothed table == what one process computes over the
    whole sequence (the contract of the N > 1 path; reference frameID/segmentation.py:35-60, 91-183)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()} (the 1-GPU emulation of the same kernels "
                    "is test_packed_exchange_equals_whole; bench.py checks the same equality at N = 2/4/8)")
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n, seed = 200_003, 17
    mp.spawn(_nccl_worker, args=(world, port, n, seed, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(seed)
    lab = np.repeat(rng.integers(0, 3, n // 53 + 2), 53)[:n].astype(np.uint8)
    lab[rng.uniform(size=n) < 0.01] = 2
    top = rng.uniform(1, 9, n).astype(np.float32)
    whole = _encode(lab, top)
    want_raw = _np(whole.to_te())
    whole.glue_orphans(100, 10)
    whole.combine_adjacent()
    want = _np(whole.to_te())
    for rank in range(world):
        got = np.load(os.path.join(tmp_path, f"rank{rank}.npz"))
        assert int(got["total"]) == n
        for k in INT_COLS:
            assert np.array_equal(got["raw_" + k], want_raw[k]), (rank, k)
            assert np.array_equal(got[k], want[k]), (rank, k)
        assert np.allclose(got["score_means"], want["score_means"], rtol=1e-5)
