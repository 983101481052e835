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
