"""Property test of the shard stitch: for ANY split of a frame sequence into consecutive time shards -- empty
shards, one-frame shards, cuts inside a run and exactly on a run boundary -- the stitched per-shard run tables equal the
run table `Segmentation.__init__` (segmentation.py:35-60) builds from the whole sequence, and glue/combine on the stitched
table give the same segments.  Pure CPU (the oracle's stitch is the checker of the CUDA stitch kernel in the -m gpu suite)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import segmentation as oseg

INT_COLS = ("end_frames", "start_frames", "run_lengths", "frame_types")


def _local(lab, top):
    if lab.size == 0:
        return {k: np.zeros(0, np.int64) for k in INT_COLS} | {"score_means": np.zeros(0, np.float32), "score_sums": np.zeros(0, np.float64)}
    te = oseg.run_table_from_labels(lab, top)
    te["score_sums"] = np.array([np.sum(top[s:e + 1], dtype=np.float64) for s, e in zip(te["start_frames"], te["end_frames"])], np.float64)
    return te


@st.composite
def _sequence_and_cuts(draw):
    n_runs = draw(st.integers(1, 25))
    labs, prev = [], -1
    for _ in range(n_runs):
        lab = draw(st.integers(0, 2))
        length = draw(st.sampled_from([1, 2, 3, 9, 10, 11, 40, 99, 100, 101, 350]))
        labs += [lab] * length          # equal neighbours are allowed: they form one run, as in the reference
    n = len(labs)
    world = draw(st.integers(1, 6))
    cuts = sorted(draw(st.lists(st.integers(0, n), min_size=world - 1, max_size=world - 1)))
    seed = draw(st.integers(0, 2 ** 16))
    return np.array(labs, np.uint8), [0] + cuts + [n], seed


@settings(max_examples=120, deadline=None)
@given(_sequence_and_cuts())
def test_stitch_of_any_split_equals_whole(case):
    lab, edges, seed = case
    top = np.random.default_rng(seed).uniform(1, 9, lab.size).astype(np.float32)
    want = oseg.run_table_from_labels(lab, top)
    shards = [_local(lab[a:b], top[a:b]) for a, b in zip(edges, edges[1:])]
    got = oseg.stitch_tables(shards, edges[:-1])
    for k in INT_COLS:
        assert np.array_equal(got[k], want[k]), k
    np.testing.assert_allclose(got["score_means"], want["score_means"], rtol=2e-6)
    assert int(got["run_lengths"].sum()) == lab.size
    # the order-dependent passes run once on the stitched table: same segments as on the whole sequence (the stitched means can
    # differ from .mean() in the last bit, which only matters if two orphan means are within rounding of each other)
    stitched = {k: got[k] for k in INT_COLS} | {"score_means": got["score_means"]}

    def segments(te):
        try:
            return oseg.csv_bytes(oseg.combine_adjacent(oseg.glue_orphans(te, 100, 10)))
        except IndexError:               # everything merged into ONE run that is still an orphan (segmentation.py:110-113)
            return "IndexError"

    assert segments(stitched) == segments(want)
