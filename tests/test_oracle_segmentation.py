"""Oracle segmentation vs tables and CSV bytes recorded from the reference's Segmentation class."""
import os

import numpy as np
import pytest

import kat_inputs
from oracle import segmentation as oseg

INT_COLS = ("end_frames", "frame_types", "run_lengths", "start_frames")


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "segmentation_kat.npz"))


def _check(te, kat, prefix):
    for k in INT_COLS:
        assert np.array_equal(te[k], kat[f"{prefix}/{k}"]), (prefix, k)
        assert te[k].dtype == np.int64
    assert np.array_equal(te["score_means"], kat[f"{prefix}/score_means"]), prefix   # float32, bit for bit


@pytest.mark.parametrize("name", sorted(kat_inputs.segmentation_cases().keys()))
def test_case(name, kat):
    scores, k1, kb = kat_inputs.segmentation_cases()[name]
    t0, t1, t2, csv = oseg.segment(scores, k1, kb)
    _check(t0, kat, f"{name}/init")
    _check(t1, kat, f"{name}/glued")
    _check(t2, kat, f"{name}/combined")
    assert csv == bytes(kat[f"{name}/csv"])


def test_survey_kat_csv(kat):
    scores, k1, kb = kat_inputs.segmentation_cases()["survey_kat"]
    assert oseg.segment(scores, k1, kb)[3] == b"0,a22\r\n300,b\r\n320,ez\r\n815,b\r\n827,a22\r\n"


def test_lone_orphan_raises(kat):
    assert bool(kat["lone_orphan/raises_index_error"][0])
    lone = kat_inputs.scores_from_runs([(0, 50)], 13)
    with pytest.raises(IndexError):
        oseg.glue_orphans(oseg.run_table(lone), 100, 10)


def test_mean_update_quirk():
    """(m_n*l_n + m_o*l_o) / l_n + l_o -- divides by the neighbour only, then adds the orphan length."""
    te = {"start_frames": np.array([0, 300]), "end_frames": np.array([299, 319]),
          "run_lengths": np.array([300, 20]), "frame_types": np.array([0, 1]),
          "score_means": np.array([5.0, 4.0], np.float32)}
    out = oseg.glue_orphans(te, 100, 10)
    assert len(out["start_frames"]) == 1 and out["frame_types"][0] == 0
    assert out["score_means"][0] == np.float32((5.0 * 300 + 4.0 * 20) / 300 + 20)
    assert out["run_lengths"][0] == 320


def test_first_index_wins_argmax_ties():
    s = np.array([[1.0, 1.0, 0.0], [0.0, 2.0, 2.0]], np.float32)
    top, lab = oseg.max_and_argmax(s)
    assert lab.tolist() == [0, 1] and top.tolist() == [1.0, 2.0]
