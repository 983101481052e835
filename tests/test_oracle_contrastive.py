"""The CPU restatement of the contrastive forward pass (oracle/contrastive.py) against what the unmodified reference
produced for the same seeded inputs (tests/golden/contrastive_kat.npz, written by tests/golden/make_contrastive_golden.py)."""
import os

import numpy as np
import pytest

import kat_inputs
from oracle import contrastive as ocon


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "contrastive_kat.npz"))


def kat_input(kat):
    n = int(kat["pairs"])
    return np.concatenate([kat_inputs.smooth_images(n, seed=21), kat_inputs.smooth_images(n, seed=22)])


def test_batchstat_forward_matches_reference(kat):
    w = {k: kat[k] for k in kat.files if k.startswith(("conv.", "linear."))}
    x = kat_input(kat)
    inter = ocon.forward_batchstats({k: v for k, v in w.items() if k.startswith("conv.")}, x, 1)
    assert inter.shape == kat["intermediate"].shape
    assert np.abs(inter - kat["intermediate"]).max() <= 2e-4
    proj = ocon.forward_batchstats({k: v for k, v in w.items() if k.startswith("linear.")}, kat["intermediate"], 1)
    assert np.abs(proj - kat["projection"]).max() <= 2e-4


def test_contrastive_loss_matches_reference(kat):
    loss, ab = ocon.contrastive_loss(kat["projection"])
    assert abs(float(loss) - float(kat["loss"])) <= 2e-5
    assert np.abs(ab - kat["logits_ab"]).max() <= 2e-6
    loss, ab = ocon.contrastive_loss(kat["projection"], temperature=0.5, h_norm=False)
    assert abs(float(loss) - float(kat["loss_t05_nonorm"])) <= 5e-5
    assert np.abs(ab - kat["logits_ab_t05_nonorm"]).max() <= 1e-4 * np.abs(kat["logits_ab_t05_nonorm"]).max()


def test_contrastive_loss_properties():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(16, 8)).astype(np.float32)
    loss, ab = ocon.contrastive_loss(x)
    # swapping the two views swaps loss_a and loss_b: the total is unchanged and logits_ab is transposed
    loss2, ab2 = ocon.contrastive_loss(np.concatenate([x[8:], x[:8]]))
    assert abs(float(loss) - float(loss2)) <= 1e-6
    assert np.allclose(ab2, ab.T, atol=1e-6)
    # identical views of well separated points: every positive is the best match, the loss is below the uniform bound 2 log(2B - 1)
    y = np.eye(8, dtype=np.float32) * 50
    low, _ = ocon.contrastive_loss(np.concatenate([y, y]), temperature=0.05)
    assert float(low) < 1e-3 < 2 * np.log(15)
