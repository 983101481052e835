"""Oracle network forward vs logits recorded from the reference's load_default_net()
(tests/golden/net_kat.npz) and the known-answer table of SURVEY.md section 8c."""
import os

import numpy as np
import pytest

import kat_inputs
from oracle import net as onet
from oracle import preprocess as opre

SURVEY_KAT = {   # SURVEY.md 8c, torch 2.11 CPU fp32, tolerance 1e-4
    "black_720": [-0.710721, -2.691549, 4.131387],
    "white_720": [4.336324, -0.971733, -3.396681],
    "hstripes_720": [11.063028, -7.746481, -8.785885],
    "vstripes_720": [-9.503363, 15.934620, -15.890486],
    "hstripes_1080": [11.063028, -7.746481, -8.785885],
    "vstripes_1080": [-9.503363, 15.934620, -15.890486],
    "noise_720": [7.636413, -4.433383, -4.424471],
    "noise_1080": [3.429621, -0.558517, -3.379377],
    "blue_720": [5.364764, -1.603385, -4.988451],
}


@pytest.fixture(scope="module")
def kat(golden_dir):
    return np.load(os.path.join(golden_dir, "net_kat.npz"))


def test_weight_fixture_shapes(prod_weights):
    w, params = prod_weights
    assert params["conv_channels"] == 48 and params["avg_pool_size"] == 4
    assert w["conv.conv_layers.0.conv.weight"].shape == (48, 3, 3, 3)
    assert w["conv.conv_layers.2.conv.weight"].shape == (48, 48, 3, 3)
    assert w["linear.layers.0.linear.weight"].shape == (32, 768)
    assert w["linear.layers.1.linear.weight"].shape == (3, 32)
    assert sum(v.size for v in w.values()) == 43200 + 24771 + 4 * 48 * 3 // 2 + 64  # + BN running stats


def test_frames_known_answers(prod_weights, kat):
    w, params = prod_weights
    for name, frame in kat_inputs.kat_frames().items():
        x = opre.preprocess_frame(frame)[None]
        got = onet.forward_f32(w, x, params["avg_pool_size"])[0]
        assert np.array_equal(got, kat["frame_" + name]), name          # bit-identical to the reference
        assert np.allclose(got, SURVEY_KAT[name], atol=1e-4), name


def test_smooth_batch_bit_identical(prod_weights, kat):
    w, params = prod_weights
    x = kat_inputs.smooth_images(48)
    got, feats = onet.forward_f32(w, x, params["avg_pool_size"], return_features=True)
    assert np.array_equal(kat["smooth48_eager"], kat["smooth48_traced"])
    # batch-48 vs recorded batch-48: same aten kernels, same blocking -> identical
    assert np.allclose(got, kat["smooth48_eager"], atol=2e-5, rtol=0)
    assert np.array_equal(got.argmax(1), kat["smooth48_eager"].argmax(1))
    got2, feats2 = onet.forward_f32(w, x[:2], params["avg_pool_size"], return_features=True)
    for i in range(3):
        assert np.allclose(feats2[i], kat[f"smooth2_layer{i}"], atol=2e-5, rtol=0), i
    # all three classes are present in the fixture
    assert set(np.unique(got.argmax(1))) == {0, 1, 2}


def test_f64_restatement_agrees(prod_weights, kat):
    w, params = prod_weights
    x = kat_inputs.smooth_images(48)[:6]
    ref = kat["smooth48_eager"][:6]
    got, feats = onet.forward_f64(w, x, params["avg_pool_size"], return_features=True)
    assert np.abs(got - ref).max() < 5e-5
    assert feats[0].shape == (6, 48, 48, 85) and feats[1].shape == (6, 48, 16, 28) and feats[2].shape == (6, 48, 5, 9)


def test_16bit_emulations_stay_within_the_stated_tolerance(prod_weights, kat):
    """The two CPU restatements of what the tensor-core kernels compute (fp32 accumulators everywhere / fp16 accumulators in
    layer 1 of the fused frames kernel, oracle.net._conv1_acc16) against the logits recorded from the reference: both must sit
    well inside the 0.05 the GPU tests allow (tests/test_gpu_net.py), with labels unchanged where the margin is >= 0.1."""
    w, params = prod_weights
    frames = kat_inputs.kat_frames()
    x = np.stack([opre.preprocess_frame(f) for f in frames.values()])
    ref = np.stack([kat["frame_" + n] for n in frames])
    for acc16 in (False, True):
        got = onet.forward_tc_emulated(w, x, params["avg_pool_size"], conv1_acc16=acc16)
        assert np.abs(got - ref).max() <= 0.02, acc16
        srt = np.sort(ref, axis=1)
        ok = (srt[:, -1] - srt[:, -2]) >= 0.1
        assert np.array_equal(got.argmax(1)[ok], ref.argmax(1)[ok])
    # the fp16-accumulator variant cannot overflow: per-channel scaling bounds every partial sum by 27 * 2^15 * 255 * 2^-24
    white = np.ones((1, 3, 144, 256), np.float32)
    assert np.isfinite(onet.forward_tc_emulated(w, white, params["avg_pool_size"], conv1_acc16=True)).all()


def test_adaptive_windows():
    assert onet.adaptive_windows(5, 4) == [(0, 2), (1, 3), (2, 4), (3, 5)]
    assert onet.adaptive_windows(9, 4) == [(0, 3), (2, 5), (4, 7), (6, 9)]
    assert onet.adaptive_windows(5, 1) == [(0, 5)]


@pytest.mark.skipif(not os.path.isdir("/root/reference/frameID"), reason="reference tree not present")
def test_live_reference_module(prod_weights):
    """Only in the build container: the real reference module, run live, equals the oracle."""
    import subprocess, sys, tempfile, textwrap
    w, params = prod_weights
    x = kat_inputs.smooth_images(5, seed=99)
    want = onet.forward_f32(w, x, params["avg_pool_size"])
    with tempfile.TemporaryDirectory() as td:
        np.save(os.path.join(td, "x.npy"), x)
        code = textwrap.dedent(f"""
            import sys, numpy as np, torch
            sys.path.insert(0, "/root/reference")
            from frameID.net import load_default_net
            net, _ = load_default_net(); net.eval()
            with torch.no_grad():
                y = net(torch.from_numpy(np.load(r"{td}/x.npy"))).numpy()
            np.save(r"{td}/y.npy", y)
        """)
        subprocess.run([sys.executable, "-c", code], check=True)
        got = np.load(os.path.join(td, "y.npy"))
    assert np.array_equal(got, want)


def test_torchscript_export_parameters(golden_dir):
    """The loader of the reference's TorchScript export recovers the shipped architecture and the
    exact parameters -- checked on the CPU against the re-encoded prod_net weights (no forward pass, no GPU)."""
    import numpy as np
    import torch
    from frameID.net import load_torchscript_net, load_default_net
    net, params = load_torchscript_net(os.path.join(golden_dir, "saved_model_trace.pt"))
    ref, ref_params = load_default_net()
    assert params == {k: ref_params[k] for k in params}
    a, b = net.state_dict(), ref.state_dict()
    assert a.keys() == b.keys()
    for k in a:
        if k.endswith("num_batches_tracked"):        # a training counter (18762 in the export), not a parameter of the forward pass
            continue
        assert torch.equal(a[k], b[k]), k
