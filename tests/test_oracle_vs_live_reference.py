"""Differential check of the CPU oracle against the UNMODIFIED reference run live (build container only).

/root/reference does not exist on the GPU box, so everything here is skipped there; the committed golden vectors
(tests/golden/, written by make_golden.py from the same reference) are what travels.  Here, where the reference is
importable, random inputs widen the pin beyond the recorded cases: `Segmentation` / `glue_orphans` /
`combine_adjacent_segments` / `write_csv` (segmentation.py:35-196) on seeded random run structures and thresholds.
The reference runs in a fresh process whose only extra path is /root/reference (this repo ships a `frameID` mirror of
the same name, which must not shadow it)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import kat_inputs
from oracle import segmentation as oseg

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "frameID")), reason="the reference tree is not on this machine")

COLS = ("end_frames", "frame_types", "run_lengths", "start_frames", "score_means")

_REF_SCRIPT = textwrap.dedent("""
    import sys, tempfile, os
    import numpy as np, torch
    sys.path.insert(0, %r)
    from frameID.segmentation import Segmentation
    import frameID
    assert frameID.__path__[0].startswith(%r), frameID.__path__
    inp = np.load(sys.argv[1])
    out = {}
    n = int(inp["n_cases"])
    for i in range(n):
        scores = torch.from_numpy(inp[f"{i}/scores"])
        k1, kb = (int(v) for v in inp[f"{i}/thresholds"])
        seg = Segmentation(scores)
        def dump(tag):
            for k, v in seg.te.items():
                out[f"{i}/{tag}/{k}"] = v.numpy().copy()
        dump("init")
        try:
            seg.glue_orphans(k1, kb)
        except IndexError:
            out[f"{i}/raised"] = np.array([1])
            continue
        dump("glued")
        seg.combine_adjacent_segments()
        dump("combined")
        assert len(seg) == len(seg.te["end_frames"])
        with tempfile.TemporaryDirectory() as d:
            p = os.path.join(d, "s.csv")
            seg.write_csv(p)
            out[f"{i}/csv"] = np.frombuffer(open(p, "rb").read(), np.uint8)
    np.savez(sys.argv[2], **out)
""") % (REF, REF)


def _cases():
    cases = []
    rng = np.random.default_rng(2026)
    for i in range(40):
        n_runs = int(rng.integers(1, 80))
        runs = kat_inputs.random_runs(5000 + i, n_runs)
        k1 = int(rng.choice([100, 100, 40, 250, 1]))
        kb = int(rng.choice([10, 10, 25, 1, 60]))
        cases.append((kat_inputs.scores_from_runs(runs, 6000 + i), k1, kb))
    for i in range(6):          # frame-level noise: runs of one or two frames, hundreds of merges
        cases.append((rng.normal(0, 1, (int(rng.integers(50, 900)), 3)).astype(np.float32), 100, 10))
    cases.append((kat_inputs.scores_from_runs([(2, 5)], 1), 100, 10))                 # a lone orphan: IndexError
    cases.append((kat_inputs.scores_from_runs([(1, 1)], 2), 100, 10))                 # one frame
    cases.append((kat_inputs.scores_from_runs([(0, 99), (1, 99)], 3), 100, 10))       # two orphans, nothing else
    return cases


def test_segmentation_equals_live_reference(tmp_path):
    cases = _cases()
    inp = {"n_cases": np.array(len(cases))}
    for i, (s, k1, kb) in enumerate(cases):
        inp[f"{i}/scores"] = s
        inp[f"{i}/thresholds"] = np.array([k1, kb])
    src, dst = str(tmp_path / "in.npz"), str(tmp_path / "out.npz")
    np.savez(src, **inp)
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", _REF_SCRIPT, src, dst], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    ref = np.load(dst)
    n_raised = 0
    for i, (s, k1, kb) in enumerate(cases):
        t0 = oseg.run_table(s)
        for k in COLS:
            assert np.array_equal(t0[k], ref[f"{i}/init/{k}"]), (i, "init", k)
        if f"{i}/raised" in ref:
            n_raised += 1
            with pytest.raises(IndexError):
                oseg.glue_orphans(t0, k1, kb)
            continue
        t0, t1, t2, csv = oseg.segment(s, k1, kb)
        for tag, te in (("glued", t1), ("combined", t2)):
            for k in COLS:
                assert np.array_equal(te[k], ref[f"{i}/{tag}/{k}"]), (i, tag, k)       # score_means bit for bit (float32)
                assert te[k].dtype == ref[f"{i}/{tag}/{k}"].dtype, (i, tag, k)
        assert csv == ref[f"{i}/csv"].tobytes(), i
    assert n_raised >= 1


_NET_SCRIPT = textwrap.dedent("""
    import sys, json
    import numpy as np, torch
    sys.path.insert(0, %r)
    from frameID.net import FrameConvNet, FrameLinearNet
    inp = np.load(sys.argv[1])
    archs = json.loads(bytes(inp["archs"]).decode())
    out = {}
    for i, a in enumerate(archs):
        conv = FrameConvNet(3, a["hidden"], a["conv_layers"], a["avg_pool"])
        lin = FrameLinearNet(a["linear_layers"], a["hidden"] * a["avg_pool"] ** 2, a["linear_size"], a["out"])
        sd_c = {k[len(f"{i}/conv."):]: torch.from_numpy(inp[k]) for k in inp.files if k.startswith(f"{i}/conv.")}
        sd_l = {k[len(f"{i}/linear."):]: torch.from_numpy(inp[k]) for k in inp.files if k.startswith(f"{i}/linear.")}
        missing = conv.load_state_dict(sd_c, strict=False)
        assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing
        missing = lin.load_state_dict(sd_l, strict=False)
        assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing
        net = torch.nn.Sequential(conv, lin).eval()
        with torch.no_grad():
            out[f"{i}/y"] = net(torch.from_numpy(inp[f"{i}/x"])).numpy()
            y = torch.from_numpy(inp[f"{i}/x"])
            for j, layer in enumerate(conv.conv_layers):          # CNNLayer.forward, net.py:34-40
                y = layer(y)
                out[f"{i}/map{j}"] = y.numpy().copy()
        out[f"{i}/num_params"] = np.array([conv.num_params(), lin.num_params()])
    np.savez(sys.argv[2], **out)
""") % (REF,)


def test_other_architectures_equal_live_reference(tmp_path):
    """FrameConvNet / FrameLinearNet (net.py:71-189) built by the reference's own constructors for architectures other than the
    shipped one -- the contrastive encoder of learn_contrasts.py:68-76, a two-layer trunk, odd resized heights (1920x800 -> 106 rows,
    854x480 -> 143) -- loaded with the oracle's random weights: logits, every layer's map and parameter counts against the oracle."""
    import json
    from oracle import net as onet
    archs = [
        dict(hidden=32, conv_layers=3, avg_pool=1, linear_layers=3, linear_size=32, out=8, h=144, w=256, batch=4),
        dict(hidden=16, conv_layers=2, avg_pool=2, linear_layers=1, linear_size=16, out=5, h=106, w=256, batch=3),
        dict(hidden=48, conv_layers=3, avg_pool=4, linear_layers=2, linear_size=32, out=3, h=143, w=256, batch=2),
        dict(hidden=8, conv_layers=4, avg_pool=1, linear_layers=2, linear_size=12, out=3, h=144, w=256, batch=2),
    ]
    inp = {"archs": np.frombuffer(json.dumps(archs).encode(), np.uint8)}
    ws, xs = [], []
    for i, a in enumerate(archs):
        w = onet.random_weights(40 + i, a["hidden"], a["conv_layers"], a["avg_pool"], a["linear_layers"], a["linear_size"], a["out"])
        x = kat_inputs.smooth_images(a["batch"], seed=50 + i, h=a["h"], w=a["w"])
        for k, v in w.items():
            inp[f"{i}/{k}"] = v
        inp[f"{i}/x"] = x
        ws.append(w)
        xs.append(x)
    src, dst = str(tmp_path / "in.npz"), str(tmp_path / "out.npz")
    np.savez(src, **inp)
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", _NET_SCRIPT, src, dst], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    ref = np.load(dst)
    for i, a in enumerate(archs):
        y, feat = onet.forward_f32(ws[i], xs[i], a["avg_pool"], return_features=True)
        assert y.shape == ref[f"{i}/y"].shape == (a["batch"], a["out"])
        assert len(feat) == a["conv_layers"]
        for j, m in enumerate(feat):
            assert m.shape == ref[f"{i}/map{j}"].shape, (i, j, m.shape)
            assert float(np.abs(m - ref[f"{i}/map{j}"]).max()) <= 1e-5, (i, j, float(np.abs(m - ref[f"{i}/map{j}"]).max()))
        assert float(np.abs(y - ref[f"{i}/y"]).max()) <= 1e-5, (i, float(np.abs(y - ref[f"{i}/y"]).max()))
        n_conv = sum(v.size for k, v in ws[i].items() if k.startswith("conv.") and "running" not in k)
        n_lin = sum(v.size for k, v in ws[i].items() if k.startswith("linear.") and "running" not in k)
        assert ref[f"{i}/num_params"].tolist() == [n_conv, n_lin]
