#!/usr/bin/env python3
"""Record what the REFERENCE produces for the seeded inputs of tests/kat_inputs.py.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's ``frameID`` package unmodified from /root/reference and writes
  tests/golden/net_kat.npz            logits of load_default_net() (eager and torch.jit.trace'd per
                                      training_scripts/make_torchscript_model.py:25-27)
  tests/golden/preprocess_video.npz   decoded frames of a small synthetic clip + what VideoDataset(resize=256)
                                      yields for them (frameID/data.py:184-234)
  tests/golden/resize_hashes.json     sha256 of cv2.resize outputs for seeded full-size frames
  tests/golden/segmentation_kat.npz   Segmentation tables after __init__ / glue_orphans / combine_adjacent_segments
                                      and the CSV bytes (frameID/segmentation.py)
  cut-detection_b200/frameID/prod_net/prod_net_weights.npz
                                      the shipped prod_net parameters re-encoded as plain float32 arrays
                                      (state_dict keys kept, 'conv.'/'linear.' prefixed) + the params JSON
No reference source is copied; only outputs are stored.
"""
import hashlib
import io
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402
import torch  # noqa: E402

from frameID.net import load_default_net  # noqa: E402  (the reference)
from frameID.data import VideoDataset  # noqa: E402
from frameID.segmentation import Segmentation  # noqa: E402
import frameID.net as ref_net_mod  # noqa: E402

import kat_inputs  # noqa: E402
from oracle import preprocess as opre  # noqa: E402

assert ref_net_mod.__file__.startswith(REF), ref_net_mod.__file__
torch.manual_seed(0)


def te_to_np(te):
    return {k: v.numpy().copy() for k, v in te.items()}


def main():
    # ------------------------------------------------------------------ weights
    net, params = load_default_net()
    net.eval()
    blob = {}
    for prefix, mod in (("conv.", net[0]), ("linear.", net[1])):
        for k, v in mod.state_dict().items():
            if k.endswith("num_batches_tracked"):
                continue
            blob[prefix + k] = v.detach().numpy().astype(np.float32)
    blob["__params_json__"] = np.frombuffer(json.dumps(params).encode("utf-8"), dtype=np.uint8)
    wdir = os.path.join(ROOT, "cut-detection_b200", "frameID", "prod_net")
    os.makedirs(wdir, exist_ok=True)
    np.savez(os.path.join(wdir, "prod_net_weights.npz"), **blob)
    print("weights:", sum(v.size for k, v in blob.items() if not k.startswith("__")), "parameters")

    # ------------------------------------------------------------------ net known answers
    traced = torch.jit.trace(net, torch.randn([1, 3, 144, 256]))
    out = {}
    with torch.no_grad():
        for name, frame in kat_inputs.kat_frames().items():
            x = torch.from_numpy(opre.preprocess_frame(frame)[None])
            # cross-check the oracle's preprocessing against cv2 + the reference's tensor ops
            nw, nh = opre.target_size(frame.shape[1], frame.shape[0])
            r = cv2.resize(frame, (nw, nh), interpolation=cv2.INTER_LINEAR)
            t = torch.flip(torch.tensor(r, dtype=torch.float).permute(2, 0, 1), (0,)) / 255
            assert torch.equal(t, x[0]), name
            out["frame_" + name] = net(x).numpy()[0]
            assert torch.equal(net(x), traced(x)), name
        xs = torch.from_numpy(kat_inputs.smooth_images(48))
        out["smooth48_eager"] = net(xs).numpy()
        out["smooth48_traced"] = traced(xs).numpy()
        # per-layer activations of the first 2 smooth images (pins the layer order conv->relu->pool->bn)
        y = xs[:2]
        for i, layer in enumerate(net[0].conv_layers):
            y = layer(y)
            out[f"smooth2_layer{i}"] = y.numpy().copy()
        out["smooth2_features"] = net[0](xs[:2]).numpy()
    np.savez_compressed(os.path.join(HERE, "net_kat.npz"), **out)
    print("net_kat: smooth48 argmax histogram", np.bincount(out["smooth48_eager"].argmax(1), minlength=3))

    # ------------------------------------------------------------------ VideoDataset on a small clip
    h, w, n = 270, 480, 6
    rng = np.random.default_rng(3)
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "clip.mp4")
        vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 30, (w, h))
        for i in range(n):
            f = np.stack([(xx * (i + 1) // 2 + yy) % 256, (yy * 2 + 17 * i) % 256,
                          ((xx // 24 + yy // 24 + i) % 2) * 200], axis=-1).astype(np.uint8)
            f[40:80, 60:200] = rng.integers(0, 256, (40, 140, 3), dtype=np.uint8)
            vw.write(f)
        vw.release()
        ds = VideoDataset(path, resize=256)
        info = dict(ds.video_info)
        tensors = [t.numpy().copy() for t in ds]
        cap = cv2.VideoCapture(path)
        decoded = []
        while True:
            ok, fr = cap.read()
            if not ok:
                break
            decoded.append(fr)
        cap.release()
    decoded = np.stack(decoded)
    tensors = np.stack(tensors)
    assert len(decoded) == len(tensors) == n
    as_u8 = np.rint(tensors * 255).astype(np.uint8)
    assert np.array_equal(as_u8.astype(np.float32) / np.float32(255), tensors)
    np.savez_compressed(os.path.join(HERE, "preprocess_video.npz"), decoded_bgr=decoded,
                        dataset_rgb_chw_u8=as_u8, info=np.frombuffer(json.dumps(info).encode(), dtype=np.uint8))
    print("preprocess_video:", decoded.shape, "->", tensors.shape, info)

    # ------------------------------------------------------------------ resize hashes at full size
    hashes = {}
    for (ww, hh) in [(1280, 720), (1920, 1080), (3840, 2160), (640, 360), (854, 480), (1920, 800), (512, 288),
                     (256, 144), (320, 180), (1000, 562)]:
        f = np.random.default_rng(ww * 10007 + hh).integers(0, 256, (hh, ww, 3), dtype=np.uint8)
        nw, nh = opre.target_size(ww, hh)
        r = cv2.resize(f, (nw, nh), interpolation=cv2.INTER_LINEAR)
        hashes[f"{ww}x{hh}"] = {"out": [nw, nh], "sha256": hashlib.sha256(r.tobytes()).hexdigest()}
    with open(os.path.join(HERE, "resize_hashes.json"), "w") as fjs:
        json.dump({"cv2": cv2.__version__, "hashes": hashes}, fjs, indent=1)

    # ------------------------------------------------------------------ segmentation
    seg_out = {}
    for name, (scores, k1, kb) in kat_inputs.segmentation_cases().items():
        seg = Segmentation(torch.from_numpy(scores))
        for k, v in te_to_np(seg.te).items():
            seg_out[f"{name}/init/{k}"] = v
        seg.glue_orphans(k1, kb)
        for k, v in te_to_np(seg.te).items():
            seg_out[f"{name}/glued/{k}"] = v
        seg.combine_adjacent_segments()
        for k, v in te_to_np(seg.te).items():
            seg_out[f"{name}/combined/{k}"] = v
        with tempfile.NamedTemporaryFile(suffix=".csv") as tf:
            seg.write_csv(tf.name)
            seg_out[f"{name}/csv"] = np.frombuffer(open(tf.name, "rb").read(), dtype=np.uint8)
        # no exact ties among run means at construction (the unstable argsort would make the
        # reference's answer build-dependent)
        m = te_to_np(Segmentation(torch.from_numpy(scores)).te)["score_means"]
        assert len(np.unique(m)) == len(m), name
        print(f"seg {name}: {len(m)} runs -> {len(seg)} segments")
    # error behaviour: a lone orphan run raises IndexError (segmentation.py:112-113)
    lone = kat_inputs.scores_from_runs([(0, 50)], 13)
    try:
        Segmentation(torch.from_numpy(lone)).glue_orphans(100, 10)
        raised = False
    except IndexError:
        raised = True
    seg_out["lone_orphan/raises_index_error"] = np.array([raised])
    np.savez_compressed(os.path.join(HERE, "segmentation_kat.npz"), **seg_out)


if __name__ == "__main__":
    main()
