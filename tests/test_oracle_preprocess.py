"""Oracle preprocessing vs (a) cv2.resize itself, every pixel, (b) hashes and VideoDataset outputs
recorded from the reference (tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import preprocess as opre

GEOMETRIES = [(1280, 720), (1920, 1080), (3840, 2160), (640, 360), (854, 480), (1920, 800), (512, 288),
              (256, 144), (320, 180), (1000, 562)]


def _frame(w, h):
    return np.random.default_rng(w * 10007 + h).integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("w,h", GEOMETRIES)
def test_resize_matches_recorded_cv2_hash(w, h, golden_dir):
    rec = json.load(open(os.path.join(golden_dir, "resize_hashes.json")))["hashes"][f"{w}x{h}"]
    nw, nh = opre.target_size(w, h)
    assert [nw, nh] == rec["out"]
    out = opre.resize_bilinear_u8(_frame(w, h), nw, nh)
    assert hashlib.sha256(out.tobytes()).hexdigest() == rec["sha256"]


@pytest.mark.parametrize("w,h", [(1280, 720), (1920, 1080), (854, 480), (300, 200), (130, 70), (512, 288)])
def test_resize_matches_live_cv2(w, h):
    cv2 = pytest.importorskip("cv2")
    f = _frame(w, h)
    nw, nh = opre.target_size(w, h)
    ref = cv2.resize(f, (nw, nh), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(ref, opre.resize_bilinear_u8(f, nw, nh))


def test_720p_is_a_pure_gather():
    """Scale exactly 5: all fractional weights vanish, output == src[5y+2, 5x+2]."""
    f = _frame(1280, 720)
    assert np.array_equal(opre.resize_bilinear_u8(f, 256, 144), f[2::5, 2::5])


def test_target_size_rule():
    assert opre.target_size(1280, 720) == (256, 144)
    assert opre.target_size(1920, 1080) == (256, 144)
    assert opre.target_size(854, 480) == (256, 143)
    assert opre.target_size(1920, 800) == (256, 106)


def test_videodataset_outputs(golden_dir):
    z = np.load(os.path.join(golden_dir, "preprocess_video.npz"))
    decoded, want = z["decoded_bgr"], z["dataset_rgb_chw_u8"]
    got = opre.preprocess_batch(decoded, resize=256)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want.astype(np.float32) / np.float32(255))


def test_channel_flip_and_division():
    f = np.zeros((144, 256, 3), np.uint8)
    f[..., 0] = 255          # blue in BGR
    t = opre.preprocess_frame(f, resize=None)
    assert t.shape == (3, 144, 256)
    assert t[0].max() == 0 and t[1].max() == 0 and t[2].min() == 1.0
    g = np.full((2, 2, 3), 37, np.uint8)
    assert opre.to_model_input(g)[0, 0, 0] == np.float32(37) / np.float32(255)
