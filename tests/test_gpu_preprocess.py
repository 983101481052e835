"""K1 on the GPU vs the oracle: bit-exact for every geometry the reference's resize rule can produce."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess as opre

pytestmark = pytest.mark.gpu

GEOMETRIES = [(1280, 720), (1920, 1080), (3840, 2160), (640, 360), (854, 480), (1920, 800), (512, 288), (256, 144),
              (320, 180), (1000, 562), (130, 70)]


def _frames(w, h, n, seed=0):
    return np.random.default_rng(seed + w * 31 + h).integers(0, 256, (n, h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("w,h", GEOMETRIES)
def test_resize_u8_bit_exact(w, h):
    from cutdet import engine
    f = _frames(w, h, 2)
    plan = engine.ResizePlan.for_video(h, w, 256)
    got = engine.preprocess_u8(plan, torch.from_numpy(f).cuda()).cpu().numpy()
    nw, nh = opre.target_size(w, h)
    want = np.stack([opre.resize_bilinear_u8(x, nw, nh) for x in f])
    assert got.shape == want.shape
    assert np.array_equal(got, want)


@pytest.mark.parametrize("w,h", [(1280, 720), (1920, 1080), (854, 480), (512, 288)])
def test_model_input_f32_bit_exact(w, h):
    from cutdet import engine
    f = _frames(w, h, 3, seed=5)
    plan = engine.ResizePlan.for_video(h, w, 256)
    got = engine.preprocess_f32(plan, torch.from_numpy(f).cuda()).cpu().numpy()
    want = opre.preprocess_batch(f, 256)
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want)          # includes the true /255 division and the BGR->RGB flip


@pytest.mark.parametrize("w,h", [(1280, 720), (1920, 1080), (854, 480)])
def test_row_compacted_frames(w, h):
    """Only the source rows the resize reads are uploaded (144 of 720 at 720p)."""
    from cutdet import engine
    f = _frames(w, h, 2, seed=9)
    plan = engine.ResizePlan.for_video(h, w, 256)
    if (w, h) == (1280, 720):
        assert np.array_equal(plan.rows, np.arange(2, 720, 5))
    if (w, h) == (1920, 1080):
        assert len(plan.rows) == 288
    compact = torch.from_numpy(np.ascontiguousarray(f[:, plan.rows])).cuda()
    got = engine.preprocess_f32(plan, compact, compact=True).cpu().numpy()
    assert np.array_equal(got, opre.preprocess_batch(f, 256))


def test_strided_and_empty_batches():
    from cutdet import engine
    f = _frames(640, 360, 5, seed=2)
    plan = engine.ResizePlan.for_video(360, 640, 256)
    dev = torch.from_numpy(f).cuda()
    got = engine.preprocess_f32(plan, dev[::2]).cpu().numpy()       # frame stride = 2 frames
    assert np.array_equal(got, opre.preprocess_batch(f[::2], 256))
    empty = engine.preprocess_f32(plan, dev[:0])
    assert tuple(empty.shape) == (0, 3, 144, 256)
    with pytest.raises(ValueError):
        engine.preprocess_f32(plan, torch.zeros((1, 10, 10, 3), dtype=torch.uint8, device="cuda"))
    with pytest.raises(RuntimeError):
        engine.preprocess_f32(plan, torch.from_numpy(f))              # CPU tensor: no fallback


def test_videodataset_golden(golden_dir):
    """Frames decoded by the reference's VideoDataset run (recorded) -> K1 == what the reference yielded."""
    from cutdet import engine
    z = np.load(os.path.join(golden_dir, "preprocess_video.npz"))
    decoded, want = z["decoded_bgr"], z["dataset_rgb_chw_u8"]
    plan = engine.ResizePlan.for_video(decoded.shape[1], decoded.shape[2], 256)
    got = engine.preprocess_f32(plan, torch.from_numpy(decoded).cuda()).cpu().numpy()
    assert np.array_equal(got, want.astype(np.float32) / np.float32(255))


def test_no_resize_is_identity_layout():
    from cutdet import engine
    f = _frames(256, 144, 2, seed=4)
    plan = engine.ResizePlan(144, 256, 144, 256)
    got = engine.preprocess_f32(plan, torch.from_numpy(f).cuda()).cpu().numpy()
    assert np.array_equal(got, opre.preprocess_batch(f, None))


@pytest.mark.parametrize("kernel", [1, 2, 3])
def test_every_k1_kernel_bit_exact(kernel):
    """The library picks one of three K1 kernels per geometry and output type (one thread per pixel, staged rows, four
    adjacent pixels per thread for two-tap resizes); cutdet_debug_k1_kernel forces each of them in turn.  Every one must give
    the oracle's bytes for both outputs, from whole and from row-compacted frames, also where a frame's last block of output
    rows is ragged (70 and 562 source rows) and where the selected kernel does not apply and the library falls back."""
    from cutdet import _cabi, engine
    lib = _cabi.lib()
    _cabi.check(lib.cutdet_debug_k1_kernel(kernel))
    try:
        for w, h in GEOMETRIES:
            if w * h > 1920 * 1080:
                continue
            f = _frames(w, h, 3, seed=11 + kernel)
            plan = engine.ResizePlan.for_video(h, w, 256)
            dev = torch.from_numpy(f).cuda()
            nw, nh = opre.target_size(w, h)
            want_u8 = np.stack([opre.resize_bilinear_u8(x, nw, nh) for x in f])
            want_f32 = opre.preprocess_batch(f, 256)
            assert np.array_equal(engine.preprocess_u8(plan, dev).cpu().numpy(), want_u8), (kernel, w, h, "u8")
            assert np.array_equal(engine.preprocess_f32(plan, dev).cpu().numpy(), want_f32), (kernel, w, h, "f32")
            compact = torch.from_numpy(np.ascontiguousarray(f[:, plan.rows])).cuda()
            assert np.array_equal(engine.preprocess_u8(plan, compact, compact=True).cpu().numpy(), want_u8), (kernel, w, h, "compact u8")
            assert np.array_equal(engine.preprocess_f32(plan, compact, compact=True).cpu().numpy(), want_f32), (kernel, w, h, "compact f32")
            # a view that starts one frame in: frame pointers that are 16-byte aligned only when a frame's size is
            assert np.array_equal(engine.preprocess_u8(plan, dev[1:]).cpu().numpy(), want_u8[1:]), (kernel, w, h, "offset")
    finally:
        _cabi.check(lib.cutdet_debug_k1_kernel(0))
