"""Oracle: frame preprocessing.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates what ``VideoDataset.__next__`` does to one decoded frame
(reference frameID/data.py:218-228):

    frame = cv2.resize(frame, (new_width, new_height), interpolation=cv2.INTER_LINEAR)
    frame = flip(tensor(frame, float).permute(2, 0, 1), (0,)) / 255

The resize arithmetic is OpenCV's (third-party, not under /root/reference;
``opencv-python`` is unpinned in the reference's requirements.txt:1, cv2 4.13.0
in this image).  Its uint8 INTER_LINEAR path is fixed point:

  * coefficients: ``f = float32((d + 0.5) * scale - 0.5)``, ``s = floor(f)``,
    ``f -= s``; weights ``round_half_even((1 - f) * 2048)``, ``round_half_even(f * 2048)``
    as int16.  Horizontally, taps that fall outside the row get ``f = 0`` and a
    clamped index; vertically the row indices are clamped and the weights kept.
  * horizontal pass (int32): ``S = a0 * p[x0] + a1 * p[x0 + 1]``
  * vertical pass: ``out = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2``
  * special case: an exact 2x downscale in both directions is rerouted by OpenCV
    to the INTER_AREA fast path = rounded 2x2 box mean ``(a + b + c + d + 2) >> 2``.
  * identical source and destination size: a copy.

Pinned by tests/test_oracle_preprocess.py against cv2.resize itself (every
pixel, many geometries) and against VideoDataset outputs recorded from the
reference (tests/golden/preprocess_video.npz).
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS  # 2048


def target_size(width: int, height: int, resize: int = 256) -> tuple[int, int]:
    """(new_width, new_height) exactly as frameID/data.py:199-202 computes them."""
    new_width = resize
    new_height = int(height * (new_width / width))
    return new_width, new_height


def linear_coeffs(src: int, dst: int, clamp_weights: bool):
    """Tap index and the two int16 weights for every destination coordinate.

    ``clamp_weights=True`` is the horizontal rule (out-of-range tap => weight 0,
    index clamped); ``False`` is the vertical rule (indices clamped later,
    weights untouched).
    Returns (i0, i1, w0, w1): int32 arrays of length ``dst``.
    """
    scale = 1.0 / (float(dst) / float(src))          # double, like cv::resize
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weights:
        low = s < 0
        f[low] = 0.0
        s[low] = 0
        high = s >= src - 1
        f[high] = 0.0
        s[high] = src - 1
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_ONE)).astype(np.int32)
    w1 = np.rint(f * np.float32(COEF_ONE)).astype(np.int32)
    i0 = np.clip(s, 0, src - 1).astype(np.int32)
    i1 = np.clip(s + 1, 0, src - 1).astype(np.int32)
    return i0, i1, w0, w1


def resize_bilinear_u8(frame: np.ndarray, new_width: int, new_height: int) -> np.ndarray:
    """Bit-exact restatement of cv2.resize(frame, (new_width, new_height), INTER_LINEAR)
    for a uint8 HWC frame."""
    frame = np.ascontiguousarray(frame)
    assert frame.dtype == np.uint8 and frame.ndim == 3
    h, w, _ = frame.shape
    if (w, h) == (new_width, new_height):
        return frame.copy()
    if w == 2 * new_width and h == 2 * new_height:
        p = frame.astype(np.int32)
        box = p[0::2, 0::2] + p[0::2, 1::2] + p[1::2, 0::2] + p[1::2, 1::2]
        return ((box + 2) >> 2).astype(np.uint8)
    x0, x1, a0, a1 = linear_coeffs(w, new_width, clamp_weights=True)
    y0, y1, b0, b1 = linear_coeffs(h, new_height, clamp_weights=False)
    p = frame.astype(np.int32)
    # horizontal pass only on the source rows the vertical pass will touch
    rows = np.unique(np.concatenate([y0, y1]))
    hpass = np.zeros((h, new_width, frame.shape[2]), dtype=np.int32)
    hpass[rows] = (p[rows][:, x0, :] * a0[None, :, None]
                   + p[rows][:, x1, :] * a1[None, :, None])
    s0 = hpass[y0] >> 4
    s1 = hpass[y1] >> 4
    out = (((b0[:, None, None] * s0) >> 16) + ((b1[:, None, None] * s1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def to_model_input(frame_bgr_u8: np.ndarray) -> np.ndarray:
    """uint8 BGR HWC -> float32 RGB CHW in [0, 1] (frameID/data.py:225-228).

    The reference divides in float32 (a true division, not a multiply by 1/255).
    """
    chw = np.transpose(frame_bgr_u8.astype(np.float32), (2, 0, 1))[::-1]
    return np.ascontiguousarray(chw / np.float32(255.0))


def preprocess_frame(frame_bgr_u8: np.ndarray, resize: int | None = 256) -> np.ndarray:
    """One decoded frame -> the tensor VideoDataset yields for it."""
    if resize is not None:
        h, w, _ = frame_bgr_u8.shape
        nw, nh = target_size(w, h, resize)
        frame_bgr_u8 = resize_bilinear_u8(frame_bgr_u8, nw, nh)
    return to_model_input(frame_bgr_u8)


def preprocess_batch(frames_bgr_u8: np.ndarray, resize: int | None = 256) -> np.ndarray:
    """[B, h, w, 3] uint8 -> [B, 3, H', W'] float32 (default_collate = stack)."""
    return np.stack([preprocess_frame(f, resize) for f in frames_bgr_u8], axis=0)
