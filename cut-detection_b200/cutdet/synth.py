"""Synthetic pre-decoded frames for tests and benchmarks (video decode is outside the hot path).

A clip is a seeded plan of runs over the three classes of the reference (frameID/data.py:116):
    a22 (0): horizontal stripes      ez (1): vertical stripes      b (2): black
drawn in the look the shipped classifier separates with a wide margin, plus a small block of
per-frame pseudo-random pixels so no two frames -- and no two run means -- are identical.  Every pixel is a pure
integer function of (seed, frame index, y, x, channel), evaluated with the same int64 arithmetic by numpy on the
host and by torch on the device, so the CPU oracle and the CUDA path see identical bytes without the frames ever
having to be copied (a full game at 720p is 896 GB).
"""
from __future__ import annotations

import numpy as np

A22, EZ, BLANK = 0, 1, 2


def plan_runs(n_frames: int, seed: int):
    """[(label, length)] covering exactly n_frames: long runs (200-5000), sub-threshold real runs (<100) and
    sub-threshold blank runs (<10) mixed, never two equal labels in a row."""
    rng = np.random.default_rng(seed)
    runs, total, prev = [], 0, -1
    while total < n_frames:
        kind = rng.uniform()
        if kind < 0.55:
            lab, length = int(rng.integers(0, 2)), int(rng.integers(200, 5001))
        elif kind < 0.70:
            lab, length = BLANK, int(rng.integers(10, 120))
        elif kind < 0.85:
            lab, length = int(rng.integers(0, 2)), int(rng.integers(3, 100))
        else:
            lab, length = BLANK, int(rng.integers(1, 10))
        if lab == prev:
            lab = (lab + 1) % 3
        length = min(length, n_frames - total)
        runs.append((lab, length))
        total += length
        prev = lab
    return runs


def _mix(v):
    v = (v * 1103515245 + 12345) & 0x7FFFFFFF
    v = v ^ (v >> 13)
    v = (v * 1103515245 + 12345) & 0x7FFFFFFF
    return (v >> 16) & 0xFF


class SyntheticClip:
    def __init__(self, height: int = 720, width: int = 1280, n_frames: int = 1800, seed: int = 0, runs=None):
        self.height, self.width, self.n_frames, self.seed = height, width, n_frames, seed
        self.runs = runs if runs is not None else plan_runs(n_frames, seed)
        assert sum(l for _, l in self.runs) == n_frames
        self.labels = np.concatenate([np.full(l, lab, dtype=np.uint8) for lab, l in self.runs])
        self.period_y = max(height // 18, 2)      # 40 rows at 720p, 60 at 1080p
        self.period_x = max(width // 32, 2)       # 40 columns at 1280, 60 at 1920
        self.patch_h, self.patch_w = max(height // 6, 1), max(width // 8, 1)

    # ------------------------------------------------------------------ shared integer recipe
    def _patch(self, f, lab, arange):
        """Noise block values (int64 in [0, 255]) and its top-left corner for frames f (int64 [n])."""
        n = f.shape[0]
        f4 = f.reshape(n, 1, 1, 1)
        py = (f * 37 + self.seed * 11) % (self.height - self.patch_h + 1)
        px = (f * 101 + self.seed * 7) % (self.width - self.patch_w + 1)
        yy = py.reshape(n, 1, 1, 1) + arange(self.patch_h).reshape(1, self.patch_h, 1, 1)
        xx = px.reshape(n, 1, 1, 1) + arange(self.patch_w).reshape(1, 1, self.patch_w, 1)
        cc = arange(3).reshape(1, 1, 1, 3)
        noise = _mix((f4 * 1000003 + yy * 10007 + xx * 101 + cc * 7 + self.seed * 7919) & 0x7FFFFFFF)
        l4 = lab.reshape(n, 1, 1, 1)
        hs = ((yy // self.period_y) % 2)              # 1 where the a22 pattern is white
        vs = ((xx // self.period_x) % 2)
        small = noise >> 3                             # [0, 32)
        a22 = hs * (255 - small) + (1 - hs) * small + 0 * xx
        ez = vs * (255 - small) + (1 - vs) * small + 0 * yy
        blank = noise >> 5                             # [0, 8): stays dark
        is_a22, is_ez = (l4 == A22) * 1, (l4 == EZ) * 1
        vals = is_a22 * a22 + is_ez * ez + (1 - is_a22 - is_ez) * blank
        return vals, py, px

    def frames_numpy(self, start: int, count: int) -> np.ndarray:
        """uint8 BGR HWC [count, h, w, 3] on the host."""
        h, w = self.height, self.width
        lab = self.labels[start:start + count].astype(np.int64)
        f = np.arange(start, start + count, dtype=np.int64)
        hs = (((np.arange(h) // self.period_y) % 2) * 255).astype(np.uint8).reshape(h, 1, 1)
        vs = (((np.arange(w) // self.period_x) % 2) * 255).astype(np.uint8).reshape(1, w, 1)
        looks = np.stack([np.broadcast_to(hs, (h, w, 3)), np.broadcast_to(vs, (h, w, 3)), np.zeros((h, w, 3), np.uint8)])
        out = looks[lab]
        for s in range(0, count, 64):
            e = min(count, s + 64)
            vals, py, px = self._patch(f[s:e], lab[s:e], lambda n: np.arange(n, dtype=np.int64))
            for i in range(e - s):
                out[s + i, py[i]:py[i] + self.patch_h, px[i]:px[i] + self.patch_w] = vals[i].astype(np.uint8)
        return out

    def frames_torch(self, start: int, count: int, device="cuda", rows=None):
        """Same bytes, generated on ``device``.  ``rows`` (array of source rows) keeps only those rows
        (row-compacted frames)."""
        import torch

        h, w = self.height, self.width
        ar = lambda n: torch.arange(n, dtype=torch.int64, device=device)
        hs = (((ar(h) // self.period_y) % 2) * 255).to(torch.uint8).reshape(h, 1, 1).expand(h, w, 3)
        vs = (((ar(w) // self.period_x) % 2) * 255).to(torch.uint8).reshape(1, w, 1).expand(h, w, 3)
        looks = torch.stack([hs, vs, torch.zeros((h, w, 3), dtype=torch.uint8, device=device)])
        out_rows = h if rows is None else len(rows)
        out = torch.empty((count, out_rows, w, 3), dtype=torch.uint8, device=device)
        row_idx = None if rows is None else torch.as_tensor(np.asarray(rows), dtype=torch.int64, device=device)
        labels = torch.from_numpy(self.labels[start:start + count].astype(np.int64)).to(device)
        for s in range(0, count, 64):
            e = min(count, s + 64)
            n = e - s
            fr = looks[labels[s:e]]
            f = torch.arange(start + s, start + e, dtype=torch.int64, device=device)
            vals, py, px = self._patch(f, labels[s:e], ar)
            ii = ar(n).reshape(n, 1, 1)
            yy = (py.reshape(n, 1) + ar(self.patch_h).reshape(1, -1)).reshape(n, -1, 1)
            xx = (px.reshape(n, 1) + ar(self.patch_w).reshape(1, -1)).reshape(n, 1, -1)
            fr[ii, yy, xx] = vals.to(torch.uint8)
            out[s:e] = fr if row_idx is None else fr[:, row_idx]
        return out
