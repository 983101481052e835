"""Seeded inputs shared by tests/golden/make_golden.py (which records what the REFERENCE
produces for them) and by the tests (which replay them through the oracle and the CUDA path).
Keeping the generators here means only outputs need to be committed as fixtures."""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- frames
def stripes(h: int, w: int, period: int, horizontal: bool) -> np.ndarray:
    f = np.zeros((h, w, 3), dtype=np.uint8)
    if horizontal:
        f[(np.arange(h) // period) % 2 == 1] = 255
    else:
        f[:, (np.arange(w) // period) % 2 == 1] = 255
    return f


def kat_frames() -> dict:
    """Named uint8 BGR HWC frames (SURVEY.md section 8c known-answer inputs)."""
    out = {
        "black_720": np.zeros((720, 1280, 3), np.uint8),
        "white_720": np.full((720, 1280, 3), 255, np.uint8),
        "hstripes_720": stripes(720, 1280, 40, True),
        "vstripes_720": stripes(720, 1280, 40, False),
        "hstripes_1080": stripes(1080, 1920, 60, True),
        "vstripes_1080": stripes(1080, 1920, 60, False),
        "noise_720": np.random.default_rng(0).integers(0, 256, (720, 1280, 3), dtype=np.uint8),
        "noise_1080": np.random.default_rng(0).integers(0, 256, (1080, 1920, 3), dtype=np.uint8),
    }
    blue = np.zeros((720, 1280, 3), np.uint8)
    blue[..., 0] = 255
    out["blue_720"] = blue
    return out


def smooth_images(n: int, seed: int = 7, h: int = 144, w: int = 256) -> np.ndarray:
    """[n, 3, h, w] float32 in [0, 1]: low-frequency fields plus a little noise -- the kind of
    image the classifier sees, with logits spread over all three classes."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    imgs = np.empty((n, 3, h, w), dtype=np.float32)
    for i in range(n):
        base = np.zeros((h, w))
        for _ in range(4):
            fx, fy = rng.uniform(0.5, 14, 2)
            ph = rng.uniform(0, 2 * np.pi)
            base += rng.uniform(0.1, 0.5) * np.sin(2 * np.pi * (fx * xx + fy * yy) + ph)
        gain = rng.uniform(0.05, 1.0)
        for c in range(3):
            img = 0.5 * rng.uniform(0.0, 1.6) + gain * base * rng.uniform(0.3, 1.0)
            img += rng.normal(0, 0.02, (h, w))
            if i % 4 == 3:            # dark frames: the 'blank' class
                img *= 0.06
            imgs[i, c] = np.clip(img, 0, 1)
    return imgs


# ----------------------------------------------------------------------------- scores
def scores_from_runs(runs, seed: int, spread: float = 3.0) -> np.ndarray:
    """[(label, length), ...] -> [N, 3] float32 scores whose argmax follows the runs, with a
    continuous random confidence per frame (no two run means tie)."""
    rng = np.random.default_rng(seed)
    n = sum(l for _, l in runs)
    s = rng.normal(0.0, 0.3, (n, 3)).astype(np.float32)
    pos = 0
    for lab, length in runs:
        s[pos:pos + length, lab] += (2.0 + rng.uniform(0, spread) + rng.uniform(0, 1.0, length)).astype(np.float32)
        pos += length
    return s


def random_runs(seed: int, n_runs: int):
    rng = np.random.default_rng(seed)
    runs = []
    prev = -1
    for _ in range(n_runs):
        lab = int(rng.integers(0, 3))
        if lab == prev:
            lab = (lab + 1) % 3
        prev = lab
        kind = rng.uniform()
        if kind < 0.35:
            length = int(rng.integers(1, 10))
        elif kind < 0.7:
            length = int(rng.integers(10, 100))
        else:
            length = int(rng.integers(100, 600))
        runs.append((lab, length))
    return runs


SURVEY_RUNS = [(0, 300), (2, 20), (1, 250), (2, 5), (0, 40), (1, 200), (2, 12), (0, 173)]


def segmentation_cases() -> dict:
    """name -> (scores, real_threshold, blank_threshold)."""
    cases = {}
    cases["survey_kat"] = (scores_from_runs(SURVEY_RUNS, 11), 100, 10)
    for i, n in enumerate([6, 17, 40, 90, 200]):
        cases[f"random_{n}"] = (scores_from_runs(random_runs(100 + i, n), 200 + i), 100, 10)
    cases["random_thresholds"] = (scores_from_runs(random_runs(300, 60), 301), 40, 25)
    cases["first_is_orphan"] = (scores_from_runs([(1, 4), (0, 300), (2, 50), (1, 150)], 5), 100, 10)
    cases["last_is_orphan"] = (scores_from_runs([(0, 300), (1, 150), (2, 3)], 6), 100, 10)
    cases["equal_neighbours"] = (scores_from_runs([(0, 200), (1, 30), (2, 200), (0, 120)], 7), 100, 10)
    cases["no_orphans"] = (scores_from_runs([(0, 300), (1, 150), (2, 30), (0, 101)], 8), 100, 10)
    cases["single_long_run"] = (scores_from_runs([(1, 400)], 9), 100, 10)
    cases["alternating_short"] = (scores_from_runs([(i % 3, 7 + (i * 5) % 11) for i in range(50)] + [(0, 500)], 10), 100, 10)
    # pure noise: many 1-2 frame runs (the worst case for the sequential pass)
    rng = np.random.default_rng(12)
    cases["noise_600"] = (rng.normal(0, 1, (600, 3)).astype(np.float32), 100, 10)
    return cases
