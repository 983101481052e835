"""TEST INFRASTRUCTURE ONLY (imported by tests/; never by the product path).

CPU restatement of the reference's contrastive forward pass:
  * ``forward_batchstats``  FrameConvNet / FrameLinearNet with every BatchNorm in training mode, as
    training_scripts/learn_contrasts.py:100-107 runs them (no ``.eval()``): CNNLayer = conv3x3(p1) -> ReLU -> MaxPool(3)
    -> BatchNorm2d(batch mean, biased batch variance) (frameID/net.py:33-40), FCLayer = Linear -> ReLU -> BatchNorm1d (net.py:62-68);
  * ``contrastive_loss``    ContrastiveLoss.forward, frameID/metrics.py:23-47, in numpy float64.
Pinned against outputs of the unmodified reference: tests/golden/contrastive_kat.npz (tests/golden/make_contrastive_golden.py).
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-5
BIG_NUMBER = 1e9


def _bn_batch(z, gamma, beta, axes):
    mean = z.mean(axis=axes, keepdims=True)
    var = z.var(axis=axes, keepdims=True)          # biased, as nn.BatchNorm normalises in training mode
    shape = [1] * z.ndim
    shape[1] = -1
    return (z - mean) / np.sqrt(var + BN_EPS) * gamma.reshape(shape) + beta.reshape(shape)


def forward_batchstats(weights: dict, x: np.ndarray, avg_pool_size: int) -> np.ndarray:
    """weights: 'conv.conv_layers.{i}.*' and/or 'linear.layers.{j}.*' (either half may be absent); x float32 NCHW or [B, F]."""
    import torch
    import torch.nn.functional as F

    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    y = np.asarray(x, np.float64)
    n_conv = len({k.split(".")[2] for k in weights if k.startswith("conv.conv_layers.")})
    for i in range(n_conv):
        p = f"conv.conv_layers.{i}"
        z = F.conv2d(t(y), t(weights[p + ".conv.weight"]), t(weights[p + ".conv.bias"]), stride=1, padding=1)
        z = F.max_pool2d(torch.relu(z), kernel_size=3).numpy()
        y = _bn_batch(z, np.asarray(weights[p + ".bn.weight"], np.float64), np.asarray(weights[p + ".bn.bias"], np.float64), (0, 2, 3))
    if n_conv:
        y = F.adaptive_avg_pool2d(t(y), avg_pool_size).numpy().reshape(y.shape[0], -1)
    n_fc = len({k.split(".")[2] for k in weights if k.startswith("linear.layers.")})
    for j in range(n_fc):
        p = f"linear.layers.{j}"
        y = y.reshape(y.shape[0], -1) @ np.asarray(weights[p + ".linear.weight"], np.float64).T + np.asarray(weights[p + ".linear.bias"], np.float64)
        if j < n_fc - 1:
            y = _bn_batch(np.maximum(y, 0.0), np.asarray(weights[p + ".bn.weight"], np.float64),
                          np.asarray(weights[p + ".bn.bias"], np.float64), (0,))
    return y.astype(np.float32)


def contrastive_loss(x: np.ndarray, temperature: float = 1.0, h_norm: bool = True):
    """-> (loss, logits_ab [B, B]); metrics.py:23-47."""
    x = np.asarray(x, np.float64)
    if h_norm:
        x = x / np.maximum(np.linalg.norm(x, axis=-1, keepdims=True), 1e-12)
    b = x.shape[0] // 2
    h1, h2 = x[:b], x[b:]
    eye = np.eye(b)
    aa = h1 @ h1.T / temperature - eye * BIG_NUMBER
    bb = h2 @ h2.T / temperature - eye * BIG_NUMBER
    ab = h1 @ h2.T / temperature
    ba = ab.T

    def ce(logits):
        m = logits.max(axis=1, keepdims=True)
        lse = m[:, 0] + np.log(np.exp(logits - m).sum(axis=1))
        return lse - logits[np.arange(b), np.arange(b)]

    loss = np.mean(ce(np.concatenate([ab, aa], axis=1)) + ce(np.concatenate([ba, bb], axis=1)))
    return np.float32(loss), ab.astype(np.float32)
