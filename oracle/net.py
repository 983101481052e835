"""Oracle: the frame classifier forward pass.  TEST INFRASTRUCTURE ONLY.

Restates ``FrameConvNet`` followed by ``FrameLinearNet`` in eval mode
(reference frameID/net.py:11-189) as plain functions over a weight dictionary
whose keys are the reference's state_dict keys (frameID/net.py:193-217 loads
two state_dicts; here the conv one is prefixed ``conv.`` and the linear one
``linear.``):

    conv.conv_layers.{i}.conv.{weight,bias}
    conv.conv_layers.{i}.bn.{weight,bias,running_mean,running_var}
    linear.layers.{j}.linear.{weight,bias}
    linear.layers.{j}.bn.{weight,bias,running_mean,running_var}   (all but the last j)

Per conv layer (frameID/net.py:33-40): conv3x3(pad 1) -> ReLU -> MaxPool(k=3, s=3,
floor) -> BatchNorm (running statistics, eps 1e-5) -- BN comes AFTER the pool.
Then AdaptiveAvgPool2d(avg_pool_size), flatten (c, i, j), and FC layers of
Linear -> ReLU -> BatchNorm1d, the last one Linear only (frameID/net.py:164-178).
The output is raw logits (no softmax anywhere in the reference).

Two restatements:
  * ``forward_f32``  -- torch CPU float32 functional ops, i.e. the same aten
    kernels the reference's modules dispatch to; bit-identical to the reference
    module (checked in tests/test_oracle_net.py when /root/reference is present,
    and against recorded logits in tests/golden/net_kat.npz everywhere).
  * ``forward_f64``  -- independent numpy float64 arithmetic (explicit window
    gathers + einsum), used to put a number on fp32/bf16 rounding.
"""
from __future__ import annotations

import json
import math

import numpy as np

BN_EPS = 1e-5


# ----------------------------------------------------------------------------- weights
def n_conv_layers(weights: dict) -> int:
    n = 0
    while f"conv.conv_layers.{n}.conv.weight" in weights:
        n += 1
    return n


def n_fc_layers(weights: dict) -> int:
    n = 0
    while f"linear.layers.{n}.linear.weight" in weights:
        n += 1
    return n


def load_weights_npz(path: str) -> tuple[dict, dict]:
    """Read the re-encoded prod_net fixture: (weights dict of float32 arrays, params dict)."""
    with np.load(path, allow_pickle=False) as z:
        weights = {k: z[k] for k in z.files if k != "__params_json__"}
        params = json.loads(bytes(z["__params_json__"]).decode("utf-8"))
    return weights, params


def random_weights(seed: int, hidden_channels: int, conv_layers: int, avg_pool_size: int,
                   linear_layers: int, linear_size: int, output_size: int,
                   input_channels: int = 3) -> dict:
    """Random-init weights of a given architecture (for the non-prod configurations)."""
    rng = np.random.default_rng(seed)
    w = {}

    def bn(prefix, c):
        w[prefix + ".weight"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
        w[prefix + ".bias"] = rng.normal(0, 0.2, c).astype(np.float32)
        w[prefix + ".running_mean"] = rng.normal(0, 0.3, c).astype(np.float32)
        w[prefix + ".running_var"] = rng.uniform(0.5, 2.0, c).astype(np.float32)

    cin = input_channels
    for i in range(conv_layers):
        bound = 1.0 / math.sqrt(cin * 9)
        w[f"conv.conv_layers.{i}.conv.weight"] = rng.uniform(
            -bound, bound, (hidden_channels, cin, 3, 3)).astype(np.float32)
        w[f"conv.conv_layers.{i}.conv.bias"] = rng.uniform(
            -bound, bound, hidden_channels).astype(np.float32)
        bn(f"conv.conv_layers.{i}.bn", hidden_channels)
        cin = hidden_channels
    sizes_in = [hidden_channels * avg_pool_size ** 2] + [linear_size] * (linear_layers - 1)
    sizes_out = [linear_size] * (linear_layers - 1) + [output_size]
    for j, (fi, fo) in enumerate(zip(sizes_in, sizes_out)):
        bound = 1.0 / math.sqrt(fi)
        w[f"linear.layers.{j}.linear.weight"] = rng.uniform(-bound, bound, (fo, fi)).astype(np.float32)
        w[f"linear.layers.{j}.linear.bias"] = rng.uniform(-bound, bound, fo).astype(np.float32)
        if j < linear_layers - 1:
            bn(f"linear.layers.{j}.bn", fo)
    return w


# ----------------------------------------------------------------------------- fp32 (torch CPU)
def forward_f32(weights: dict, x: np.ndarray, avg_pool_size: int, return_features: bool = False):
    """x: [B, 3, H, W] float32 -> logits [B, out] float32, with torch CPU fp32 ops."""
    import torch
    import torch.nn.functional as F

    t = lambda k: torch.from_numpy(np.ascontiguousarray(weights[k], dtype=np.float32))
    feats = []
    with torch.no_grad():
        y = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        for i in range(n_conv_layers(weights)):
            p = f"conv.conv_layers.{i}"
            y = F.conv2d(y, t(p + ".conv.weight"), t(p + ".conv.bias"), stride=1, padding=1)
            y = F.relu(y)
            y = F.max_pool2d(y, kernel_size=3)
            y = F.batch_norm(y, t(p + ".bn.running_mean"), t(p + ".bn.running_var"),
                             t(p + ".bn.weight"), t(p + ".bn.bias"), training=False, eps=BN_EPS)
            feats.append(y.numpy().copy())
        y = F.adaptive_avg_pool2d(y, avg_pool_size)
        y = torch.reshape(y, [y.shape[0], -1])
        nfc = n_fc_layers(weights)
        for j in range(nfc):
            p = f"linear.layers.{j}"
            y = F.linear(y, t(p + ".linear.weight"), t(p + ".linear.bias"))
            if j < nfc - 1:
                y = F.relu(y)
                y = F.batch_norm(y, t(p + ".bn.running_mean"), t(p + ".bn.running_var"),
                                 t(p + ".bn.weight"), t(p + ".bn.bias"), training=False, eps=BN_EPS)
        out = y.numpy().copy()
    return (out, feats) if return_features else out


# ----------------------------------------------------------------------------- fp64 (numpy)
def _conv3x3_f64(x, w, b):
    B, C, H, W = x.shape
    xp = np.zeros((B, C, H + 2, W + 2), dtype=np.float64)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((B, w.shape[0], H, W), dtype=np.float64)
    for ky in range(3):
        for kx in range(3):
            out += np.einsum("bchw,oc->bohw", xp[:, :, ky:ky + H, kx:kx + W], w[:, :, ky, kx],
                             optimize=True)
    return out + b[None, :, None, None]


def _maxpool3_f64(x):
    B, C, H, W = x.shape
    ph, pw = H // 3, W // 3
    v = x[:, :, :ph * 3, :pw * 3].reshape(B, C, ph, 3, pw, 3)
    return v.max(axis=(3, 5))


def _bn_f64(x, g, b, m, v, axis_shape):
    s = g / np.sqrt(v + BN_EPS)
    return (x - m.reshape(axis_shape)) * s.reshape(axis_shape) + b.reshape(axis_shape)


def adaptive_windows(n_in: int, n_out: int):
    """[(start, end)] of torch's adaptive average pool: floor(i*in/out), ceil((i+1)*in/out)."""
    return [((i * n_in) // n_out, -((-(i + 1) * n_in) // n_out)) for i in range(n_out)]


def _adaptive_avg_f64(x, size):
    B, C, H, W = x.shape
    out = np.zeros((B, C, size, size), dtype=np.float64)
    for i, (r0, r1) in enumerate(adaptive_windows(H, size)):
        for j, (c0, c1) in enumerate(adaptive_windows(W, size)):
            out[:, :, i, j] = x[:, :, r0:r1, c0:c1].mean(axis=(2, 3))
    return out


def forward_f64(weights: dict, x: np.ndarray, avg_pool_size: int, return_features: bool = False):
    g = lambda k: np.asarray(weights[k], dtype=np.float64)
    y = np.asarray(x, dtype=np.float64)
    feats = []
    for i in range(n_conv_layers(weights)):
        p = f"conv.conv_layers.{i}"
        y = _conv3x3_f64(y, g(p + ".conv.weight"), g(p + ".conv.bias"))
        y = np.maximum(y, 0.0)
        y = _maxpool3_f64(y)
        y = _bn_f64(y, g(p + ".bn.weight"), g(p + ".bn.bias"), g(p + ".bn.running_mean"),
                    g(p + ".bn.running_var"), (1, -1, 1, 1))
        feats.append(y.copy())
    y = _adaptive_avg_f64(y, avg_pool_size).reshape(y.shape[0], -1)
    nfc = n_fc_layers(weights)
    for j in range(nfc):
        p = f"linear.layers.{j}"
        y = y @ g(p + ".linear.weight").T + g(p + ".linear.bias")
        if j < nfc - 1:
            y = np.maximum(y, 0.0)
            y = _bn_f64(y, g(p + ".bn.weight"), g(p + ".bn.bias"), g(p + ".bn.running_mean"),
                        g(p + ".bn.running_var"), (1, -1))
    return (y, feats) if return_features else y


# ----------------------------------------------------------------------------- 16-bit operand emulation
def _bf16_round(a: np.ndarray) -> np.ndarray:
    """Round float32 values to the nearest bfloat16 (ties to even), returned as float32."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    r = ((u.astype(np.uint64) + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32)


def _f16_round(a: np.ndarray) -> np.ndarray:
    return np.asarray(a, dtype=np.float32).astype(np.float16).astype(np.float32)


def _conv1_acc16(y, w_folded: np.ndarray, bias_folded: np.ndarray, sign: np.ndarray, shift: np.ndarray):
    """Layer 1 as conv1_fused_tc_kernel<.., ACC16> computes it (cut-detection_b200/csrc/conv_tc.cu, tc_prepare): pixels v = 256 y
    enter as v * 2^-24, the taps of channel co as fp16(w * 2^a), a such that the largest lands in [2^14, 2^15); the fp16
    accumulator takes the three kernel rows one after the other (one rounding each), the bias rides on a constant 1.0 in three
    fp16 pieces; 9-way max and ReLU on the fp16 values; one fused multiply-add by +-2^(16-a) and the fp16 shift."""
    import torch
    import torch.nn.functional as F

    C = w_folded.shape[0]
    wmax = np.abs(w_folded.reshape(C, -1)).max(axis=1)
    e = np.where(wmax > 0, np.frexp(wmax)[1], 0).astype(np.int64)
    ae = 15 - e
    up = np.ldexp(np.float32(1), ae).astype(np.float32)
    taps = _f16_round(w_folded * up.reshape(-1, 1, 1, 1)).astype(np.float64)
    pieces = np.zeros((3, C))
    rest = bias_folded.astype(np.float64) * np.ldexp(1.0, ae - 16)
    for ky in range(3):
        pieces[ky] = _f16_round(rest.astype(np.float32)).astype(np.float64)
        rest = rest - pieces[ky]
    v = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)) * 256.0          # integers 0..255 for decoded frames
    vp = F.pad(v, (1, 1, 1, 1))
    H = v.shape[2]
    acc = None
    for ky in range(3):
        part = F.conv2d(vp[:, :, ky:ky + H, :], torch.from_numpy(taps[:, :, ky:ky + 1, :].copy()), None) * 2.0 ** -24
        part = part + torch.from_numpy(pieces[ky]).view(1, -1, 1, 1)
        tot = part if acc is None else acc.double() + part
        acc = tot.half()                                                            # round to nearest even, subnormals kept
    m = torch.relu(F.max_pool2d(acc.double(), kernel_size=3))
    sc = torch.from_numpy(sign.astype(np.float64) * np.ldexp(1.0, 16 - ae)).view(1, -1, 1, 1)
    sh = torch.from_numpy(_f16_round(shift).astype(np.float64)).view(1, -1, 1, 1)
    return (m * sc + sh).half().float()


def forward_tc_emulated(weights: dict, x: np.ndarray, avg_pool_size: int, fmt: str = "f16", return_features: bool = False,
                        conv1_acc16: bool = False):
    """What the tensor-core path computes, restated on the CPU: conv operands rounded to 16 bits (fmt 'f16' as
    cut-detection_b200/csrc/conv_tc.cu ships, or 'bf16'), layer-1 input stored as x*255/256 with 256/255 folded into its
    weights (bf16: x*255 and 1/255), exact products, float32 epilogue (max-pool of the raw sums, +bias, ReLU, folded
    BatchNorm affine; layer 1 carries |scale| in its taps), 16-bit inter-layer activations, float32 head.  It separates 'the kernel is wrong' from
    '16-bit operands round': the CUDA path must match THIS to ~1e-3, and this differs from the fp32 reference by the
    rounding the format implies.  conv1_acc16: layer 1 as the fused frames kernel computes it by default (fp16 accumulators,
    see _conv1_acc16)."""
    import torch
    import torch.nn.functional as F

    rnd = _f16_round if fmt == "f16" else _bf16_round
    in_scale, w_scale = (np.float32(255.0 / 256.0), np.float32(256.0 / 255.0)) if fmt == "f16" else \
        (np.float32(255.0), np.float32(1.0 / 255.0))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    n = n_conv_layers(weights)
    feats = []
    with torch.no_grad():
        y = t(rnd(np.asarray(x, np.float32) * in_scale))
        for i in range(n):
            p = f"conv.conv_layers.{i}"
            w = np.asarray(weights[p + ".conv.weight"], np.float32)
            if i == 0:
                w = w * w_scale
            g = np.asarray(weights[p + ".bn.weight"], np.float64)
            s = (g / np.sqrt(np.asarray(weights[p + ".bn.running_var"], np.float64) + BN_EPS)).astype(np.float32)
            sh = (np.asarray(weights[p + ".bn.bias"], np.float64)
                  - np.asarray(weights[p + ".bn.running_mean"], np.float64) * s.astype(np.float64)).astype(np.float32)
            bias = np.asarray(weights[p + ".conv.bias"], np.float32)
            if i == 0 and fmt == "f16":
                # conv_tc.cu folds |BatchNorm scale| into the layer-1 taps BEFORE rounding them (s * relu(u) = sign(s) * relu(|s| u)),
                # which frees the fused kernel's epilogue of the bias add and the ReLU
                w = w * np.abs(s).reshape(-1, 1, 1, 1)
                bias = bias * np.abs(s)
                s = np.where(s < 0, np.float32(-1), np.float32(1)).astype(np.float32)
            if i == 0 and conv1_acc16 and fmt == "f16":
                z = _conv1_acc16(y.numpy(), w, bias, s, sh)
            else:
                z = F.conv2d(y.double(), t(rnd(w)).double(), None, stride=1, padding=1).float()
                z = F.max_pool2d(z, kernel_size=3)
                z = torch.relu(z + t(bias).view(1, -1, 1, 1))
                z = z * t(s).view(1, -1, 1, 1) + t(sh).view(1, -1, 1, 1)
            if i < n - 1:
                z = t(rnd(z.numpy()))
            feats.append(z.numpy().copy())
            y = z
        y = F.adaptive_avg_pool2d(y, avg_pool_size)
        y = torch.reshape(y, [y.shape[0], -1])
        nfc = n_fc_layers(weights)
        for j in range(nfc):
            p = f"linear.layers.{j}"
            y = F.linear(y, t(weights[p + ".linear.weight"]), t(weights[p + ".linear.bias"]))
            if j < nfc - 1:
                y = F.relu(y)
                y = F.batch_norm(y, t(weights[p + ".bn.running_mean"]), t(weights[p + ".bn.running_var"]),
                                 t(weights[p + ".bn.weight"]), t(weights[p + ".bn.bias"]), training=False, eps=BN_EPS)
        out = y.numpy().copy()
    return (out, feats) if return_features else out
