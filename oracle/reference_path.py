"""Oracle: the reference's CPU path end to end, with the reference's own library calls, for TIMING and for
whole-clip parity.  TEST/BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline and --impl reference legs, tests).

/root/reference does not exist on the GPU box, so this module restates the three stages of
``segment_video.py`` (reference segment_video.py:28-77) with the same third-party calls the reference makes:

  1. preprocessing   cv2.resize(INTER_LINEAR) + torch.tensor/permute/flip//255 per frame, torch.stack per batch
                     (frameID/data.py:218-228, DataLoader default_collate)
  2. inference       a torch.jit.trace'd nn.Sequential of Conv2d/ReLU/MaxPool2d/BatchNorm2d ... Linear, eval mode,
                     under torch.no_grad(), float32 on the CPU with all host threads
                     (frameID/net.py:11-189; training_scripts/make_torchscript_model.py:25-27 for the trace)
  3. segmentation    run table, glue_orphans, combine_adjacent_segments, CSV (oracle.segmentation)

``tests/test_oracle_net.py`` checks (in the build container) that the module built here is bit-identical to the
reference's ``load_default_net()``.
"""
from __future__ import annotations

import time

import numpy as np

from . import net as onet
from . import segmentation as oseg


def build_torch_net(weights: dict, avg_pool_size: int):
    """nn.Sequential(trunk, head) with the oracle weights loaded, eval mode."""
    import torch
    import torch.nn as nn

    class Flatten2(nn.Module):
        def forward(self, x):
            return torch.reshape(x, [x.shape[0], -1])

    def load_bn(bn, prefix):
        bn.weight.data = torch.from_numpy(weights[prefix + ".weight"].copy())
        bn.bias.data = torch.from_numpy(weights[prefix + ".bias"].copy())
        bn.running_mean.data = torch.from_numpy(weights[prefix + ".running_mean"].copy())
        bn.running_var.data = torch.from_numpy(weights[prefix + ".running_var"].copy())

    trunk = []
    for i in range(onet.n_conv_layers(weights)):
        p = f"conv.conv_layers.{i}"
        w = weights[p + ".conv.weight"]
        conv = nn.Conv2d(w.shape[1], w.shape[0], kernel_size=3, padding=1)
        conv.weight.data = torch.from_numpy(w.copy())
        conv.bias.data = torch.from_numpy(weights[p + ".conv.bias"].copy())
        bn = nn.BatchNorm2d(w.shape[0])
        load_bn(bn, p + ".bn")
        trunk += [conv, nn.ReLU(), nn.MaxPool2d(kernel_size=3), bn]
    trunk += [nn.AdaptiveAvgPool2d(avg_pool_size), Flatten2()]
    head = []
    nfc = onet.n_fc_layers(weights)
    for j in range(nfc):
        p = f"linear.layers.{j}"
        w = weights[p + ".linear.weight"]
        lin = nn.Linear(w.shape[1], w.shape[0])
        lin.weight.data = torch.from_numpy(w.copy())
        lin.bias.data = torch.from_numpy(weights[p + ".linear.bias"].copy())
        head.append(lin)
        if j < nfc - 1:
            bn = nn.BatchNorm1d(w.shape[0])
            load_bn(bn, p + ".bn")
            head += [nn.ReLU(), bn]
    net = nn.Sequential(nn.Sequential(*trunk), nn.Sequential(*head))
    net.eval()
    return net


def trace(net, height: int = 144, width: int = 256):
    import torch

    with torch.no_grad():
        return torch.jit.trace(net, torch.randn([1, 3, height, width]))


def preprocess_frame_like_reference(frame_bgr_u8: np.ndarray, new_width: int, new_height: int):
    import cv2
    import torch

    frame = cv2.resize(frame_bgr_u8, (new_width, new_height), interpolation=cv2.INTER_LINEAR)
    return torch.flip(torch.tensor(frame, dtype=torch.float).permute(2, 0, 1), (0,)) / 255


class CpuReferencePath:
    """The three stages over pre-decoded frames, timed separately with perf_counter."""

    def __init__(self, weights: dict, params: dict, threads: int | None = None):
        import os
        import torch

        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.net = trace(build_torch_net(weights, params["avg_pool_size"]))
        self.t_pre = self.t_net = self.t_seg = 0.0
        self.frames = 0

    def score_batch(self, frames_bgr_u8: np.ndarray):
        """One DataLoader batch: preprocess every frame, stack, run the traced net.  Returns logits (torch)."""
        import torch
        from .preprocess import target_size

        h, w = frames_bgr_u8.shape[1:3]
        nw, nh = target_size(w, h, 256)
        t0 = time.perf_counter()
        batch = torch.stack([preprocess_frame_like_reference(f, nw, nh) for f in frames_bgr_u8])
        t1 = time.perf_counter()
        with torch.no_grad():
            y = self.net(batch)
        t2 = time.perf_counter()
        self.t_pre += t1 - t0
        self.t_net += t2 - t1
        self.frames += frames_bgr_u8.shape[0]
        return y

    def segment(self, logits: np.ndarray, real_threshold: int = 100, blank_threshold: int = 10):
        t0 = time.perf_counter()
        out = oseg.segment(np.asarray(logits, dtype=np.float32), real_threshold, blank_threshold)
        self.t_seg += time.perf_counter() - t0
        return out

    def total_seconds(self) -> float:
        return self.t_pre + self.t_net + self.t_seg
