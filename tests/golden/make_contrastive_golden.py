#!/usr/bin/env python3
"""Record the reference's contrastive forward pass (BASELINE config 5 / SURVEY 8f rank 2) for seeded inputs:
    FrameConvNet(hidden_channels=32, n_conv_layers=3) + FrameLinearNet(3, 32, 32, 8), both LEFT IN TRAINING MODE exactly as
    training_scripts/learn_contrasts.py:68-76,100-108 runs them, then ContrastiveLoss(batch_size) (frameID/metrics.py).
Imports the unmodified reference from /root/reference (build container only) and writes tests/golden/contrastive_kat.npz:
the random-init parameters (state_dict keys, 'conv.'/'linear.' prefixed), the encoder output, the projection,
the loss and logits_ab.

    python tests/golden/make_contrastive_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
import torch  # noqa: E402
import frameID.net as ref_net  # noqa: E402
from frameID.metrics import ContrastiveLoss  # noqa: E402
import kat_inputs  # noqa: E402

assert ref_net.__file__.startswith(REF), ref_net.__file__
torch.manual_seed(5)
PAIRS = 6
conv_net = ref_net.FrameConvNet(hidden_channels=32, n_conv_layers=3)
linear_net = ref_net.FrameLinearNet(n_layers=3, input_size=32, hidden_size=32, output_size=8)
with torch.no_grad():                         # non-trivial BatchNorm parameters (fresh modules have gamma = 1, beta = 0)
    for m in list(conv_net.modules()) + list(linear_net.modules()):
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm1d)):
            m.weight.uniform_(-1.5, 1.5)
            m.bias.uniform_(-0.5, 0.5)
assert conv_net.training and linear_net.training
x = torch.from_numpy(np.concatenate([kat_inputs.smooth_images(PAIRS, seed=21), kat_inputs.smooth_images(PAIRS, seed=22)]))
blob = {"pairs": np.int64(PAIRS)}     # the inputs are kat_inputs.smooth_images(PAIRS, seed=21 | 22): not stored
for prefix, net in (("conv.", conv_net), ("linear.", linear_net)):
    for k, v in net.state_dict().items():
        if not k.endswith("num_batches_tracked") and "running" not in k:
            blob[prefix + k] = v.detach().numpy().copy()
with torch.no_grad():
    inter = conv_net(x)
    res = linear_net(inter)
    loss, logits_ab, labels = ContrastiveLoss(batch_size=PAIRS)(res)
    loss_t, logits_t, _ = ContrastiveLoss(batch_size=PAIRS, temperature=0.5, h_norm=False)(res)
blob.update(intermediate=inter.numpy(), projection=res.numpy(), loss=np.float32(loss.item()), logits_ab=logits_ab.numpy(),
            loss_t05_nonorm=np.float32(loss_t.item()), logits_ab_t05_nonorm=logits_t.numpy())
out = os.path.join(HERE, "contrastive_kat.npz")
np.savez_compressed(out, **blob)
print(out, os.path.getsize(out), "bytes; loss", loss.item(), "loss(T=0.5, no norm)", loss_t.item())
