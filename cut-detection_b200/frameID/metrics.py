"""Mirror of the reference's ``frameID/metrics.py`` (forward pass): ``ContrastiveLoss`` keeps the constructor, the buffers and
the ``(loss, logits_ab, labels)`` return of metrics.py:9-47, but the arithmetic runs in libcutdet_b200.so
(``cutdet_contrastive_loss``, csrc/contrastive.cu).  Forward only: the loss carries no autograd graph."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from cutdet import engine as _engine

BIG_NUMBER = 1e9


class ContrastiveLoss(nn.Module):
    """NT-Xent over a batch of ``batch_size`` image pairs: ``forward(x)`` takes the ``[2 * batch_size, D]`` projections of
    ``torch.cat((x_t1, x_t2))`` (learn_contrasts.py:104-108)."""

    def __init__(self, batch_size: int = 32, h_norm=True, temperature=1.0):
        super().__init__()
        self.h_norm = h_norm
        self.temperature = temperature
        self.register_buffer("labels", torch.arange(batch_size))
        self.register_buffer("masks", F.one_hot(torch.arange(batch_size), batch_size))

    def forward(self, x):
        if x.shape[0] != 2 * self.labels.shape[0]:
            # the reference fails inside its broadcasts/cross_entropy for any other batch; say why
            raise RuntimeError(f"ContrastiveLoss(batch_size={self.labels.shape[0]}) got {x.shape[0]} rows, expected "
                               f"{2 * self.labels.shape[0]}")
        loss, logits_ab = _engine.contrastive_loss(x, temperature=self.temperature, h_norm=bool(self.h_norm))
        return loss, logits_ab, self.labels


class SummedCrossEntropy(nn.Module):
    """Addition to the reference API: what ``torch.nn.CrossEntropyLoss(reduction="sum")`` computes in the reference's supervised
    script (training_scripts/supervised_training.py:132, 148, 186) on the native path, forward only.  ``forward(pred, labels)``
    returns the summed loss; ``accuracy_counts(pred, labels)`` also returns the validation loop's per-class ``(correct, total)``
    tensors (:188-193), all from one kernel."""

    def forward(self, pred, labels):
        return _engine.cross_entropy_sum(pred, labels)

    @staticmethod
    def accuracy_counts(pred, labels):
        return _engine.cross_entropy_sum(pred, labels, with_counts=True)
