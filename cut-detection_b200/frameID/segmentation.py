"""frameID.segmentation -- the reference's Segmentation API (reference frameID/segmentation.py) over the B200 kernels.

    Segmentation(scores)              .te dict of five tensors, len()                      segmentation.py:26-63
    .glue_orphans(real_threshold=100, blank_threshold=10)                                   segmentation.py:91-166
    .combine_adjacent_segments()                                                            segmentation.py:168-183
    .write_csv(file_path)             "start_frame,label" rows, CRLF, no header             segmentation.py:185-196

``scores`` is the [N, classes] float32 tensor of raw logits; it may live on the CPU (as in the reference CLI, which
moves it there first) or on the GPU.  Construction runs K4 (max / first argmax per frame) and K5 (run-length
encoding with per-run mean of the max logit) on the GPU; the two smoothing passes run K6.  ``.te`` holds CPU tensors
with the reference's keys and dtypes and may be edited between calls, exactly as with the reference.

Two documented deviations, both inside what the reference leaves unspecified or at rounding level:
  * ties between exactly equal orphan means go to the lowest run index (torch.argsort is unstable);
  * per-run means are float64 sums rounded once to float32 (the reference's float32 .mean() differs in the last ulp).
"""
from __future__ import annotations

import csv

import torch

from cutdet import engine as _engine
from .data import SupervisedFrameDataset

_TYPE_MAP = SupervisedFrameDataset.lab_enum
_INVERSE_TYPE_MAP = {v: k for k, v in _TYPE_MAP.items()}


def _device_for(t: torch.Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("Segmentation needs a CUDA device: this build has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


class Segmentation:
    """Per-frame scores -> run table -> smoothed segments."""

    def __init__(self, scores):
        dev = _device_for(scores)
        scores = scores.to(device=dev, dtype=torch.float32)
        if scores.dim() != 2 or scores.shape[0] == 0:
            raise IndexError(f"scores must be a non-empty [N, classes] tensor, got {tuple(scores.shape)}")
        self._device = dev
        self.te = _engine.run_table_from_scores(scores).to_te()

    @classmethod
    def from_table(cls, table: "_engine.DeviceRunTable") -> "Segmentation":
        """Wrap a run table that is already on the device (streaming / multi-GPU path)."""
        self = cls.__new__(cls)
        self._device = table.device
        self.te = table.to_te()
        return self

    def __len__(self):
        return self.te["end_frames"].shape[0]

    def _run(self, method: str, *args) -> None:
        table = _engine.DeviceRunTable.from_te(self.te, self._device)
        getattr(table, method)(*args)
        self.te = table.to_te()

    def glue_orphans(self, real_threshold=100, blank_threshold=10):
        """Merge too-short runs into a neighbour, least confident first (raises IndexError on a lone orphan run,
        like the reference)."""
        self._run("glue_orphans", real_threshold, blank_threshold)

    def combine_adjacent_segments(self):
        """Combine adjacent segments of the same type."""
        self._run("combine_adjacent")

    def write_csv(self, file_path):
        """One ``start_frame,label`` row per segment; csv module defaults give CRLF line ends."""
        with open(file_path, "w", newline="") as f:
            out = csv.writer(f, delimiter=",")
            for start, kind in zip(self.te["start_frames"].tolist(), self.te["frame_types"].tolist()):
                out.writerow((start, _INVERSE_TYPE_MAP[kind]))
