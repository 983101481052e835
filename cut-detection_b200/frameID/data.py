"""frameID.data -- the reference's frame source API (reference frameID/data.py) over the B200 kernels.

    open_video(path) -> (cv2.VideoCapture, {"fps", "length", "width", "height"})          data.py:13-31
    VideoDataset(file_path, resize=None)   iterable of [3, H', W'] float32 RGB tensors       data.py:184-234
    SupervisedFrameDataset.lab_enum        {"a22": 0, "ez": 1, "b": 2}                        data.py:116

Decoding stays on the host (cv2 / FFmpeg: outside the hot path, excluded from every timing).  What the reference
does to each decoded frame on the CPU -- cv2.resize(INTER_LINEAR), float conversion, HWC->CHW, BGR->RGB, /255
(data.py:218-228) -- is the K1 CUDA kernel here, bit-exact with the reference's result.  Consequently the tensors
this dataset yields live on the GPU; ``batch.to(device)`` in the reference's loop becomes a no-op.

``VideoDataset`` keeps the reference's one-frame-at-a-time contract.  The CLI does not go through it: it decodes with worker
processes into a reusable pinned ring and streams chunks through ``cutdet.pipeline.FramePipeline`` (cutdet/decode.py,
cut-detection_b200/segment_video.py).
"""
from __future__ import annotations

import cv2
import numpy as np
import torch
from torch.utils.data import IterableDataset

from cutdet import engine as _engine


def open_video(video_path):
    """Open a video with OpenCV and report its basic properties as ints."""
    cap = cv2.VideoCapture(video_path)
    props = {"fps": cv2.CAP_PROP_FPS, "length": cv2.CAP_PROP_FRAME_COUNT, "width": cv2.CAP_PROP_FRAME_WIDTH,
             "height": cv2.CAP_PROP_FRAME_HEIGHT}
    return cap, {name: int(cap.get(prop)) for name, prop in props.items()}


class SupervisedFrameDataset:
    """Only the label enumeration is on the inference path (segmentation imports it); the training-time image
    folder dataset itself is out of scope for this build."""

    lab_enum = {"a22": 0, "ez": 1, "b": 2}

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("training datasets are outside the hot path this build implements")


class VideoDataset(IterableDataset):
    """Frames of a video, preprocessed for the classifier.  Single pass, like the reference (``__iter__`` returns
    self; a second iteration yields nothing because the capture is exhausted)."""

    def __init__(self, file_path, resize=None, device="cuda"):
        super().__init__()
        self.cap, self.video_info = open_video(file_path)
        self.device = torch.device(device)
        if resize is not None:
            self.new_width, self.new_height = _engine.target_size(self.video_info["width"], self.video_info["height"], resize)
        else:
            self.new_width = None
            self.new_height = None
        self._plan = None

    @property
    def plan(self) -> _engine.ResizePlan:
        """Resize geometry, built on first use (the capture reports the true frame size only after a read for some
        containers, so the decoded frame's shape is authoritative)."""
        return self._plan

    def _plan_for(self, frame: np.ndarray) -> _engine.ResizePlan:
        h, w = frame.shape[:2]
        if self._plan is None or (self._plan.src_h, self._plan.src_w) != (h, w):
            if self.new_width is None:
                self._plan = _engine.ResizePlan(h, w, h, w)
            else:
                self._plan = _engine.ResizePlan(h, w, self.new_height, self.new_width)
        return self._plan

    def __iter__(self):
        return self

    def __next__(self):
        ret, frame = self.cap.read()
        if not ret:
            raise StopIteration
        plan = self._plan_for(frame)
        dev = torch.from_numpy(np.ascontiguousarray(frame)).to(self.device)[None]
        return _engine.preprocess_f32(plan, dev)[0]

    def __len__(self):
        """Frame count reported by the container."""
        return self.video_info["length"]
