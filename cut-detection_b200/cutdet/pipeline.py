"""The per-GPU frame pipeline: decoded frames in (device-resident or pinned host memory), run table out.

    frames --K1--> conv stack --K3--> logits --K4--> (label, max logit) --K5--> run table        (per chunk)
    ... finish(): close the open run;  smooth(): K6 glue_orphans + combine_adjacent_segments

One instance serves one contiguous time range of a video (the whole video on one GPU, or one rank's shard) -- or, with
``n_ranges`` > 1, several time ranges that arrive interleaved (the decode workers of cutdet.decode each own one): every range
has its own streaming run-length encoder, the kernels, buffers and staging are shared, and ``finish_ranges()`` joins the
ranges' tables on the device.  Host frames are uploaded on a copy stream into one of two device buffers (only the source rows
the resize reads), so the copy of chunk i+1 overlaps the kernels of chunk i.

``lanes`` > 1: consecutive chunks are scored on alternating streams, each with its own workspace and result buffers, so the
kernels of chunk i+1 take the SMs chunk i leaves idle -- the last, partial round of its persistent frame kernel (4,050 frames =
27.4 rounds of 148) and the narrow kernels behind it (conv3, the head, K4, K5).  Only the run-length encoder is shared: its
appends are chained by events in chunk order, so the table is the same as with one lane.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi, engine


class _Lane:
    """Where one chunk is scored: a stream (None = the caller's current stream), a conv workspace (None = the net's own) and
    the chunk's logits / labels / max logits."""

    def __init__(self, net: engine.NativeNet, plan: engine.ResizePlan, max_chunk: int, device, side: bool):
        self.stream = torch.cuda.Stream(device=device) if side else None
        self.ws = torch.empty(net.workspace_bytes(max_chunk, plan.dst_h, plan.dst_w), dtype=torch.uint8, device=device) if side else None
        self.logits = torch.empty((max_chunk, net.out_features), dtype=torch.float32, device=device)
        self.labels = torch.empty(max_chunk, dtype=torch.uint8, device=device)
        self.top = torch.empty(max_chunk, dtype=torch.float32, device=device)
        self.done = None            # the lane's most recent chunk has been scored and appended


class FramePipeline:
    def __init__(self, net: engine.NativeNet, plan: engine.ResizePlan, max_chunk: int, table_capacity: int,
                 device="cuda", n_ranges: int = 1, lanes: int = 1):
        self.net, self.plan = net, plan
        self.device = torch.device(device)
        self.max_chunk = int(max_chunk)
        self.encoders = [engine.RunLengthEncoder(table_capacity, self.device) for _ in range(max(1, int(n_ranges)))]
        self.encoder = self.encoders[0]
        self.range_frames = [0] * len(self.encoders)
        lanes = max(1, int(lanes))
        self._lanes = [_Lane(net, plan, self.max_chunk, self.device, side=lanes > 1) for _ in range(lanes)]
        self._lane_turn = 0
        self._appended = [None] * len(self.encoders)    # per range: the event behind its most recent append (lanes > 1)
        # the most recent chunk's results (with lanes > 1: valid on the caller's stream after wait_results())
        self.logits, self.labels, self.top = self._lanes[0].logits, self._lanes[0].labels, self._lanes[0].top
        self.n_frames = 0
        self.h2d_bytes = 0
        self._stage = None          # two row-compacted device buffers for host uploads
        self._stage_events = None
        self._copy_stream = None
        self._turn = 0
        if lanes == 1:
            net.workspace(self.max_chunk, plan.dst_h, plan.dst_w, self.device)

    # ------------------------------------------------------------------ per chunk
    def _score(self, frames: torch.Tensor, compact: bool, rng: int = 0, ready: "torch.cuda.Event | None" = None):
        """Score one chunk and append it to range ``rng``'s encoder.  ``ready``: an event the frames become valid behind (the
        upload); work already queued on the caller's stream is always waited for.  Returns the event that marks the chunk done."""
        n = frames.shape[0]
        if n > self.max_chunk:
            raise ValueError(f"chunk of {n} frames exceeds max_chunk={self.max_chunk}")
        lane = self._lanes[self._lane_turn]
        self._lane_turn = (self._lane_turn + 1) % len(self._lanes)
        caller = torch.cuda.current_stream()
        stream = lane.stream or caller
        if lane.stream is not None:
            # frames produced on the caller's stream, a reset() queued there, the caller's reads of this lane's previous results
            queued = torch.cuda.Event()
            queued.record(caller)
            stream.wait_event(queued)
            frames.record_stream(stream)        # the allocator must not hand the frames' memory out while the lane reads it
        if ready is not None:
            stream.wait_event(ready)
        with torch.cuda.stream(stream):
            logits, labels, top = lane.logits[:n], lane.labels[:n], lane.top[:n]
            self.net.forward_frames(self.plan, frames, compact, out=logits, ws=lane.ws)
            _cabi.check(_cabi.lib().cutdet_argmax(logits.data_ptr(), n, logits.shape[1], labels.data_ptr(), top.data_ptr(),
                                                  stream.cuda_stream))
            if lane.stream is not None and self._appended[rng] is not None:
                stream.wait_event(self._appended[rng])          # the encoder takes the chunks in order
            self.encoders[rng].append(labels, top)
            done = torch.cuda.Event()
            done.record(stream)
        lane.done = done
        self._appended[rng] = done
        self.logits, self.labels, self.top = lane.logits, lane.labels, lane.top
        self.range_frames[rng] += n
        self.n_frames += n
        return done

    def _join(self) -> None:
        """The caller's stream waits for every lane (before the encoders are finished or reset there)."""
        caller = torch.cuda.current_stream()
        for lane in self._lanes:
            if lane.stream is not None and lane.done is not None:
                caller.wait_event(lane.done)

    def wait_results(self) -> None:
        """Make the caller's current stream wait for the most recent chunk: ``logits`` / ``labels`` / ``top`` may then be read
        there (nothing to do with one lane, where the chunk was scored on that stream)."""
        lane = self._lanes[(self._lane_turn - 1) % len(self._lanes)]
        if lane.stream is not None and lane.done is not None:
            torch.cuda.current_stream().wait_event(lane.done)

    def push_device(self, frames: torch.Tensor, compact: bool = False, rng: int = 0) -> None:
        """Frames already in HBM: uint8 BGR HWC [n, rows, w, 3]."""
        self._score(frames, compact, rng)

    def push_host(self, frames: torch.Tensor, compact: bool = False, rng: int = 0) -> torch.cuda.Event:
        """Decoded frames in host memory, uint8 BGR HWC [n, h, w, 3] (``compact``: [n, len(plan.rows), w, 3], only the source
        rows the resize reads, as the decode workers store them): upload the needed rows, then score.

        OWNERSHIP OF THE HOST BUFFER.  From PINNED memory the copy is asynchronous: the call returns as soon as it is queued and
        the buffer must stay untouched until the returned event has completed (``event.synchronize()`` / ``event.query()``, or
        ``wait_uploaded()`` for the most recent call) -- a decode loop that refills the buffer earlier overwrites frames still
        in flight.  From pageable memory the driver stages the copy, and the call returns only after the buffer has been read."""
        if frames.is_cuda or frames.dtype != torch.uint8 or frames.dim() != 4:
            raise ValueError("push_host takes a uint8 [n, h, w, 3] host tensor")
        n = frames.shape[0]
        rows_in = len(self.plan.rows) if compact else self.plan.src_h
        if tuple(frames.shape[1:]) != (rows_in, self.plan.src_w, 3) or frames.stride(3) != 1 or frames.stride(2) != 3:
            raise ValueError(f"host frames must be dense [{rows_in}, {self.plan.src_w}, 3] rows")
        if self._stage is None:
            shape = (self.max_chunk, len(self.plan.rows), self.plan.src_w, 3)
            self._stage = [torch.empty(shape, dtype=torch.uint8, device=self.device) for _ in range(2)]
            self._stage_events = [None, None]
            self._copy_stream = torch.cuda.Stream(device=self.device)
        slot = self._turn
        self._turn ^= 1
        if self._stage_events[slot] is not None:
            self._copy_stream.wait_event(self._stage_events[slot])      # kernels that read this buffer are done
        if compact:         # already row-compacted on the host: one contiguous copy
            with torch.cuda.stream(self._copy_stream):
                self._stage[slot][:n].copy_(frames, non_blocking=True)
            self.h2d_bytes += frames.numel()
        else:
            copied = C.c_int64()
            _cabi.check(_cabi.lib().cutdet_upload_frames(self.plan.handle, frames.data_ptr(), n, frames.stride(0), frames.stride(1),
                                                         self._stage[slot].data_ptr(), self._copy_stream.cuda_stream,
                                                         C.byref(copied)))
            self.h2d_bytes += copied.value
        uploaded = torch.cuda.Event()
        uploaded.record(self._copy_stream)
        self._last_upload = uploaded
        if not frames.is_pinned():
            uploaded.synchronize()          # pageable source: never leave a copy in flight behind the caller's back
        self._stage_events[slot] = self._score(self._stage[slot][:n], True, rng, ready=uploaded)
        return uploaded

    def wait_uploaded(self) -> None:
        """Block until the host buffer of the most recent push_host() has been read (it may then be refilled)."""
        if getattr(self, "_last_upload", None) is not None:
            self._last_upload.synchronize()

    # ------------------------------------------------------------------ end of the range
    def finish(self) -> engine.DeviceRunTable:
        """Close the open run; the table then equals Segmentation(scores).te for this range (local frame numbers)."""
        if len(self.encoders) > 1:
            return self.finish_ranges()[0]
        self._join()
        return self.encoder.finish()

    def finish_ranges(self, capacity: int | None = None):
        """Several time ranges: close every range's open run and join the tables in range order (cutdet_shard_pack +
        cutdet_stitch_packed, no host synchronisation).  Returns (table of the whole sequence, total_frames device tensor)."""
        from . import shard
        self._join()
        tables = [e.finish() for e in self.encoders]
        return shard.stitch_local(tables, self.range_frames, capacity or shard.DEFAULT_CAPACITY)

    def reset(self) -> None:
        self._join()
        for e in self.encoders:
            e.reset()
        self.range_frames = [0] * len(self.encoders)
        self._appended = [None] * len(self.encoders)
        self.n_frames = 0
        self.h2d_bytes = 0


def smooth(table: engine.DeviceRunTable, real_threshold: int = 100, blank_threshold: int = 10) -> engine.DeviceRunTable:
    """K6 on a finished table: glue_orphans then combine_adjacent_segments (segment_video.py:64-66).  Queued without a host
    synchronisation: a lone-orphan run (the reference's IndexError) is raised by the table's next count() / to_te()."""
    table.glue_orphans(real_threshold, blank_threshold, defer_status=True)
    table.combine_adjacent()
    return table
