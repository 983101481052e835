"""Thin torch-facing wrappers over the C ABI: torch only supplies device memory and the stream.

Every function here launches CUDA kernels from libcutdet_b200.so on the current torch stream and raises if the
library is missing or the tensors are not on a CUDA device -- nothing is computed by PyTorch or on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi

TABLE_COLUMNS = ("end_frames", "frame_types", "run_lengths", "start_frames", "score_means")  # Segmentation.te keys


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: this build has no CPU path (got device {t.device})")


def device_check() -> dict:
    sm, major, minor = C.c_int(), C.c_int(), C.c_int()
    _cabi.check(_cabi.lib().cutdet_device_check(C.byref(sm), C.byref(major), C.byref(minor)))
    return {"sm_count": sm.value, "cc": (major.value, minor.value)}


def target_size(width: int, height: int, resize: int = 256) -> tuple[int, int]:
    """(new_width, new_height) per frameID/data.py:199-202."""
    nw, nh = C.c_int(), C.c_int()
    _cabi.check(_cabi.lib().cutdet_target_size(width, height, resize, C.byref(nw), C.byref(nh)))
    return nw.value, nh.value


def resize_rows(height: int, width: int, resize: int = 256) -> np.ndarray:
    """Source rows the resize of a height x width frame reads (what ``ResizePlan.for_video(...).rows`` lists), computed on the
    host without any CUDA call: a frame source can start gathering rows before the process has a CUDA context."""
    nw, nh = target_size(width, height, resize)
    n = C.c_int()
    _cabi.check(_cabi.lib().cutdet_resize_rows(height, width, nh, nw, None, C.byref(n)))
    rows = (C.c_int * max(n.value, 1))()
    _cabi.check(_cabi.lib().cutdet_resize_rows(height, width, nh, nw, rows, C.byref(n)))
    return np.frombuffer(rows, dtype=np.int32)[:n.value].copy()


# ----------------------------------------------------------------------------------------------- K1
class ResizePlan:
    """Tap tables of one cv2.resize(INTER_LINEAR) geometry, resident on the current device."""

    def __init__(self, src_h: int, src_w: int, dst_h: int, dst_w: int):
        self.src_h, self.src_w, self.dst_h, self.dst_w = src_h, src_w, dst_h, dst_w
        handle = C.c_void_p()
        _cabi.check(_cabi.lib().cutdet_resize_plan_create(src_h, src_w, dst_h, dst_w, C.byref(handle)))
        self.handle = handle
        n = C.c_int()
        _cabi.check(_cabi.lib().cutdet_resize_plan_rows(self.handle, None, C.byref(n)))
        rows = (C.c_int * n.value)()
        _cabi.check(_cabi.lib().cutdet_resize_plan_rows(self.handle, rows, C.byref(n)))
        self.rows = np.frombuffer(rows, dtype=np.int32).copy()   # source rows the resize reads

    @classmethod
    def for_video(cls, height: int, width: int, resize: int = 256) -> "ResizePlan":
        nw, nh = target_size(width, height, resize)
        return cls(height, width, nh, nw)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _cabi.lib().cutdet_resize_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def _frames_struct(plan: ResizePlan, frames: torch.Tensor, compact: bool) -> _cabi.Frames:
    _need_cuda(frames, "frames")
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
        raise ValueError(f"frames must be uint8 [B, rows, width, 3], got {frames.dtype} {tuple(frames.shape)}")
    rows = len(plan.rows) if compact else plan.src_h
    if frames.shape[1] != rows or frames.shape[2] != plan.src_w:
        raise ValueError(f"frames are {tuple(frames.shape[1:3])}, the plan expects ({rows}, {plan.src_w})")
    if frames.stride(3) != 1 or frames.stride(2) != 3:
        frames = frames.contiguous()
    return _cabi.Frames(frames.data_ptr(), frames.stride(0), frames.stride(1), frames.shape[0], 1 if compact else 0), frames


def preprocess_f32(plan: ResizePlan, frames: torch.Tensor, compact: bool = False, out: torch.Tensor | None = None):
    """uint8 BGR HWC frames -> float32 RGB [B,3,H2,W2] in [0,1] (what VideoDataset yields, stacked)."""
    fs, keep = _frames_struct(plan, frames, compact)
    if out is None:
        out = torch.empty((frames.shape[0], 3, plan.dst_h, plan.dst_w), dtype=torch.float32, device=frames.device)
    _cabi.check(_cabi.lib().cutdet_preprocess_f32(plan.handle, C.byref(fs), out.data_ptr(), _stream()))
    return out


def preprocess_u8(plan: ResizePlan, frames: torch.Tensor, compact: bool = False):
    """uint8 BGR HWC frames -> cv2.resize'd uint8 BGR HWC [B,H2,W2,3]."""
    fs, keep = _frames_struct(plan, frames, compact)
    out = torch.empty((frames.shape[0], plan.dst_h, plan.dst_w, 3), dtype=torch.uint8, device=frames.device)
    _cabi.check(_cabi.lib().cutdet_preprocess_u8(plan.handle, C.byref(fs), out.data_ptr(), _stream()))
    return out


# ----------------------------------------------------------------------------------------------- K2/K3
def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


class NativeNet:
    """A cutdet_net built from state_dict-style float32 arrays.

    ``weights`` uses the reference's keys with the conv state_dict prefixed ``conv.`` and the linear one
    ``linear.`` (the layout of oracle.net / prod_net_weights.npz); either part may be absent (a bare
    FrameConvNet or FrameLinearNet)."""

    def __init__(self, weights: dict, avg_pool_size: int, bn_eps: float = 1e-5, input_channels: int | None = None):
        n_conv = 0
        while f"conv.conv_layers.{n_conv}.conv.weight" in weights:
            n_conv += 1
        n_fc = 0
        while f"linear.layers.{n_fc}.linear.weight" in weights:
            n_fc += 1
        if n_conv + n_fc == 0:
            raise ValueError("no layers found in the weight dictionary")
        cfg = _cabi.NetConfig()
        if n_conv:
            w0 = weights["conv.conv_layers.0.conv.weight"]
            cfg.input_channels, cfg.hidden_channels = int(w0.shape[1]), int(w0.shape[0])
            cfg.avg_pool_size = int(avg_pool_size)
            cfg.fc_input_size = cfg.hidden_channels * cfg.avg_pool_size ** 2
        else:
            fin = int(weights["linear.layers.0.linear.weight"].shape[1])
            cfg.input_channels, cfg.hidden_channels, cfg.avg_pool_size, cfg.fc_input_size = fin, fin, 1, fin
        cfg.n_conv_layers, cfg.n_fc_layers = n_conv, n_fc
        if n_fc:
            cfg.fc_hidden_size = int(weights["linear.layers.0.linear.weight"].shape[0])
            cfg.fc_output_size = int(weights[f"linear.layers.{n_fc - 1}.linear.weight"].shape[0])
            if int(weights["linear.layers.0.linear.weight"].shape[1]) != cfg.fc_input_size:
                raise ValueError("first FC layer does not match conv_channels * avg_pool_size^2")
        else:
            cfg.fc_hidden_size = cfg.fc_output_size = cfg.fc_input_size
        self.cfg = cfg
        self.out_features = cfg.fc_output_size if n_fc else cfg.fc_input_size
        lib = _cabi.lib()
        handle = C.c_void_p()
        _cabi.check(lib.cutdet_net_create(C.byref(cfg), C.byref(handle)))
        self.handle = handle
        for i in range(n_conv):
            p = f"conv.conv_layers.{i}"
            arrs = [_f32(weights[p + s]) for s in (".conv.weight", ".conv.bias", ".bn.weight", ".bn.bias",
                                                   ".bn.running_mean", ".bn.running_var")]
            _cabi.check(lib.cutdet_net_set_conv_layer(self.handle, i, *[_cabi.fptr(a) for a in arrs], bn_eps))
        for j in range(n_fc):
            p = f"linear.layers.{j}"
            arrs = [_f32(weights[p + ".linear.weight"]), _f32(weights[p + ".linear.bias"])]
            if j < n_fc - 1 or (n_conv == 0 and n_fc == 1 and p + ".bn.weight" in weights):     # a lone FCLayer may carry a BatchNorm
                arrs += [_f32(weights[p + s]) for s in (".bn.weight", ".bn.bias", ".bn.running_mean", ".bn.running_var")]
            else:
                arrs += [None] * 4
            _cabi.check(lib.cutdet_net_set_fc_layer(self.handle, j, *[_cabi.fptr(a) for a in arrs], bn_eps))
        _cabi.check(lib.cutdet_net_finalize(self.handle))
        self._ws = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _cabi.lib().cutdet_net_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def set_option(self, name: str, value: int) -> None:
        """Scheduling / precision switches of the tensor-core path (include/cutdet_b200.h CUTDET_OPT_*): ``conv1_acc32``,
        ``sub_batch``, ``group_frames``, ``no_pdl``, ``conv1_grid``.  They are per net; nothing is read from the environment."""
        if name not in _cabi.NET_OPTIONS:
            raise ValueError(f"unknown option {name!r}; known: {sorted(_cabi.NET_OPTIONS)}")
        _cabi.check(_cabi.lib().cutdet_net_set_option(self.handle, _cabi.NET_OPTIONS[name], int(value)))
        self._ws = None             # the workspace layout depends on sub_batch / group_frames

    def get_option(self, name: str) -> int:
        v = C.c_int()
        _cabi.check(_cabi.lib().cutdet_net_get_option(self.handle, _cabi.NET_OPTIONS[name], C.byref(v)))
        return v.value

    def forward_conv_layer(self, layer: int, x: torch.Tensor, bn_mode: int = 1) -> torch.Tensor:
        """CNNLayer.forward (reference frameID/net.py:33-40) of conv layer ``layer`` alone: float32 [B,Cin,H,W] ->
        [B,Cout,H//3,W//3].  bn_mode 0 = no BatchNorm, 1 = running statistics, 2 = batch statistics."""
        _need_cuda(x, "input")
        if x.dtype != torch.float32 or x.dim() != 4:
            raise ValueError(f"input must be float32 [B,C,H,W], got {x.dtype} {tuple(x.shape)}")
        cin = self.cfg.input_channels if layer == 0 else self.cfg.hidden_channels
        if x.shape[1] != cin:
            raise ValueError(f"input has {x.shape[1]} channels, the layer expects {cin}")
        x = x.contiguous()
        b, _, h, w = x.shape
        if h < 3 or w < 3:
            raise ValueError(f"input {h}x{w} is smaller than the 3x3 pool")
        out = torch.empty((b, self.cfg.hidden_channels, h // 3, w // 3), dtype=torch.float32, device=x.device)
        _cabi.check(_cabi.lib().cutdet_net_forward_conv_layer(self.handle, layer, x.data_ptr(), b, h, w, out.data_ptr(),
                                                              int(bn_mode), _stream()))
        return out

    def forward_fc_layer(self, layer: int, x: torch.Tensor, relu: bool, bn_mode: int) -> torch.Tensor:
        """FCLayer.forward (reference frameID/net.py:62-68) of FC layer ``layer`` alone: float32 [B,in] -> [B,out]."""
        _need_cuda(x, "input")
        if x.dtype != torch.float32 or x.dim() != 2:
            raise ValueError(f"input must be float32 [B,in], got {x.dtype} {tuple(x.shape)}")
        x = x.contiguous()
        n_out = self.cfg.fc_output_size if layer == self.cfg.n_fc_layers - 1 else self.cfg.fc_hidden_size
        n_in = self.cfg.fc_input_size if layer == 0 else self.cfg.fc_hidden_size
        if x.shape[1] != n_in:
            raise ValueError(f"input has {x.shape[1]} features, the layer expects {n_in}")
        out = torch.empty((x.shape[0], n_out), dtype=torch.float32, device=x.device)
        _cabi.check(_cabi.lib().cutdet_net_forward_fc_layer(self.handle, layer, x.data_ptr(), x.shape[0], out.data_ptr(),
                                                            1 if relu else 0, int(bn_mode), _stream()))
        return out

    def uses_tensor_cores(self, height: int, width: int) -> bool:
        return bool(_cabi.lib().cutdet_net_uses_tensor_cores(self.handle, height, width))

    def workspace(self, batch: int, height: int, width: int, device) -> torch.Tensor:
        need = C.c_size_t()
        _cabi.check(_cabi.lib().cutdet_net_workspace_bytes(self.handle, batch, height, width, C.byref(need)))
        if self._ws is None or self._ws.numel() < need.value or self._ws.device != torch.device(device):
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=device)
        return self._ws

    def forward_f32(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """x float32 [B,C,H,W] -> float32 [B, out_features]."""
        _need_cuda(x, "input")
        if x.dim() == 2:
            x = x[:, :, None, None]
        if x.dtype != torch.float32 or x.dim() != 4:
            raise ValueError(f"input must be float32 [B,C,H,W], got {x.dtype} {tuple(x.shape)}")
        if x.shape[1] != self.cfg.input_channels:
            raise ValueError(f"input has {x.shape[1]} channels, the net expects {self.cfg.input_channels}")
        x = x.contiguous()
        b, _, h, w = x.shape
        if out is None:
            out = torch.empty((b, self.out_features), dtype=torch.float32, device=x.device)
        ws = self.workspace(b, h, w, x.device)
        _cabi.check(_cabi.lib().cutdet_net_forward_f32(self.handle, x.data_ptr(), b, h, w, out.data_ptr(), ws.data_ptr(),
                                                       ws.numel(), _stream()))
        return out

    def forward_f32_batchstats(self, x: torch.Tensor, out: torch.Tensor | None = None, tensor_cores: bool = True) -> torch.Tensor:
        """The forward pass of a module left in training mode: every BatchNorm normalises with this batch's mean and biased
        variance (reference training_scripts/learn_contrasts.py:100-107).  Forward only.  ``tensor_cores=False`` keeps the
        float32 CUDA-core kernels (also used for architectures and batch sizes the tensor-core path does not cover)."""
        _need_cuda(x, "input")
        if x.dim() == 2:
            x = x[:, :, None, None]
        if x.dtype != torch.float32 or x.dim() != 4:
            raise ValueError(f"input must be float32 [B,C,H,W], got {x.dtype} {tuple(x.shape)}")
        if x.shape[1] != self.cfg.input_channels:
            raise ValueError(f"input has {x.shape[1]} channels, the net expects {self.cfg.input_channels}")
        if x.shape[0] < 2:
            raise ValueError("Expected more than 1 value per channel when training")      # what nn.BatchNorm raises
        x = x.contiguous()
        b, _, h, w = x.shape
        if out is None:
            out = torch.empty((b, self.out_features), dtype=torch.float32, device=x.device)
        ws = self.workspace(b, h, w, x.device)
        _cabi.check(_cabi.lib().cutdet_net_forward_f32_batchstats(self.handle, x.data_ptr(), b, h, w, out.data_ptr(),
                                                                  ws.data_ptr(), ws.numel(), 1 if tensor_cores else 0, _stream()))
        return out

    def workspace_bytes(self, batch: int, height: int, width: int) -> int:
        need = C.c_size_t()
        _cabi.check(_cabi.lib().cutdet_net_workspace_bytes(self.handle, batch, height, width, C.byref(need)))
        return need.value

    def forward_frames(self, plan: ResizePlan, frames: torch.Tensor, compact: bool = False,
                       out: torch.Tensor | None = None, ws: torch.Tensor | None = None) -> torch.Tensor:
        """Decoded uint8 BGR HWC frames -> logits, K1 fused in front of the conv stack.  ``ws``: a workspace of the caller's
        (``workspace_bytes``) instead of the net's own -- calls that run side by side on different streams each need one."""
        fs, keep = _frames_struct(plan, frames, compact)
        b = frames.shape[0]
        if out is None:
            out = torch.empty((b, self.out_features), dtype=torch.float32, device=frames.device)
        if ws is None:
            ws = self.workspace(b, plan.dst_h, plan.dst_w, frames.device)
        _cabi.check(_cabi.lib().cutdet_net_forward_frames(self.handle, plan.handle, C.byref(fs), out.data_ptr(),
                                                          ws.data_ptr(), ws.numel(), _stream()))
        return out

    def debug_conv_output(self, layer: int, batch: int, height: int, width: int) -> torch.Tensor:
        """float32 NCHW output of conv layer ``layer`` from the last forward (test hook)."""
        c = self.cfg.hidden_channels
        h, w = height, width
        for _ in range(layer + 1):
            h, w = h // 3, w // 3
        out = torch.empty((batch, c, h, w), dtype=torch.float32, device=self._ws.device)
        _cabi.check(_cabi.lib().cutdet_net_debug_conv_output(self.handle, layer, batch, height, width,
                                                             self._ws.data_ptr(), out.data_ptr(), _stream()))
        return out


# ----------------------------------------------------------------------------------------------- K4
def argmax(scores: torch.Tensor):
    """scores [N,C] float32 -> (labels uint8 [N], top float32 [N]): torch.max(scores, 1) with first-index ties."""
    _need_cuda(scores, "scores")
    if scores.dtype != torch.float32 or scores.dim() != 2:
        raise ValueError(f"scores must be float32 [N,C], got {scores.dtype} {tuple(scores.shape)}")
    scores = scores.contiguous()
    n, c = scores.shape
    labels = torch.empty(n, dtype=torch.uint8, device=scores.device)
    top = torch.empty(n, dtype=torch.float32, device=scores.device)
    _cabi.check(_cabi.lib().cutdet_argmax(scores.data_ptr(), n, c, labels.data_ptr(), top.data_ptr(), _stream()))
    return labels, top


# ----------------------------------------------------------------------------------------------- K5/K6
class ShardOverflow(RuntimeError):
    """A shard's run table had more rows than the gather capacity of the exchange step (cutdet.shard)."""

    def __init__(self, needed: int, capacity):
        super().__init__(f"a shard's run table has {needed} runs, the gather capacity is {capacity}")
        self.needed = needed


class DeviceRunTable:
    """A cutdet_run_table backed by torch tensors."""

    def __init__(self, capacity: int, device):
        capacity = max(int(capacity), 1)
        self.capacity = capacity
        self.device = torch.device(device)
        self.end_frames = torch.empty(capacity, dtype=torch.int64, device=device)
        self.start_frames = torch.empty(capacity, dtype=torch.int64, device=device)
        self.run_lengths = torch.empty(capacity, dtype=torch.int64, device=device)
        self.frame_types = torch.empty(capacity, dtype=torch.int32, device=device)
        self.score_means = torch.empty(capacity, dtype=torch.float32, device=device)
        self.score_sums = torch.empty(capacity, dtype=torch.float64, device=device)
        # run count and K6 status side by side: ONE 16-byte device->host copy reads both (to_te / count)
        self._head = torch.zeros(2, dtype=torch.int64, device=device)
        self.n_runs = self._head[0:1]
        self.status = self._head[1:2].view(torch.int32)[0:1]
        self._glue_ws = None        # K6 scratch, owned by this table (the library keeps none of its own)
        self.shard_capacity = None  # set by shard.stitch_packed: what a negative run count is reported against

    def struct(self) -> _cabi.RunTable:
        return _cabi.RunTable(self.end_frames.data_ptr(), self.start_frames.data_ptr(), self.run_lengths.data_ptr(),
                              self.frame_types.data_ptr(), self.score_means.data_ptr(), self.score_sums.data_ptr(),
                              self.capacity)

    @classmethod
    def from_te(cls, te: dict, device, capacity: int | None = None) -> "DeviceRunTable":
        """Upload the five Segmentation.te columns (sums are reconstructed as mean * length)."""
        n = int(te["end_frames"].shape[0])
        t = cls(capacity or n, device)
        t.end_frames[:n] = te["end_frames"].to(device=device, dtype=torch.int64)
        t.start_frames[:n] = te["start_frames"].to(device=device, dtype=torch.int64)
        t.run_lengths[:n] = te["run_lengths"].to(device=device, dtype=torch.int64)
        t.frame_types[:n] = te["frame_types"].to(device=device, dtype=torch.int32)
        t.score_means[:n] = te["score_means"].to(device=device, dtype=torch.float32)
        t.score_sums[:n] = t.score_means[:n].double() * t.run_lengths[:n].double()
        t.n_runs.fill_(n)
        return t

    def count(self) -> int:
        """Run count (one device->host copy = one synchronisation).  Raises what the queued kernels reported: a deferred K6
        status (IndexError for a lone orphan run), ShardOverflow from the stitch, RuntimeError for a table overflow."""
        head = self._head.cpu()
        n, status = int(head[0]), int(head[1]) & 0xFFFFFFFF
        if n < 0:
            raise ShardOverflow(-n - 1, self.shard_capacity)
        if status:
            self.status.zero_()
            _cabi.check(status)
        if n > self.capacity:
            raise RuntimeError(f"run table overflow: {n} runs, capacity {self.capacity}")
        return n

    def to_te(self) -> dict:
        """The five columns as CPU tensors with the reference's dtypes (int64 / float32): the run count, then ONE packed
        device->host copy of the rows (cutdet_shard_pack's layout)."""
        from . import shard
        n = self.count()
        if n == 0:
            rows = np.zeros(0, dtype=shard.ROW_DTYPE)
        else:
            packed = shard.pack_table(self, 0, n).cpu().numpy()
            rows = packed[shard.HEADER_BYTES:].view(shard.ROW_DTYPE)
        return {
            "end_frames": torch.from_numpy(rows["end"].copy()),
            "frame_types": torch.from_numpy(rows["type"].astype(np.int64)),
            "run_lengths": torch.from_numpy(rows["length"].copy()),
            "start_frames": torch.from_numpy(rows["start"].copy()),
            "score_means": torch.from_numpy(rows["mean"].copy()),
        }

    def glue_orphans(self, real_threshold: int = 100, blank_threshold: int = 10, defer_status: bool = False) -> None:
        """K6.  ``defer_status``: do not synchronise here; a lone-orphan status is raised by the next count() / to_te()."""
        ts = self.struct()
        lib = _cabi.lib()
        if self._glue_ws is None:
            self._glue_ws = torch.empty(lib.cutdet_glue_orphans_workspace_bytes(self.capacity), dtype=torch.uint8,
                                        device=self.device)
        _cabi.check(lib.cutdet_glue_orphans(C.byref(ts), self.n_runs.data_ptr(), int(real_threshold), int(blank_threshold),
                                            self.status.data_ptr(), self._glue_ws.data_ptr(), self._glue_ws.numel(), _stream()))
        if not defer_status:
            _cabi.check(int(self.status.item()))      # ELONE_ORPHAN -> IndexError, like the reference

    def combine_adjacent(self) -> None:
        ts = self.struct()
        _cabi.check(_cabi.lib().cutdet_combine_adjacent(C.byref(ts), self.n_runs.data_ptr(), _stream()))


class RunLengthEncoder:
    """Streaming K5: feed (labels, top) chunks in frame order, then ``finish()``."""

    def __init__(self, capacity: int, device):
        self.table = DeviceRunTable(capacity, device)
        self.state = torch.empty(_cabi.lib().cutdet_rle_state_bytes(), dtype=torch.uint8, device=device)
        self.reset()

    def reset(self) -> None:
        _cabi.check(_cabi.lib().cutdet_rle_reset(self.state.data_ptr(), _stream()))

    def append(self, labels: torch.Tensor, top: torch.Tensor) -> None:
        _need_cuda(labels, "labels")
        if labels.dtype != torch.uint8 or top.dtype != torch.float32 or labels.shape != top.shape or labels.dim() != 1:
            raise ValueError("labels must be uint8 [N] and top float32 [N]")
        labels, top = labels.contiguous(), top.contiguous()
        ts = self.table.struct()
        _cabi.check(_cabi.lib().cutdet_rle_append(self.state.data_ptr(), labels.data_ptr(), top.data_ptr(),
                                                  labels.shape[0], C.byref(ts), _stream()))

    def finish(self) -> DeviceRunTable:
        ts = self.table.struct()
        _cabi.check(_cabi.lib().cutdet_rle_finish(self.state.data_ptr(), C.byref(ts), self.table.n_runs.data_ptr(),
                                                  _stream()))
        # no synchronisation here: a table that overflowed reports n_runs > capacity, which count() / to_te() raise on
        return self.table


def run_table_from_scores(scores: torch.Tensor, capacity: int | None = None) -> DeviceRunTable:
    """K4 + K5 over a whole [N,C] score tensor."""
    labels, top = argmax(scores)
    enc = RunLengthEncoder(capacity or max(int(scores.shape[0]), 1), scores.device)
    enc.append(labels, top)
    return enc.finish()


def stitch_shards(shards: "list[DeviceRunTable] | DeviceRunTable", n_runs: torch.Tensor, frame_offsets: torch.Tensor,
                  shard_capacity: int) -> DeviceRunTable:
    """Join per-shard run tables (local frame numbers) into one global table.

    ``shards`` is ONE DeviceRunTable whose rows [i*shard_capacity, ...) hold shard i (e.g. the output of an
    all_gather of equally sized tables)."""
    src = shards
    n_shards = int(n_runs.shape[0])
    dst = DeviceRunTable(src.capacity, src.device)
    s, d = src.struct(), dst.struct()
    _cabi.check(_cabi.lib().cutdet_stitch_shards(C.byref(s), n_shards, int(shard_capacity), n_runs.data_ptr(),
                                                 frame_offsets.data_ptr(), C.byref(d), dst.n_runs.data_ptr(), _stream()))
    return dst


def contrastive_loss(x: torch.Tensor, temperature: float = 1.0, h_norm: bool = True):
    """ContrastiveLoss.forward of the reference (frameID/metrics.py:23-47) on the GPU: x float32 [2B, D] ->
    (loss 0-d tensor, logits_ab [B, B])."""
    _need_cuda(x, "input")
    if x.dtype != torch.float32 or x.dim() != 2 or x.shape[0] % 2 or x.shape[0] == 0:
        raise ValueError(f"input must be float32 [2B, D], got {x.dtype} {tuple(x.shape)}")
    x = x.contiguous()
    pairs, dim = x.shape[0] // 2, x.shape[1]
    loss = torch.empty((), dtype=torch.float32, device=x.device)
    logits_ab = torch.empty((pairs, pairs), dtype=torch.float32, device=x.device)
    lib = _cabi.lib()
    ws = torch.empty(lib.cutdet_contrastive_loss_workspace_bytes(pairs), dtype=torch.uint8, device=x.device)
    _cabi.check(lib.cutdet_contrastive_loss(x.data_ptr(), pairs, dim, float(temperature), 1 if h_norm else 0, loss.data_ptr(),
                                            logits_ab.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return loss, logits_ab


def cross_entropy_sum(logits: torch.Tensor, labels: torch.Tensor, with_counts: bool = False, check_labels: bool = True):
    """torch.nn.CrossEntropyLoss(reduction="sum")(logits, labels) on the GPU (forward only; reference
    training_scripts/supervised_training.py:132, 148) and, with ``with_counts``, the validation loop's per-class
    (correct, total) counters (:188-193).  Returns loss (0-d float32 tensor) or (loss, correct int64 [C], total int64 [C])."""
    _need_cuda(logits, "logits")
    _need_cuda(labels, "labels")
    if logits.dtype != torch.float32 or logits.dim() != 2 or labels.dim() != 1 or labels.shape[0] != logits.shape[0]:
        raise ValueError(f"logits must be float32 [N, C] and labels [N], got {tuple(logits.shape)} / {tuple(labels.shape)}")
    logits, labels = logits.contiguous(), labels.to(torch.int64).contiguous()
    n, c = logits.shape
    lib = _cabi.lib()
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    correct = torch.empty(c, dtype=torch.int64, device=logits.device) if with_counts else None
    total = torch.empty(c, dtype=torch.int64, device=logits.device) if with_counts else None
    ws = torch.empty(lib.cutdet_cross_entropy_workspace_bytes(c) // 8 + 1, dtype=torch.int64, device=logits.device)
    bad = C.c_int(0)
    rc = lib.cutdet_cross_entropy_sum(logits.data_ptr(), labels.data_ptr(), n, c, loss.data_ptr(),
                                      correct.data_ptr() if with_counts else None, total.data_ptr() if with_counts else None,
                                      ws.data_ptr(), ws.numel() * 8, C.byref(bad) if check_labels else None, _stream())
    if rc == _cabi.EINVAL and bad.value:
        raise IndexError("Target out of bounds")          # what torch's cross_entropy raises on the CPU
    _cabi.check(rc)
    return (loss, correct, total) if with_counts else loss
