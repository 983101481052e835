"""frameID.net -- the reference's network API (reference frameID/net.py) over the B200 kernels.

Same public names and constructor arguments as the reference:

    CNNLayer, FCLayer                       net.py:11-68    (parameter containers here)
    FrameConvNet(input_channels=3, hidden_channels=32, n_conv_layers=3, average_pool_size=1)     net.py:71-136
    FrameLinearNet(n_layers=3, input_size=32, hidden_size=32, output_size=8)                     net.py:139-189
    load_and_glue_nets(param_file, conv_file, linear_file) -> (net, params)                      net.py:193-217
    load_default_net() -> (net, params)                                                          net.py:221-233

The modules hold ordinary torch parameters under the reference's state_dict keys, so the reference's
checkpoint files load with ``load_state_dict`` unchanged, ``.eval()``, ``.to(device)`` and ``num_params()``
behave as before -- but ``forward`` does not run PyTorch operators: it hands the input to libcutdet_b200.so
(hand-written sm_100a kernels) through the C ABI.  Forward passes only (no autograd): after ``.eval()`` BatchNorm uses
its running statistics; a module left in training mode, as in the reference's learn_contrasts.py, normalises with the
statistics of the batch (float32 CUDA-core kernels) without updating the running ones.  A CPU input raises -- there is
no fallback path.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch
import torch.nn as nn

from cutdet import engine as _engine

package_directory = os.path.dirname(os.path.abspath(__file__))


def _weights_of(module: nn.Module, prefix: str) -> dict:
    return {prefix + k: v.detach().to("cpu", torch.float32).numpy()
            for k, v in module.state_dict().items() if not k.endswith("num_batches_tracked")}


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


class _LayerBacked(nn.Module):
    """A single layer run on its own: a one-layer native net, rebuilt when the parameters change."""

    def _layer_weights(self):   # -> weights dict in NativeNet's key layout
        raise NotImplementedError

    def _native_layer(self) -> _engine.NativeNet:
        fp = tuple((id(p), p._version, p.data_ptr()) for p in list(self.parameters()) + list(self.buffers()))
        cache = self.__dict__.get("_native_cache")
        if cache is None or cache[0] != fp:
            cache = (fp, _engine.NativeNet(self._layer_weights(), 1))
            self.__dict__["_native_cache"] = cache
        return cache[1]

    def _bn_mode(self) -> int:
        if not self.batch_norm:
            return 0
        return 2 if self.training else 1


class CNNLayer(_LayerBacked):
    """conv3x3 -> activation -> max-pool -> batch-norm (in that order; reference net.py:33-40).  Inside a FrameConvNet the whole
    trunk is launched by the enclosing module (tensor-core kernels); called on its own, the layer runs the float32 kernel
    ``cutdet_net_forward_conv_layer`` -- for the one shape of layer the reference's networks are made of: k3 / stride 1 /
    padding 1 convolution, ReLU, MaxPool2d(3).  Other arguments are accepted by the constructor (the parameters load and save
    as in the reference) and refused by ``forward``."""

    def __init__(self, conv_args: dict, max_pool_args: dict, activation=nn.ReLU, batch_norm=True):
        super().__init__()
        self.batch_norm = batch_norm
        self.conv = nn.Conv2d(**conv_args)
        self.activation = activation()
        self.max_pool = nn.MaxPool2d(**max_pool_args)
        self.bn = nn.BatchNorm2d(conv_args["out_channels"]) if batch_norm else nn.Identity()

    def _check_shape(self):
        c, p = self.conv, self.max_pool
        ok = (_pair(c.kernel_size) == (3, 3) and _pair(c.stride) == (1, 1) and _pair(c.padding) == (1, 1) and
              _pair(c.dilation) == (1, 1) and c.groups == 1 and c.bias is not None and c.padding_mode == "zeros" and
              _pair(p.kernel_size) == (3, 3) and _pair(p.stride if p.stride is not None else p.kernel_size) == (3, 3) and
              _pair(p.padding) == (0, 0) and _pair(p.dilation) == (1, 1) and not p.ceil_mode and
              isinstance(self.activation, nn.ReLU))
        if not ok:
            raise NotImplementedError("the native kernels implement conv3x3(stride 1, padding 1) -> ReLU -> MaxPool2d(3) -> "
                                      "BatchNorm2d, the only layer shape FrameConvNet builds (reference net.py:91-120)")

    def _layer_weights(self):
        w = {"conv.conv_layers.0.conv.weight": self.conv.weight, "conv.conv_layers.0.conv.bias": self.conv.bias}
        n = self.conv.out_channels
        if self.batch_norm:
            w.update({"conv.conv_layers.0.bn.weight": self.bn.weight, "conv.conv_layers.0.bn.bias": self.bn.bias,
                      "conv.conv_layers.0.bn.running_mean": self.bn.running_mean,
                      "conv.conv_layers.0.bn.running_var": self.bn.running_var})
        else:
            w.update({"conv.conv_layers.0.bn.weight": torch.ones(n), "conv.conv_layers.0.bn.bias": torch.zeros(n),
                      "conv.conv_layers.0.bn.running_mean": torch.zeros(n), "conv.conv_layers.0.bn.running_var": torch.ones(n)})
        return {k: v.detach().to("cpu", torch.float32).numpy() for k, v in w.items()}

    def forward(self, x):
        """float32 [B, Cin, H, W] on the GPU -> [B, Cout, H//3, W//3]."""
        self._check_shape()
        return self._native_layer().forward_conv_layer(0, x, self._bn_mode())


class FCLayer(_LayerBacked):
    """linear -> activation -> batch-norm (reference net.py:62-68).  Inside a FrameLinearNet the enclosing module launches the
    head; called on its own the layer runs ``cutdet_net_forward_fc_layer`` (activation ReLU or Identity)."""

    def __init__(self, linear_args: dict, activation=nn.ReLU, batch_norm=True):
        super().__init__()
        self.batch_norm = batch_norm
        self.linear = nn.Linear(**linear_args)
        self.activation = activation()
        self.bn = nn.BatchNorm1d(linear_args["out_features"]) if batch_norm else nn.Identity()

    def _layer_weights(self):
        if self.linear.bias is None:
            raise NotImplementedError("the native FC kernel expects a bias (nn.Linear default, as FrameLinearNet builds it)")
        w = {"linear.layers.0.linear.weight": self.linear.weight, "linear.layers.0.linear.bias": self.linear.bias}
        if self.batch_norm:
            w.update({"linear.layers.0.bn.weight": self.bn.weight, "linear.layers.0.bn.bias": self.bn.bias,
                      "linear.layers.0.bn.running_mean": self.bn.running_mean,
                      "linear.layers.0.bn.running_var": self.bn.running_var})
        return {k: v.detach().to("cpu", torch.float32).numpy() for k, v in w.items()}

    def forward(self, x):
        """float32 [B, in_features] on the GPU -> [B, out_features]."""
        if isinstance(self.activation, nn.ReLU):
            relu = True
        elif isinstance(self.activation, nn.Identity):
            relu = False
        else:
            raise NotImplementedError("the native FC kernel implements ReLU and Identity activations (reference net.py:164-178)")
        return self._native_layer().forward_fc_layer(0, x, relu, self._bn_mode())


class _NativeBacked(nn.Module):
    """Builds (and caches) the native net for the module's current parameters."""

    # options applied to the native net whenever it is (re)built: see engine.NativeNet.set_option
    native_options: dict = {}

    # training-mode forward: tcgen05 kernels + batch-statistics kernels where the tensor-core path applies (batches of up to 148
    # frames); False keeps the float32 CUDA-core kernels everywhere
    batchstats_tensor_cores = True

    def _native_parts(self):   # -> (weights dict, avg_pool_size)
        raise NotImplementedError

    def _fingerprint(self):
        return tuple((id(p), p._version, p.data_ptr()) for p in list(self.parameters()) + list(self.buffers()))

    def _native(self) -> _engine.NativeNet:
        fp = self._fingerprint()
        cache = self.__dict__.get("_native_cache")
        if cache is None or cache[0] != fp:
            weights, pool = self._native_parts()
            native = _engine.NativeNet(weights, pool)
            for name, value in self.native_options.items():
                native.set_option(name, value)
            cache = (fp, native)
            self.__dict__["_native_cache"] = cache
        return cache[1]

    def _run(self, x):
        """eval(): BatchNorm running statistics, tensor-core path where it applies.  A module left in training mode (the
        reference's learn_contrasts.py never calls .eval()) normalises with the statistics of the batch: forward only --
        the running statistics are not updated and no autograd graph is built."""
        if self.training:
            return self._native().forward_f32_batchstats(x, tensor_cores=self.batchstats_tensor_cores)
        return self._native().forward_f32(x)

    def set_native_option(self, name: str, value: int) -> None:
        """E.g. ``net.set_native_option("conv1_acc32", 1)``: layer 1 of the fused frames kernel accumulates in fp32."""
        self.native_options = dict(self.native_options, **{name: int(value)})
        self.__dict__.pop("_native_cache", None)

    def num_params(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)


class FrameConvNet(_NativeBacked):
    """The convolutional trunk: n_conv_layers x CNNLayer(k3 p1 conv, k3 pool, BN) -> AdaptiveAvgPool2d -> flatten."""

    def __init__(self, input_channels=3, hidden_channels=32, n_conv_layers=3, average_pool_size=1):
        super().__init__()
        self.input_channels = input_channels
        self.hidden_channels = hidden_channels
        self.n_conv_layers = n_conv_layers
        self.average_pool_size = average_pool_size
        self.conv_layers = nn.ModuleList()
        self.average_pool = nn.AdaptiveAvgPool2d(average_pool_size)
        channels = [input_channels] + [hidden_channels] * n_conv_layers
        for cin, cout in zip(channels[:-1], channels[1:]):
            self.conv_layers.append(CNNLayer(
                conv_args={"in_channels": cin, "out_channels": cout, "kernel_size": 3, "padding": 1},
                max_pool_args={"kernel_size": 3}, activation=nn.ReLU, batch_norm=True))

    def _native_parts(self):
        return _weights_of(self, "conv."), self.average_pool_size

    def forward(self, x):
        """float32 [B, Cin, H, W] on the GPU -> [B, hidden_channels * average_pool_size^2]."""
        return self._run(x)


class FrameLinearNet(_NativeBacked):
    """The fully connected head: Linear -> ReLU -> BatchNorm1d on all but the last layer, which is Linear only."""

    def __init__(self, n_layers: int = 3, input_size: int = 32, hidden_size: int = 32, output_size: int = 8):
        super().__init__()
        self.n_layers = n_layers
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.output_size = output_size
        self.layers = nn.ModuleList()
        widths = [input_size] + [hidden_size] * (n_layers - 1) + [output_size]
        for j, (fin, fout) in enumerate(zip(widths[:-1], widths[1:])):
            last = j == n_layers - 1
            self.layers.append(FCLayer(linear_args={"in_features": fin, "out_features": fout},
                                       activation=nn.Identity if last else nn.ReLU, batch_norm=not last))

    def _native_parts(self):
        return _weights_of(self, "linear."), 1

    def forward(self, x):
        """float32 [B, input_size] on the GPU -> [B, output_size] raw scores."""
        return self._run(x)


class GluedNet(nn.Sequential, _NativeBacked):
    """nn.Sequential(conv_net, linear_net) whose forward is ONE native pipeline (trunk and head fused),
    plus ``forward_frames`` which also fuses the frame preprocessing in front."""

    def __init__(self, conv_net: FrameConvNet, linear_net: FrameLinearNet):
        nn.Sequential.__init__(self, conv_net, linear_net)

    def _native_parts(self):
        weights = _weights_of(self[0], "conv.")
        weights.update(_weights_of(self[1], "linear."))
        return weights, self[0].average_pool_size

    def forward(self, x):
        """float32 [B, 3, H', W'] RGB in [0, 1] -> raw logits [B, output_size] (class order a22, ez, b)."""
        return self._run(x)

    def forward_frames(self, plan: _engine.ResizePlan, frames, compact: bool = False):
        """Addition to the reference API: decoded uint8 BGR HWC frames [B, h, w, 3] on the GPU -> logits."""
        return self._native().forward_frames(plan, frames, compact)


def _build(model_params: dict):
    conv_net = FrameConvNet(hidden_channels=model_params["conv_channels"], n_conv_layers=model_params["conv_layers"],
                            average_pool_size=model_params["avg_pool_size"])
    linear_net = FrameLinearNet(n_layers=model_params["linear_layers"],
                                input_size=model_params["conv_channels"] * model_params["avg_pool_size"] ** 2,
                                hidden_size=model_params["linear_size"], output_size=model_params["linear_output_size"])
    return conv_net, linear_net


def load_and_glue_nets(param_file, conv_file, linear_file):
    """Read the reference's three-file checkpoint (JSON + two torch state_dicts) into one callable."""
    with open(param_file, "r") as f:
        model_params = json.load(f)
    conv_net, linear_net = _build(model_params)
    conv_net.load_state_dict(torch.load(conv_file, map_location="cpu"))
    linear_net.load_state_dict(torch.load(linear_file, map_location="cpu"))
    return GluedNet(conv_net, linear_net), model_params


def load_torchscript_net(trace_file):
    """Addition to the reference API: take the parameters out of a TorchScript export of the
    glued net -- ``saved_model_trace.pt`` as written by the reference's training_scripts/make_torchscript_model.py:17-34
    (``torch.jit.trace(nn.Sequential(conv_net, linear_net), ...).save``) -- so that a deployment that ships only the
    traced file runs on the native path too.  The architecture is read off the tensor shapes: state-dict keys
    ``0.conv_layers.{i}.*`` / ``1.layers.{j}.*``, ``avg_pool_size = sqrt(linear_in / conv_channels)``.
    Returns ``(net, model_params)`` like ``load_and_glue_nets``."""
    scripted = torch.jit.load(trace_file, map_location="cpu")
    sd = {k: v.detach().clone() for k, v in scripted.state_dict().items()}
    conv = {k[2:]: v for k, v in sd.items() if k.startswith("0.")}
    linear = {k[2:]: v for k, v in sd.items() if k.startswith("1.")}
    if not conv or not linear or len(conv) + len(linear) != len(sd):
        raise ValueError(f"{trace_file}: not a traced nn.Sequential(FrameConvNet, FrameLinearNet)")
    n_conv = 1 + max(int(k.split(".")[1]) for k in conv if k.startswith("conv_layers."))
    n_lin = 1 + max(int(k.split(".")[1]) for k in linear if k.startswith("layers."))
    channels = int(conv["conv_layers.0.conv.weight"].shape[0])
    lin_in = int(linear["layers.0.linear.weight"].shape[1])
    pool = int(round((lin_in / channels) ** 0.5))
    if channels * pool * pool != lin_in:
        raise ValueError(f"{trace_file}: linear input {lin_in} is not conv_channels {channels} x a square pool size")
    model_params = {
        "conv_layers": n_conv, "conv_channels": channels, "avg_pool_size": pool, "linear_layers": n_lin,
        "linear_size": int(linear["layers.0.linear.weight"].shape[0]) if n_lin > 1 else 0,
        "linear_output_size": int(linear[f"layers.{n_lin - 1}.linear.weight"].shape[0]),
    }
    conv_net, linear_net = _build(model_params)
    conv_net.load_state_dict(conv)          # strict: a file of another architecture fails loudly, as in the reference
    linear_net.load_state_dict(linear)
    return GluedNet(conv_net, linear_net), model_params


def _load_npz(path):
    with np.load(path, allow_pickle=False) as z:
        model_params = json.loads(bytes(z["__params_json__"]).decode("utf-8"))
        arrays = {k: torch.from_numpy(z[k].copy()) for k in z.files if k != "__params_json__"}
    conv_net, linear_net = _build(model_params)
    conv_net.load_state_dict({k[len("conv."):]: v for k, v in arrays.items() if k.startswith("conv.")}, strict=False)
    linear_net.load_state_dict({k[len("linear."):]: v for k, v in arrays.items() if k.startswith("linear.")}, strict=False)
    return GluedNet(conv_net, linear_net), model_params


def load_default_net():
    """The classifier shipped with the package (prod_net/).  Uses the reference's own three files when they have been
    dropped into prod_net/, else the re-encoded copy of the same parameters (prod_net_weights.npz)."""
    d = os.path.join(package_directory, "prod_net")
    three = [os.path.join(d, n) for n in ("init_model_model_params.json", "init_model_classifier_conv.pt",
                                          "init_model_classifier_linear.pt")]
    if all(os.path.isfile(p) for p in three):
        return load_and_glue_nets(*three)
    return _load_npz(os.path.join(d, "prod_net_weights.npz"))
