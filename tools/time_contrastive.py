"""BASELINE config 5: forward pass of the contrastive encoder (FrameConvNet(32 ch, 3 layers) + FrameLinearNet(3, 32, 32, 8)) on
synthetic frame pairs, device-timed: eval() (running statistics, tensor-core path) and training mode (batch statistics, as the
reference's learn_contrasts.py runs it; float32 CUDA-core path), plus ContrastiveLoss.
    python tools/time_contrastive.py 512        # pairs per batch"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]
from frameID.net import FrameConvNet, FrameLinearNet
from frameID.metrics import ContrastiveLoss

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
conv_net = FrameConvNet(hidden_channels=32, n_conv_layers=3).to("cuda")
linear_net = FrameLinearNet(n_layers=3, input_size=32, hidden_size=32, output_size=8).to("cuda")
crit = ContrastiveLoss(batch_size=pairs).to("cuda")
x = torch.rand((2 * pairs, 3, 144, 256), device="cuda")


def run():
    return crit(linear_net(conv_net(x)))[0]


for mode in ("train", "eval"):
    conv_net.train(mode == "train"); linear_net.train(mode == "train")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 10
    for _ in range(n):
        loss = run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print(f"contrastive forward, {pairs} pairs ({2 * pairs} frames 256x144), BatchNorm in {mode} mode: {ms:.3f} ms, "
          f"{2 * pairs / ms * 1e3:,.0f} frames/s, loss {float(loss):.4f}")
