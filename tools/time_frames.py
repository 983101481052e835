"""Device-timed frames/s of the fused path (K1 + CNN) for one resolution and batch size, frames resident in HBM.
    python tools/time_frames.py 1080 1920 1184"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]
from cutdet import engine
from frameID.net import load_default_net

h, w, batch = (int(a) for a in sys.argv[1:4])
net, _ = load_default_net()
net.eval().to("cuda:0")
plan = engine.ResizePlan.for_video(h, w, 256)
g = torch.Generator(device="cuda").manual_seed(1)
frames = torch.randint(0, 256, (batch, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
with torch.no_grad():
    for _ in range(3):
        net.forward_frames(plan, frames)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 10
    for _ in range(n):
        net.forward_frames(plan, frames)
    b.record()
    torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
print(f"{w}x{h} batch {batch}: {ms:.3f} ms per batch, {batch / ms * 1e3:,.0f} frames/s (frames resident in HBM: {frames.numel() / 1e9:.2f} GB)")
