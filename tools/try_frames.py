"""First runs of the frame-by-frame conv1+conv2 kernel (net option conv1_variant=2) against the default path: bits and time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]
import numpy as np
import torch
from cutdet import engine, synth
from frameID.net import load_default_net

net, _ = load_default_net()
native = net.eval().to("cuda")._native()
for (h, w, n) in ((720, 1280, 148), (720, 1280, 1184), (720, 1280, 4050), (1080, 1920, 600), (360, 640, 700)):
    plan = engine.ResizePlan.for_video(h, w, 256)
    frames = synth.SyntheticClip(h, w, n, seed=1).frames_torch(0, n, device="cuda")
    rng = np.random.default_rng(n)
    frames[: min(n, 64)] = torch.from_numpy(rng.integers(0, 256, (min(n, 64), h, w, 3), dtype=np.uint8)).cuda()
    res = {}
    for variant in (3, 2):
        native.set_option("conv1_variant", variant)
        out = native.forward_frames(plan, frames).clone()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(3):
            native.forward_frames(plan, frames)
        ev[0].record()
        for _ in range(10):
            native.forward_frames(plan, frames)
        ev[1].record()
        torch.cuda.synchronize()
        res[variant] = (out, ev[0].elapsed_time(ev[1]) / 10)
    same = torch.equal(res[3][0], res[2][0])
    d = float((res[3][0] - res[2][0]).abs().max())
    print(f"{h}p n={n}: default {res[3][1]:.3f} ms ({n / res[3][1] / 1e3:.3f} M f/s)  frames {res[2][1]:.3f} ms ({n / res[2][1] / 1e3:.3f} M f/s)  "
          f"bit-equal {same} max|d| {d:.4g} nan {bool(torch.isnan(res[2][0]).any())}", flush=True)
