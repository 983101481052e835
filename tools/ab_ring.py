"""Same-process A/B of conv12_frames' row-supply options: the operand-ring capacity where a resized row reads two source rows
(net option ring_cap: 2 = 768 positions and a larger raw-row ring, 1 = the full 1,024; 0 = the library's choice) and the loaders'
L2 prefetch of the source rows (src_prefetch: 0 = never, 1 = two-row geometries, 2 = every tensor map): ms per forward_frames
call, interleaved, and bit-equality.
    python tools/ab_ring.py [libcutdet_b200.so of another build]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]
from cutdet import _cabi
if len(sys.argv) > 1:                 # another build of the library (e.g. compiled with -DF1_RESIZE_ILP=8)
    _cabi.LIB_OVERRIDE = os.path.abspath(sys.argv[1])
    print("library:", sys.argv[1])
from cutdet import engine, synth
from frameID.net import load_default_net

nets = {}
VARIANTS = (("default", {}), ("prefetch_two_rows", {"src_prefetch": 1}), ("prefetch_all", {"src_prefetch": 2}), ("full_ring", {"ring_cap": 1}),
            ("full_ring_prefetch", {"ring_cap": 1, "src_prefetch": 1}))
for name, opts in VARIANTS:
    net, _ = load_default_net()
    nets[name] = net.eval().to("cuda")._native()
    for k, v in opts.items():
        nets[name].set_option(k, v)
for h, w, n in ((1080, 1920, 1184), (360, 640, 2368), (720, 1280, 1184)):
    frames = synth.SyntheticClip(h, w, n, seed=3).frames_torch(0, n, device="cuda")
    plan = engine.ResizePlan.for_video(h, w, 256)
    outs = {}
    for name, net in nets.items():
        for _ in range(2):
            outs[name] = net.forward_frames(plan, frames)
    torch.cuda.synchronize()
    print(f"{w}x{h} x{n}: logits bit-equal: {all(bool(torch.equal(outs['default'], o)) for o in outs.values())}", flush=True)
    for rep in range(3):
        for name, net in nets.items():
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                net.forward_frames(plan, frames)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            print(f"  rep {rep} {name:22s} {ms:.3f} ms per {n} frames = {n / ms * 1e3:,.0f} frames/s", flush=True)
    del frames
    torch.cuda.empty_cache()
