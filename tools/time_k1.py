"""Device-timed throughput of the STANDALONE preprocessing kernel K1 (cutdet_preprocess_f32 / _u8), frames resident in HBM,
against its HBM roofline: bytes = the source rows the resize reads + the output tensor.
    python tools/time_k1.py 720 1280 1184 [kernel]        kernel: 0 = the library's choice, 1 = one thread per pixel, 2 = row kernel"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]
from cutdet import _cabi, engine

h, w, batch = (int(a) for a in sys.argv[1:4])
kernel = int(sys.argv[4]) if len(sys.argv) > 4 else 0
_cabi.check(_cabi.lib().cutdet_debug_k1_kernel(kernel))
plan = engine.ResizePlan.for_video(h, w, 256)
g = torch.Generator(device="cuda").manual_seed(1)
frames = torch.randint(0, 256, (batch, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
peak = 6453.1
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
n_rows = len(plan.rows)
for name, fn, out_bytes in (("f32 NCHW", engine.preprocess_f32, 3 * plan.dst_h * plan.dst_w * 4),
                            ("u8 HWC", engine.preprocess_u8, 3 * plan.dst_h * plan.dst_w)):
    for _ in range(3):
        fn(plan, frames)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 10
    for _ in range(n):
        fn(plan, frames)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    bytes_per_frame = n_rows * 3 * w + out_bytes
    gbs = batch * bytes_per_frame / ms / 1e6
    print(f"K1[kernel {kernel}] {w}x{h} -> {plan.dst_w}x{plan.dst_h} {name}: {ms:.3f} ms per {batch} frames, {batch / ms * 1e3:,.0f} frames/s, "
          f"{bytes_per_frame:,} B/frame ({n_rows} source rows + output) -> {gbs:,.0f} GB/s = {100 * gbs / peak:.1f} % of {peak:.0f} GB/s")
