"""The standalone K1 kernels side by side in ONE process (profiles/r02_k1_matrix*.txt): every kernel (1 = one thread per pixel,
2 = staged rows, 3 = four adjacent pixels per thread, 0 = the library's choice) x geometry x output type, frames resident in HBM,
a warm loop of 10 launches timed with CUDA events; bytes = the source rows the resize reads + the output tensor.
    python tools/k1_matrix.py [kernels, e.g. 0123]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]
from cutdet import _cabi, engine

kernels = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "1230")]
peak = 6453.1
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
GEOMS = [(720, 1280, 1184), (1080, 1920, 512), (360, 640, 2368), (480, 854, 1184), (288, 512, 2368), (2160, 3840, 148)]
lib = _cabi.lib()
for h, w, batch in GEOMS:
    plan = engine.ResizePlan.for_video(h, w, 256)
    g = torch.Generator(device="cuda").manual_seed(1)
    frames = torch.randint(0, 256, (batch, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
    n_rows = len(plan.rows)
    for name, fn, out_bytes in (("f32 NCHW", engine.preprocess_f32, 3 * plan.dst_h * plan.dst_w * 4),
                                ("u8 HWC", engine.preprocess_u8, 3 * plan.dst_h * plan.dst_w)):
        cells = []
        for kernel in kernels:
            _cabi.check(lib.cutdet_debug_k1_kernel(kernel))
            for _ in range(3):
                fn(plan, frames)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                fn(plan, frames)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 10
            bytes_per_frame = n_rows * 3 * w + out_bytes
            gbs = batch * bytes_per_frame / ms / 1e6
            cells.append(f"k{kernel}: {ms:.3f} ms {gbs:,.0f} GB/s {100 * gbs / peak:.1f} %")
        print(f"{w}x{h} x{batch} -> {plan.dst_w}x{plan.dst_h} {name:8s} ({n_rows} rows + out = {bytes_per_frame:,} B/frame)  " + "   ".join(cells), flush=True)
    _cabi.check(lib.cutdet_debug_k1_kernel(0))
    del frames
    torch.cuda.empty_cache()
