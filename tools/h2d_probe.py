"""What the host link gives: pinned host -> device copies of one 4,050-frame chunk's source rows (2.24 GB), timed with CUDA events.
(a) one contiguous copy, (b) cutdet_upload_frames (strided 2-D DMA: 3,840-byte rows at a 19,200-byte pitch), (c) the same split
over two streams.     python tools/h2d_probe.py"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]
from cutdet import _cabi, engine

n, h, w = 4050, 720, 1280
plan = engine.ResizePlan.for_video(h, w, 256)
rows = len(plan.rows)
host = torch.empty((n, h, w, 3), dtype=torch.uint8).pin_memory()
host.view(-1)[::4096] = 1
flat = torch.empty(n * rows * w * 3, dtype=torch.uint8).pin_memory()
dev = torch.empty((n, rows, w, 3), dtype=torch.uint8, device="cuda")
lib = _cabi.lib()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def contiguous():
    dev.view(-1).copy_(flat, non_blocking=True)


def strided(frames=host, out=dev, stream=None):
    st = stream or torch.cuda.current_stream()
    _cabi.check(lib.cutdet_upload_frames(plan.handle, frames.data_ptr(), frames.shape[0], frames.stride(0), frames.stride(1),
                                         out.data_ptr(), st.cuda_stream, None))


def two_streams():
    cur = torch.cuda.current_stream()
    e = torch.cuda.Event(); e.record(cur)
    for st, lo, hi in ((s1, 0, n // 2), (s2, n // 2, n)):
        st.wait_event(e)
        strided(host[lo:hi], dev[lo:hi], st)
        d = torch.cuda.Event(); d.record(st); cur.wait_event(d)


gb = n * rows * w * 3 / 1e9
for name, fn in (("contiguous", contiguous), ("strided 2-D (cutdet_upload_frames)", strided), ("strided, two streams", two_streams)):
    ms = timed(fn)
    print(f"{name:38s} {ms:8.2f} ms  {gb / ms * 1e3:6.2f} GB/s", flush=True)

# ---- the same uploads with the kernels of the previous chunk running beside them (FramePipeline.push_host), per step
from cutdet import pipeline
from frameID.net import load_default_net
net, _ = load_default_net()
native = net.eval().to("cuda")._native()
del dev, flat
for lanes in (1, 2):
    pipe = pipeline.FramePipeline(native, plan, n, 16 * n, "cuda", lanes=lanes)
    for _ in range(2):
        pipe.push_host(host)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(8):
        pipe.push_host(host)
    pipe.finish()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 8
    print(f"push_host pipeline, lanes={lanes}:          {ms:8.2f} ms per chunk  {gb / ms * 1e3:6.2f} GB/s", flush=True)
    del pipe
# uploads alone through the pipeline's copy stream and staging buffers, no kernels
stage = [torch.empty((n, rows, w, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
cs = torch.cuda.Stream()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(cs)
for i in range(8):
    strided(host, stage[i & 1], cs)
b.record(cs); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 8
print(f"uploads alone, two staging buffers:        {ms:8.2f} ms per chunk  {gb / ms * 1e3:6.2f} GB/s", flush=True)
