"""Device-timed frames/s of the float-tensor entry net(x) (the reference's own call, segment_video.py:45): x float32 [B,3,H,W]
resident in HBM.     python tools/time_f32.py 144 256 1184"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]
from cutdet import _cabi
from frameID.net import load_default_net

h, w, batch = (int(a) for a in sys.argv[1:4])
net, _ = load_default_net()
net.eval().to("cuda:0")
x = torch.rand((batch, 3, h, w), device="cuda")
with torch.no_grad():
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    lib = _cabi.lib()
    lib.cutdet_profile_begin()
    net(x)
    torch.cuda.synchronize()
    import ctypes
    buf = ctypes.create_string_buffer(1 << 16)
    lib.cutdet_profile_end(buf, len(buf))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 10
    for _ in range(n):
        net(x)
    b.record()
    torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
print(f"net(x) {w}x{h} batch {batch}: {ms:.3f} ms per batch, {batch / ms * 1e3:,.0f} frames/s")
print(buf.value.decode()[:1500])
