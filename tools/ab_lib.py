"""Same-box A/B of two builds of the library on the frames path: every (library, geometry) case in its own process (a kernel that
hangs costs its timeout, not the call), ms per forward_frames call and a SHA-256 of the logits, so that builds can be compared
for speed and for bits.     python tools/ab_lib.py - build/libcutdet_other.so        ("-" = the in-tree build)"""
import hashlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]
CASES = ((720, 1280, 4050), (720, 1280, 1184), (1080, 1920, 1184))
if sys.argv[1] == "--case":
    lib_path, h, w, n = sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    from cutdet import _cabi
    if lib_path != "-":
        _cabi.LIB_OVERRIDE = os.path.abspath(lib_path)
    import torch
    from cutdet import engine, synth
    from frameID.net import load_default_net
    net, _ = load_default_net()
    native = net.eval().to("cuda")._native()
    plan = engine.ResizePlan.for_video(h, w, 256)
    frames = synth.SyntheticClip(h, w, n, seed=1).frames_torch(0, n, device="cuda")
    for _ in range(3):
        out = native.forward_frames(plan, frames)
    torch.cuda.synchronize()
    digest = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
    times = []
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            native.forward_frames(plan, frames)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b) / 10)
    print(f"{lib_path:32s} {w}x{h} x{n}: " + " ".join(f"{t:.3f}" for t in times) + f" ms  best {n / min(times) / 1e3:.3f} M frames/s  logits {digest}", flush=True)
else:
    for h, w, n in CASES:
        for lib_path in sys.argv[1:] * 2:          # each library twice, interleaved
            try:
                r = subprocess.run([sys.executable, __file__, "--case", lib_path, str(h), str(w), str(n)], capture_output=True, text=True, timeout=90)
                print(r.stdout.strip() or ("rc %d %s" % (r.returncode, r.stderr[-300:])), flush=True)
            except subprocess.TimeoutExpired:
                print(lib_path, h, w, n, "HANG", flush=True)
