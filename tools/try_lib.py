"""Guarded run of an alternative library build: python tools/try_lib.py LIB [frames ...] -- each case in its own process."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]
if sys.argv[1] == "--case":
    lib_path, n = sys.argv[2], int(sys.argv[3])
    from cutdet import _cabi
    if lib_path != "-":
        _cabi.LIB_OVERRIDE = os.path.abspath(lib_path)
    import torch
    from cutdet import engine, synth
    from frameID.net import load_default_net
    net, _ = load_default_net()
    native = net.eval().to("cuda")._native()
    plan = engine.ResizePlan.for_video(720, 1280, 256)
    frames = synth.SyntheticClip(720, 1280, n, seed=1).frames_torch(0, n, device="cuda")
    native.set_option("conv1_variant", 3)
    want = native.forward_frames(plan, frames).clone()
    native.set_option("conv1_variant", 2)
    for _ in range(3):
        got = native.forward_frames(plan, frames)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        native.forward_frames(plan, frames)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    print(f"{lib_path} frames {n} ok {torch.equal(got, want)} {ms:.3f} ms {n / ms / 1e3:.3f} M f/s", flush=True)
else:
    lib_path = sys.argv[1]
    for n in [int(a) for a in sys.argv[2:]] or [296, 4050]:
        try:
            r = subprocess.run([sys.executable, __file__, "--case", lib_path, str(n)], capture_output=True, text=True, timeout=60)
            print(r.stdout.strip() or ("rc %d %s" % (r.returncode, r.stderr[-400:])), flush=True)
        except subprocess.TimeoutExpired:
            print(lib_path, "frames", n, "HANG", flush=True)
            sys.exit(1)
