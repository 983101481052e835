"""Guarded first run of a kernel change: each case in its own process with a short timeout, so a hang costs seconds."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]
if len(sys.argv) > 1:
    import torch
    from cutdet import engine, synth
    from frameID.net import load_default_net
    n = int(sys.argv[1])
    net, _ = load_default_net()
    native = net.eval().to("cuda")._native()
    plan = engine.ResizePlan.for_video(720, 1280, 256)
    frames = synth.SyntheticClip(720, 1280, n, seed=1).frames_torch(0, n, device="cuda")
    native.set_option("conv1_variant", 3)
    want = native.forward_frames(plan, frames).clone()
    native.set_option("conv1_variant", 2)
    for _ in range(5):
        got = native.forward_frames(plan, frames)
    torch.cuda.synchronize()
    print("frames", n, "ok", torch.equal(got, want), flush=True)
else:
    for n in (296, 1000):
        try:
            r = subprocess.run([sys.executable, __file__, str(n)], capture_output=True, text=True, timeout=45)
            print(r.stdout.strip() or ("n %d rc %d %s" % (n, r.returncode, r.stderr[-300:])), flush=True)
        except subprocess.TimeoutExpired:
            print("frames", n, "HANG", flush=True)
            sys.exit(1)
