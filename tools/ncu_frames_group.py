#!/usr/bin/env python3
"""A FULL launch of the default frames path for ncu: 1,184 frames of 720p = eight frames per persistent CTA, the shape the
bench launches (tools/ncu_case.py captures 296 frames = two per CTA, where ramp and drain weigh four times as much).

    python tools/ncu_frames_group.py && ncu --set full -k regex:conv12_frames -c 2 python tools/ncu_frames_group.py
    python tools/ncu_frames_group.py 1184 1080 1920         (frames, height, width: another geometry)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]


def main():
    import torch
    from cutdet import engine, synth
    from frameID.net import load_default_net
    net, _ = load_default_net()
    native = net.eval().to("cuda")._native()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
    h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (720, 1280)
    frames = synth.SyntheticClip(h, w, n, seed=1).frames_torch(0, n, device="cuda")
    plan = engine.ResizePlan.for_video(h, w, 256)
    for _ in range(3):
        logits = native.forward_frames(plan, frames)
    torch.cuda.synchronize()
    print("NCU_FRAMES_GROUP_OK", tuple(logits.shape))


if __name__ == "__main__":
    main()
