#!/usr/bin/env python3
"""The quad kernel of K1 (two-tap resizes) for ncu: 640x360 and 1920x1080 frames, float32 and uint8 outputs, two launches each.

    python tools/ncu_k1_quads.py && ncu --set full -k regex:preprocess_quads -c 8 python tools/ncu_k1_quads.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]


def main():
    import torch
    from cutdet import engine
    for h, w, batch in ((360, 640, 2368), (1080, 1920, 512)):
        plan = engine.ResizePlan.for_video(h, w, 256)
        g = torch.Generator(device="cuda").manual_seed(1)
        frames = torch.randint(0, 256, (batch, h, w, 3), dtype=torch.uint8, device="cuda", generator=g)
        for _ in range(2):
            engine.preprocess_f32(plan, frames)
            engine.preprocess_u8(plan, frames)
        torch.cuda.synchronize()
        print(f"{w}x{h} x{batch}: source rows {len(plan.rows) * 3 * w * batch / 1e6:.2f} MB, "
              f"f32 out {batch * 3 * plan.dst_h * plan.dst_w * 4 / 1e6:.2f} MB, u8 out {batch * 3 * plan.dst_h * plan.dst_w / 1e6:.2f} MB")
        del frames
    print("NCU_K1_QUADS_OK")


if __name__ == "__main__":
    main()
