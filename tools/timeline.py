#!/usr/bin/env python3
"""Clock stamps of CTA 0 of conv1_fused_tc (kernel 1) or conv2_tc (kernel 2): cutdet_net_debug_timeline arms a caller-owned
device buffer, one full sub-batch is run, the stamps come back relative to the kernel's first stamp.

    python tools/timeline.py --kernel 1 --out gpurun_out/timeline1.txt [--height 720 --width 1280]

Stamp indices are documented next to the `tl[...]` stores in csrc/conv_tc.cu."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]


def main():
    import torch
    from cutdet import _cabi, engine, synth
    from frameID.net import load_default_net
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", type=int, default=1)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--frames", type=int, default=444, help="for --kernel 3 use a multiple of 148, e.g. 1184")
    ap.add_argument("--out", default="gpurun_out/timeline.txt")
    ap.add_argument("--net-opt", action="append", default=[])
    a = ap.parse_args()
    net, _ = load_default_net()
    native = net.eval().to("cuda")._native()
    native.set_option("conv1_variant", 2 if a.kernel == 3 else 3)       # kernels 1 and 2 belong to the two-kernel path
    for kv in a.net_opt:
        k, _, v = kv.partition("=")
        native.set_option(k, int(v))
    plan = engine.ResizePlan.for_video(a.height, a.width, 256)
    frames = synth.SyntheticClip(a.height, a.width, a.frames, seed=1).frames_torch(0, a.frames, device="cuda")
    for _ in range(3):
        native.forward_frames(plan, frames)
    torch.cuda.synchronize()
    stamps = torch.zeros(4096, dtype=torch.int64, device="cuda")
    _cabi.check(_cabi.lib().cutdet_net_debug_timeline(native.handle, a.kernel, stamps.data_ptr(), 4096))
    native.forward_frames(plan, frames)
    torch.cuda.synchronize()
    _cabi.check(_cabi.lib().cutdet_net_debug_timeline(native.handle, 0, None, 0))
    h = stamps.cpu().tolist()
    t0 = h[2047] if a.kernel == 1 else h[0]
    if a.kernel == 3:
        # conv12_frames_kernel, CTA 0: [0] start, then per frame [1 + 3 it] layer 1 set up, [2 + 3 it] layer 1 done, [3 + 3 it] layer 2 done
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            f.write("# frame: setup_done layer1_done layer2_done (cycles since CTA 0's start) | layer1 layer2 cycles\n")
            it = 0
            prev = 0
            while 3 + 3 * it < 2048 and h[3 + 3 * it]:
                b, m, e = (h[1 + 3 * it] - t0, h[2 + 3 * it] - t0, h[3 + 3 * it] - t0)
                f.write(f"{it}: {b} {m} {e} | setup {b - prev} layer1 {m - b} layer2 {e - m}\n")
                prev = e
                it += 1
            if it > 2 and h[700]:
                # the third frame in detail (cycles since its layer 1 was set up)
                f0 = h[1 + 3 * 2]
                f.write("# third frame, layer 1: tile: rows_there issued stored | period\n")
                t = 0
                while t < 40 and h[700 + 2 * t]:
                    f.write(f"tile {t}: {h[700 + 2 * t] - f0} {h[701 + 2 * t] - f0} {h[800 + t] - f0} | {h[700 + 2 * t] - h[700 + 2 * (t - 1)] if t else 0}\n")
                    t += 1
                f.write("# loaders: row: issued (after its slot was free)\n")
                f.write(" ".join(f"{n}:{h[40 + n] - f0}" for n in range(0, 200) if h[40 + n]) + "\n")
                f.write("# unfold warp 0: its k-th row: start, wait for ring space and source row, work, fence\n")
                k = 0
                while k < 30 and h[256 + 4 * k]:
                    a0, b0, c0, d0 = (h[256 + 4 * k + j] for j in range(4))
                    f.write(f"urow {k}: start {a0 - f0} wait {b0 - a0} work {c0 - b0} fence {d0 - c0}\n")
                    k += 1
            g0 = min(v for v in h[2048::2] if v)
            for b in range(148):
                if h[2048 + 2 * b]:
                    f.write(f"cta {b} {h[2048 + 2 * b] - g0} {h[2049 + 2 * b] - g0}\n")
        print(f"wrote {it} frames to {a.out}")
        return
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        for i, v in enumerate(h[:2048]):
            if v:
                f.write(f"{i} {v - t0}\n")
        if a.kernel == 1:          # per-CTA entry / exit in ns since the first CTA's entry
            g0 = min(v for v in h[2048::2] if v)
            for b in range(148):
                if h[2048 + 2 * b]:
                    f.write(f"cta {b} {h[2048 + 2 * b] - g0} {h[2049 + 2 * b] - g0}\n")
    print(f"wrote {sum(1 for v in h if v)} stamps to {a.out}")


if __name__ == "__main__":
    main()
