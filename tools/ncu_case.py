#!/usr/bin/env python3
"""One short run of every kernel the round's ncu captures look at (profiles/README.md): the standalone K1 kernels (both of
them, 720p and 1080p, float32 and uint8 outputs), K4/K5 over a full game's worth of frames (324,000), K6, and two full
sub-batches of the fused frames path, by the two-kernel path (conv1_fused_tc -> conv2_tc -> conv3_tc -> head) and by the default
one (conv12_frames -> conv3_tc -> head).

    python tools/ncu_case.py && ncu --set full -k regex:... python tools/ncu_case.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]


def main():
    import torch
    from cutdet import _cabi, engine, pipeline, synth
    from frameID.net import load_default_net
    net, _ = load_default_net()
    native = net.eval().to("cuda")._native()
    lib = _cabi.lib()
    for h, w, batch in ((720, 1280, 592), (1080, 1920, 296)):
        frames = synth.SyntheticClip(h, w, batch, seed=1).frames_torch(0, batch, device="cuda")
        plan = engine.ResizePlan.for_video(h, w, 256)
        for kernel in (1, 2):
            _cabi.check(lib.cutdet_debug_k1_kernel(kernel))
            for _ in range(2):
                engine.preprocess_f32(plan, frames)
                engine.preprocess_u8(plan, frames)
        _cabi.check(lib.cutdet_debug_k1_kernel(0))
        for variant in (3, 0):              # the two-kernel path, then conv12_frames (two frames per CTA)
            native.set_option("conv1_variant", variant)
            for _ in range(2):
                native.forward_frames(plan, frames[:296])
        torch.cuda.synchronize()
        del frames
    n = 324_000
    rng = np.random.default_rng(0)
    lab = np.repeat(rng.integers(0, 3, n // 400 + 2), 400)[:n]
    lab[rng.uniform(size=n) < 0.003] = 2
    scores = np.full((n, 3), -1.0, np.float32)
    scores[np.arange(n), lab] = rng.uniform(2, 9, n).astype(np.float32)
    dev = torch.from_numpy(scores).cuda()
    for _ in range(2):
        table = engine.run_table_from_scores(dev)
        pipeline.smooth(table, 100, 10)
        table.to_te()
    torch.cuda.synchronize()
    print("NCU_CASE_OK")


if __name__ == "__main__":
    main()
