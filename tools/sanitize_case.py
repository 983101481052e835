#!/usr/bin/env python3
"""The smallest run that exercises every hand-rolled synchronisation protocol of the library, for compute-sanitizer
(one tool per gpurun call: memcheck, racecheck, synccheck; see profiles/README.md):

  * forward_frames on a few 720p frames (integer-scale gather: bulk-copy raw ring -> unfold -> operand ring -> tcgen05 MMAs ->
    TMEM epilogue, conv2/conv3 with TMA stages, programmatic dependent launch between them) and a few 1080p frames (two source
    rows per output row, the fixed-point bilinear path), with the default kernels and with the conv1_teams experiment;
  * K4 argmax + a multi-block rle_append (decoupled look-back: blocks spin on their predecessor's flag) + rle_finish;
  * K6 glue_orphans / combine_adjacent, the shard pack + stitch kernels, K1 standalone (row kernel and generic kernel).
Results are checked against the oracle so that a tool that perturbs timing cannot hide a wrong answer."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cut-detection_b200")]


def main():
    import torch
    from cutdet import engine, pipeline, shard
    from frameID.net import load_default_net
    from oracle import net as onet, preprocess as opre, segmentation as oseg
    frames_per_case = int(os.environ.get("SANITIZE_FRAMES", "3"))
    net, params = load_default_net()
    native = net.eval().to("cuda")._native()
    weights, wparams = onet.load_weights_npz(os.path.join(ROOT, "cut-detection_b200", "frameID", "prod_net", "prod_net_weights.npz"))
    rng = np.random.default_rng(0)
    for teams in (0, 1):     # default fused conv1 kernel / the two-epilogue-set experiment
        native.set_option("conv1_variant", teams)
        for h, w in ((720, 1280), (1080, 1920)):
            frames = rng.integers(0, 256, (frames_per_case, h, w, 3), dtype=np.uint8)
            frames[0, : h // 2] = 255
            plan = engine.ResizePlan.for_video(h, w, 256)
            got = native.forward_frames(plan, torch.from_numpy(frames).cuda()).cpu().numpy()
            want = onet.forward_f32(weights, opre.preprocess_batch(frames, 256), wparams["avg_pool_size"])
            err = float(np.abs(got - want).max())
            print(f"forward_frames {w}x{h} x{frames_per_case} teams={teams}: max|dlogit| {err:.4f}")
            assert err <= 0.05
            x = engine.preprocess_f32(plan, torch.from_numpy(frames).cuda())
            assert np.array_equal(x.cpu().numpy(), opre.preprocess_batch(frames, 256))
    native.set_option("conv1_variant", 0)
    n = 20_011
    lab = np.repeat(rng.integers(0, 3, n // 37 + 2), 37)[:n].astype(np.uint8)
    lab[rng.uniform(size=n) < 0.02] = 2
    top = rng.uniform(1, 9, n).astype(np.float32)
    scores = np.full((n, 3), -1.0, np.float32)
    scores[np.arange(n), lab] = top
    table = engine.run_table_from_scores(torch.from_numpy(scores).cuda())
    want0 = oseg.run_table(scores)
    te = table.to_te()
    assert np.array_equal(te["end_frames"].numpy(), want0["end_frames"]) and np.array_equal(te["frame_types"].numpy(), want0["frame_types"])
    halves = []
    for lo, hi in ((0, 9_000), (9_000, n)):
        enc = engine.RunLengthEncoder(hi - lo, "cuda")
        enc.append(torch.from_numpy(lab[lo:hi]).cuda(), torch.from_numpy(top[lo:hi]).cuda())
        halves.append(enc.finish())
    joined, total = shard.stitch_local(halves, [9_000, n - 9_000], 1024)
    pipeline.smooth(joined, 100, 10)
    got = joined.to_te()
    want = oseg.combine_adjacent(oseg.glue_orphans(want0, 100, 10))
    for k in ("end_frames", "frame_types", "run_lengths", "start_frames"):
        assert np.array_equal(got[k].numpy(), want[k]), k
    print(f"segmentation: {len(want0['end_frames'])} runs -> {len(want['end_frames'])} segments, stitched + smoothed == oracle")
    torch.cuda.synchronize()
    print("SANITIZE_CASE_OK")


if __name__ == "__main__":
    main()
