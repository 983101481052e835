"""Repetition tests of the hand-rolled synchronisation protocols (compute-sanitizer is closed on this pool -- see
profiles/README.md, round 2 -- so races are hunted the way they show: as results that change from run to run).

conv1_fused_tc joins loaders, unfold warps, an MMA issuer and epilogue warps with monotone counters, mbarrier rings and
programmatic dependent launch; conv_mid_tc has a TMA producer, an MMA issuer and epilogue warps; rle_append chains blocks with a
decoupled look-back.  Every case runs many times -- alone, back to back without host synchronisation, and on two streams at
once so that two instances of the kernels share the SMs -- and every run must reproduce the first one's bits."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [0, 1, 3])
def test_forward_frames_is_reproducible(prod_weights, variant):
    from cutdet import engine
    wts, params = prod_weights
    nets = [engine.NativeNet(wts, params["avg_pool_size"]) for _ in range(2)]
    for n in nets:
        n.set_option("conv1_variant", variant)
    rng = np.random.default_rng(variant)
    for h, w, batch in ((720, 1280, 460), (1080, 1920, 190), (360, 640, 333)):
        frames = torch.from_numpy(rng.integers(0, 256, (batch, h, w, 3), dtype=np.uint8)).cuda()
        plan = engine.ResizePlan.for_video(h, w, 256)
        want = nets[0].forward_frames(plan, frames).clone()
        torch.cuda.synchronize()
        # back to back on one stream, no host synchronisation in between
        outs = [nets[0].forward_frames(plan, frames, out=torch.empty_like(want)) for _ in range(12)]
        torch.cuda.synchronize()
        for o in outs:
            assert torch.equal(o, want), (h, "sequential")
        # two nets on two streams at once: the kernels of both compete for the SMs (each net owns its workspace)
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        outs = []
        for rep in range(6):
            for n, st in zip(nets, streams):
                with torch.cuda.stream(st):
                    outs.append(n.forward_frames(plan, frames, out=torch.empty_like(want)))
        torch.cuda.synchronize()
        for o in outs:
            assert torch.equal(o, want), (h, "two streams")


def test_run_length_encoding_is_reproducible():
    from cutdet import engine
    n = 300_007                                    # 147 blocks per launch: a long look-back chain
    rng = np.random.default_rng(3)
    lab = np.repeat(rng.integers(0, 3, n // 11 + 2), 11)[:n].astype(np.uint8)
    lab[rng.uniform(size=n) < 0.05] = 1
    top = rng.uniform(1, 9, n).astype(np.float32)
    dl, dt = torch.from_numpy(lab).cuda(), torch.from_numpy(top).cuda()
    first = None
    for rep in range(25):
        enc = engine.RunLengthEncoder(n, "cuda")
        pos = 0
        for c in (100_000, 1, 150_006, 50_000):   # several launches carry the open run across
            enc.append(dl[pos:pos + c], dt[pos:pos + c])
            pos += c
        te = enc.finish().to_te()
        if first is None:
            first = te
            assert int(te["run_lengths"].sum()) == n
        else:
            for k in first:
                assert torch.equal(te[k], first[k]), (rep, k)
