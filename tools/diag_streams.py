import sys, numpy as np, torch
sys.path[:0] = ["/root/repo", "/root/repo/cut-detection_b200"]
from cutdet import engine
from oracle import net as onet
wts, params = onet.load_weights_npz("/root/repo/cut-detection_b200/frameID/prod_net/prod_net_weights.npz")
def run(variant, mode, no_pdl=0, reps=8):
    nets = [engine.NativeNet(wts, params["avg_pool_size"]) for _ in range(2)]
    for n in nets:
        n.set_option("conv1_variant", variant); n.set_option("no_pdl", no_pdl)
    rng = np.random.default_rng(0)
    h, w, batch = 720, 1280, 460
    frames = torch.from_numpy(rng.integers(0, 256, (batch, h, w, 3), dtype=np.uint8)).cuda()
    plan = engine.ResizePlan.for_video(h, w, 256)
    want = nets[0].forward_frames(plan, frames).clone()
    w1 = nets[1].forward_frames(plan, frames).clone()      # warm nets[1] alone first
    torch.cuda.synchronize()
    assert torch.equal(want, w1)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = []
    for rep in range(reps):
        for i, (n, st) in enumerate(zip(nets, streams)):
            if mode == "one_net" and i == 1: continue
            with torch.cuda.stream(st):
                outs.append((rep, i, n.forward_frames(plan, frames, out=torch.empty_like(want))))
    torch.cuda.synchronize()
    bad = []
    for rep, i, o in outs:
        if not torch.equal(o, want):
            rows = (o != want).any(1).nonzero().flatten().cpu().numpy()
            bad.append((rep, i, len(rows), int(rows.min()), int(rows.max()), float((o - want).abs().max())))
    print(f"variant={variant} mode={mode} no_pdl={no_pdl}: {len(bad)} of {len(outs)} outputs differ", bad[:6])
for variant in (0, 1):
    for no_pdl in (0, 1):
        run(variant, "two_nets", no_pdl)
run(0, "one_net")
