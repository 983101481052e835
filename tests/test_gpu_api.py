"""The reference-facing Python surface (frameID.*) end to end on the GPU, as segment_video.py uses it."""
import os

import numpy as np
import pytest
import torch

import kat_inputs
from oracle import net as onet
from oracle import segmentation as oseg

pytestmark = pytest.mark.gpu


def test_load_default_net_contract(prod_weights, golden_dir):
    from frameID.net import load_default_net, FrameConvNet, FrameLinearNet
    net, params = load_default_net()
    assert params["conv_channels"] == 48 and params["linear_output_size"] == 3
    assert isinstance(net, torch.nn.Sequential) and isinstance(net[0], FrameConvNet) and isinstance(net[1], FrameLinearNet)
    assert net[0].num_params() == 43200 and net[1].num_params() == 24771
    keys = set(net[0].state_dict().keys())
    assert "conv_layers.0.conv.weight" in keys and "conv_layers.2.bn.running_var" in keys
    assert "layers.1.linear.bias" in net[1].state_dict()
    net.eval()
    net.to("cuda:0")
    x = torch.from_numpy(kat_inputs.smooth_images(48)).to("cuda:0")
    with torch.no_grad():
        y = net(x)
    assert y.dtype == torch.float32 and tuple(y.shape) == (48, 3)
    kat = np.load(os.path.join(golden_dir, "net_kat.npz"))
    assert np.abs(y.cpu().numpy() - kat["smooth48_eager"]).max() <= 0.05
    # trunk and head also work on their own, like the reference's modules
    feats = net[0](x[:4])
    assert tuple(feats.shape) == (4, 768)
    assert np.abs(net[1](feats).cpu().numpy() - kat["smooth48_eager"][:4]).max() <= 0.05


def test_training_mode_and_cpu_inputs_raise():
    from frameID.net import load_default_net
    net, _ = load_default_net()
    net.to("cuda:0")
    with pytest.raises(ValueError):
        net(torch.zeros((1, 3, 144, 256), device="cuda:0"))           # training mode: batch statistics of ONE frame (as nn.BatchNorm)
    net.eval()
    with pytest.raises(RuntimeError):
        net(torch.zeros((1, 3, 144, 256)))                            # CPU input: no fallback


def test_three_file_checkpoint_format(tmp_path, prod_weights):
    """load_and_glue_nets reads the reference's JSON + two state_dict files (net.py:193-217)."""
    import json
    from frameID.net import load_and_glue_nets, load_default_net
    net, params = load_default_net()
    torch.save(net[0].state_dict(), tmp_path / "m_conv.pt")
    torch.save(net[1].state_dict(), tmp_path / "m_linear.pt")
    json.dump(params, open(tmp_path / "m.json", "w"))
    net2, params2 = load_and_glue_nets(str(tmp_path / "m.json"), str(tmp_path / "m_conv.pt"), str(tmp_path / "m_linear.pt"))
    assert params2 == params
    net.eval().to("cuda")
    net2.eval().to("cuda")
    x = torch.from_numpy(kat_inputs.smooth_images(3)).cuda()
    assert torch.equal(net(x), net2(x))


def test_torchscript_export_as_weight_source(golden_dir):
    """the reference's saved_model_trace.pt (training_scripts/make_torchscript_model.py:17-34, written by the
    unmodified reference: tests/golden/make_trace_golden.py) loads into the native path and gives the shipped net's logits."""
    from frameID.net import load_default_net, load_torchscript_net
    net, params = load_default_net()
    net2, params2 = load_torchscript_net(os.path.join(golden_dir, "saved_model_trace.pt"))
    assert {k: params[k] for k in params2} == params2
    net.eval().to("cuda")
    net2.eval().to("cuda")
    x = torch.from_numpy(kat_inputs.smooth_images(5, seed=11)).cuda()
    assert torch.equal(net(x), net2(x))
    kat = np.load(os.path.join(golden_dir, "net_kat.npz"))      # what the reference's traced module gave for these inputs
    got = net2(torch.from_numpy(kat_inputs.smooth_images(48)).cuda()).cpu().numpy()
    assert np.abs(got - kat["smooth48_traced"]).max() <= 0.05


def test_cli_on_a_synthetic_clip(tmp_path, prod_weights):
    """segment_video.py on a small synthetic mp4: CSV equals the oracle's run over the same decoded frames."""
    import cv2
    import subprocess, sys
    from oracle import preprocess as opre
    w, h, n = 640, 360, 330
    path = str(tmp_path / "clip.mp4")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 30, (w, h))
    plan = [(kat_inputs.stripes(h, w, 20, True), 150), (np.zeros((h, w, 3), np.uint8), 30),
            (kat_inputs.stripes(h, w, 20, False), 150)]
    for frame, count in plan:
        for _ in range(count):
            vw.write(frame)
    vw.release()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out_csv = str(tmp_path / "out.csv")
    r = subprocess.run([sys.executable, os.path.join(root, "cut-detection_b200", "segment_video.py"), path,
                        "--output_path", out_csv, "--batch-size", "64", "--print-every", "2"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Found" in r.stderr and "Writing" in r.stderr
    # oracle on the same decoded frames
    cap = cv2.VideoCapture(path)
    frames = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        frames.append(f)
    wts, params = prod_weights
    logits = onet.forward_f32(wts, opre.preprocess_batch(np.stack(frames), 256), params["avg_pool_size"])
    want = oseg.segment(logits, 100, 10)[3]
    assert open(out_csv, "rb").read() == want
    assert want == b"0,a22\r\n150,b\r\n180,ez\r\n"
    # any number of decode workers (time ranges joined on the device) gives the same bytes; 1 = sequential decode
    for workers in ("1", "2", "3"):
        r = subprocess.run([sys.executable, os.path.join(root, "cut-detection_b200", "segment_video.py"), path,
                            "--output_path", out_csv, "--batch-size", "32", "--print-every", "0", "--decode-workers", workers],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert open(out_csv, "rb").read() == want, workers
    # default output path and the --frame-limit quirk (check happens after the batch is scored, strict >)
    r = subprocess.run([sys.executable, os.path.join(root, "cut-detection_b200", "segment_video.py"), path,
                        "--batch-size", "64", "--frame-limit", "64", "--print-every", "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    default_csv = os.path.splitext(path)[0] + "_segments.csv"
    got = open(default_csv, "rb").read()
    assert got == oseg.segment(logits[:128], 100, 10)[3]
    r = subprocess.run([sys.executable, os.path.join(root, "cut-detection_b200", "segment_video.py"),
                        str(tmp_path / "missing.mp4")], capture_output=True, text=True)
    assert r.returncode != 0 and "does not exist" in r.stderr


def test_videodataset_iteration(tmp_path):
    import cv2
    from frameID.data import VideoDataset
    from oracle import preprocess as opre
    w, h = 480, 270
    path = str(tmp_path / "v.mp4")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 30, (w, h))
    rng = np.random.default_rng(0)
    for i in range(5):
        vw.write(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
    vw.release()
    ds = VideoDataset(path, resize=256)
    assert ds.video_info == {"fps": 30, "length": 5, "width": w, "height": h} and len(ds) == 5
    got = [t.cpu().numpy() for t in ds]
    cap = cv2.VideoCapture(path)
    for t in got:
        ok, f = cap.read()
        assert ok and np.array_equal(t, opre.preprocess_frame(f, 256))
    assert len(got) == 5 and list(ds) == []        # single pass


def test_split_video_script_writes_the_references_jpegs(tmp_path):
    """training_scripts/split_video.py: the resize runs in K1 on the GPU; the files must be byte-identical to
    what the reference's loop writes (cv2.resize(INTER_LINEAR) + cv2.imwrite, split_video.py:40-53), here redone with cv2."""
    import cv2
    import subprocess, sys
    w, h, n = 640, 360, 70
    path = str(tmp_path / "clip.mp4")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 30, (w, h))
    rng = np.random.default_rng(2)
    for i in range(n):
        frame = kat_inputs.stripes(h, w, 20, i % 2 == 0).copy()
        frame[: h // 3] = rng.integers(0, 256, (h // 3, w, 3), dtype=np.uint8)
        vw.write(frame)
    vw.release()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "frames"
    r = subprocess.run([sys.executable, os.path.join(root, "cut-detection_b200", "training_scripts", "split_video.py"), path, str(out),
                        "--resize", "256", "--max-frames", "66"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.splitlines()[0] == f"Processing 66 frames from {path}." and r.stdout.splitlines()[-1] == "Done"
    cap = cv2.VideoCapture(path)
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    for i in range(66):
        ok, frame = cap.read()
        assert ok
        frame = cv2.resize(frame, (256, int(h * (256 / w))), interpolation=cv2.INTER_LINEAR)
        cv2.imwrite(f"{ref_dir}/frame_{i:07}.jpg", frame)
    names = sorted(os.listdir(out))
    assert names == sorted(os.listdir(ref_dir)) and len(names) == 66
    for name in names:
        assert open(out / name, "rb").read() == open(ref_dir / name, "rb").read(), name


def test_decode_pool_ranges_equal_sequential_decode(tmp_path):
    """cutdet.decode: W worker processes on contiguous time ranges write row-compacted frames into the pinned ring; put back in
    frame order they are byte for byte what one cv2.VideoCapture reads in sequence (reference frameID/data.py:211-213), the
    seam check passes, and the streamed run table of noisy content equals Segmentation(scores) over the whole clip."""
    import cv2
    from cutdet import decode, engine, pipeline
    from frameID.net import load_default_net
    from frameID.segmentation import Segmentation
    w, h, n = 640, 360, 300
    path = str(tmp_path / "noisy.mp4")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 30, (w, h))
    rng = np.random.default_rng(4)
    for i in range(n):
        frame = kat_inputs.stripes(h, w, 20, (i // 40) % 2 == 0).copy()
        frame[: h // 4, : w // 4] = rng.integers(0, 256, (h // 4, w // 4, 3), dtype=np.uint8)
        vw.write(frame)
    vw.release()
    cap = cv2.VideoCapture(path)
    seq = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        seq.append(f)
    seq = np.stack(seq)
    n_meta, hh, ww = decode.probe_video(path)
    assert (n_meta, hh, ww) == (n, h, w)
    plan = engine.ResizePlan.for_video(h, w, 256)
    net, _ = load_default_net()
    net.eval().to("cuda")
    with torch.no_grad():
        want_scores = net.forward_frames(plan, torch.from_numpy(seq).cuda())
    want = Segmentation(want_scores).te
    for workers in (1, 2, 4):
        pool = decode.DecodePool(path, plan.rows, h, w, 48, workers, n)
        pipe = pipeline.FramePipeline(net._native(), plan, 48, n, "cuda", n_ranges=pool.n_workers)
        got = np.zeros((n, len(plan.rows), w, 3), np.uint8)
        for wk, frames, first, slot in pool:
            assert frames.is_pinned()                     # slots are pinned as they are first handed over
            got[first:first + frames.shape[0]] = frames.numpy()
            pool.release(slot, pipe.push_host(frames, compact=True, rng=wk))
        table, total = pipe.finish_ranges()
        te = table.to_te()
        pool.close()
        assert np.array_equal(got, seq[:, plan.rows]), workers
        assert int(total.item()) == n and sum(pool.frames_decoded) == n
        for k in ("end_frames", "frame_types", "run_lengths", "start_frames"):
            assert torch.equal(te[k], want[k]), (workers, k)
        assert torch.allclose(te["score_means"], want["score_means"], rtol=2e-6)


def test_push_host_ownership_contract():
    """FramePipeline.push_host from pinned memory is asynchronous: the returned event says when the buffer may be refilled.
    Scribbling over it after the event has completed must not change the result; pageable memory is safe at once."""
    from cutdet import engine, pipeline, synth
    from frameID.net import load_default_net
    net, _ = load_default_net()
    net.eval().to("cuda")
    h, w, n = 360, 640, 200
    clip = synth.SyntheticClip(h, w, n, seed=9, runs=[(0, 70), (2, 30), (1, 100)])
    frames = torch.from_numpy(clip.frames_numpy(0, n))
    plan = engine.ResizePlan.for_video(h, w, 256)
    results = []
    for pinned in (True, False):
        pipe = pipeline.FramePipeline(net._native(), plan, 100, n, "cuda")
        buf = torch.empty((100, h, w, 3), dtype=torch.uint8)
        if pinned:
            buf = buf.pin_memory()
        for c in range(2):
            buf.copy_(frames[c * 100:(c + 1) * 100])
            ev = pipe.push_host(buf)
            if pinned:
                ev.synchronize()
            buf.fill_(255)                       # refill: must not reach the frames already handed over
        results.append(pipe.finish().to_te())
    for te in results:
        assert te["frame_types"].tolist() == [0, 2, 1] and te["run_lengths"].tolist() == [70, 30, 100]


@pytest.mark.parametrize("lanes", [2, 3])
def test_pipeline_lanes_change_nothing(lanes):
    """FramePipeline(lanes > 1) scores consecutive chunks on alternating streams with their own workspaces; the shared
    run-length encoder takes them in order.  Same table, bit for bit, as one lane -- from device frames and from pinned host
    frames, over several repetitions (reset in between) and with ragged chunk sizes; per-chunk results read behind
    wait_results() are the chunk's own."""
    from cutdet import engine, pipeline, synth
    from frameID.net import load_default_net
    net, _ = load_default_net()
    native = net.eval().to("cuda")._native()
    h, w, n, chunk = 720, 1280, 1500, 300
    runs = [(0, 310), (2, 7), (1, 290), (2, 40), (0, 153), (1, 300), (2, 3), (1, 97), (0, 300)]
    clip = synth.SyntheticClip(h, w, n, seed=21, runs=runs)
    frames = clip.frames_torch(0, n, device="cuda")
    host = frames.cpu().pin_memory()
    plan = engine.ResizePlan.for_video(h, w, 256)
    sizes = [300, 300, 148, 300, 2, 300, 150]
    assert sum(sizes) == n

    def run(pipe, source):
        pipe.reset()
        per_chunk, f0 = [], 0
        for sz in sizes:
            if source == "device":
                pipe.push_device(frames[f0:f0 + sz])
            else:
                pipe.push_host(host[f0:f0 + sz])
            pipe.wait_results()
            per_chunk.append(pipe.labels[:sz].clone())
            f0 += sz
        te = pipe.finish().to_te()
        return te, torch.cat(per_chunk).cpu()

    one = pipeline.FramePipeline(native, plan, chunk, n, "cuda")
    want, want_labels = run(one, "device")
    assert want["frame_types"].tolist() == [r[0] for r in runs] and want["run_lengths"].tolist() == [r[1] for r in runs]
    many = pipeline.FramePipeline(native, plan, chunk, n, "cuda", lanes=lanes)
    for rep in range(3):
        for source in ("device", "host"):
            te, labels = run(many, source)
            assert torch.equal(labels, want_labels), (rep, source)
            for k in ("end_frames", "frame_types", "run_lengths", "start_frames", "score_means"):
                assert torch.equal(te[k], want[k]), (rep, source, k)


def test_forward_frames_in_a_cuda_graph():
    """After its first call for a geometry the library allocates and synchronises nothing, so its launches -- programmatic
    dependent launches and tensor maps included -- can be captured into a CUDA graph as they are.  A replayed
    graph must give the eager call's logits bit for bit, also after the input buffer's contents change."""
    from cutdet import engine, synth
    from frameID.net import load_default_net
    net, _ = load_default_net()
    native = net.eval().to("cuda")._native()
    h, w, n = 720, 1280, 300
    clip = synth.SyntheticClip(h, w, 2 * n, seed=5)
    a, b = clip.frames_torch(0, n, device="cuda"), clip.frames_torch(n, n, device="cuda")
    plan = engine.ResizePlan.for_video(h, w, 256)
    want_a, want_b = native.forward_frames(plan, a).clone(), native.forward_frames(plan, b).clone()
    static_in = a.clone()
    out = torch.empty_like(want_a)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        native.forward_frames(plan, static_in, out=out)          # warm on the capture stream
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            native.forward_frames(plan, static_in, out=out)
    torch.cuda.current_stream().wait_stream(side)
    for frames, want in ((a, want_a), (b, want_b), (a, want_a)):
        static_in.copy_(frames)
        out.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, want)
