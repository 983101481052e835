"""segment_video.py -- the reference CLI (reference segment_video.py:20-126), same flags, log lines and CSV bytes,
running on the B200 kernels.

    python segment_video.py <video> [--output_path CSV] [--base-threshold 100] [--blank-threshold 10]
                            [--batch-size 128] [--print-every 50] [--frame-limit N] [--cpu] [@argsfile]
                            [--decode-workers W]

Differences from the reference, all behind the same interface:
  * the frames are decoded by W worker processes, each on its own contiguous time range of the video, into one reusable
    pinned ring (cutdet.decode; W = 1 is the reference's sequential decode moved off the main thread);
  * decoded frames go to the GPU as uint8 (only the source rows the resize reads) through ``FramePipeline.push_host``: K1 does
    resize/layout/normalise there and feeds the classifier directly, K4/K5 turn the logits into a run table chunk by chunk --
    logits never leave the device and nothing per-frame is kept until the end;
  * ``--cpu`` is accepted for compatibility but refused: this build has no CPU path.
"""
from frameID.net import load_default_net
from frameID.data import open_video
from frameID.segmentation import Segmentation

import numpy as np
import torch

import argparse
import logging
import os

logging.basicConfig(
    level="INFO",
    format="[%(asctime)s] %(levelname)s [%(name)s.%(funcName)s:%(lineno)d] %(message)s",
)


def open_decode(path, batch_size, frame_limit=None, decode_workers=None, timings=None):
    """Start the decoder processes on their time ranges of the video.  No CUDA call is made here on purpose: call it BEFORE the
    process creates its CUDA context (forking a process that holds one costs ~70 ms per worker).  Returns the state
    ``score_video`` consumes."""
    import time
    from cutdet import decode, engine
    t0 = time.perf_counter()
    n_meta, h, w = decode.probe_video(path)
    if h == 0:
        raise IndexError("the video has no decodable frame")          # what Segmentation(empty scores) ends in
    # the reference stops after the first batch that takes it past the limit: (i + 1) * batch_size > frame_limit
    n_frames = n_meta
    if frame_limit is not None:
        n_frames = (frame_limit // batch_size + 1) * batch_size
        if n_meta > 0:
            n_frames = min(n_meta, n_frames)
    to_eof = frame_limit is None           # the last worker reads until cap.read() fails, like the reference
    workers = decode.default_workers() if decode_workers is None else max(1, int(decode_workers))
    if n_frames <= 0:
        workers = 1
    # the ring's slots hold at most 32 frames: the workers hand over small chunks (a short clip keeps all of them busy and the
    # pinned ring stays small: registering it is a fixed cost), the kernels take whatever arrives, the log still counts batches
    chunk = max(1, min(batch_size, 32))
    rows = engine.resize_rows(h, w, 256)
    pool = decode.DecodePool(path, rows, h, w, chunk, workers, max(n_frames, 0), to_eof=to_eof, slots_per_worker=3, pin=False)
    if timings is not None:
        timings["probe_and_workers"] = time.perf_counter() - t0
    return {"path": path, "pool": pool, "h": h, "w": w, "chunk": chunk, "rows": rows, "n_frames": n_frames, "to_eof": to_eof,
            "batch_size": batch_size, "workers": workers}


def score_video(dec, net, print_every=0, device="cuda:0", timings=None):
    """Decode -> K1 -> CNN -> K4 -> K5 for a whole video: returns (DeviceRunTable of the initial runs, frames scored).
    What the reference's batch loop plus ``Segmentation.__init__`` compute (segment_video.py:38-62), streamed.
    ``dec``: the state ``open_decode`` returned.  ``timings``: a dict that receives wall-clock seconds per phase."""
    import time
    from cutdet import decode, engine, pipeline
    t_start = time.perf_counter()

    def lap(name):
        nonlocal t_start
        if timings is not None:
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + now - t_start
            t_start = now

    h, w, chunk, batch_size = dec["h"], dec["w"], dec["chunk"], dec["batch_size"]
    native = net._native()
    plan = engine.ResizePlan.for_video(h, w, 256)
    assert np.array_equal(plan.rows, dec["rows"])
    pool = dec["pool"]
    for attempt in range(2):
        retry = None
        try:
            pool.pin()
            # a run table never has more rows than frames; a range that outgrows this (a container that under-reports its
            # length) is reported as an overflow by the table, not silently truncated
            capacity = max(hi - lo for lo, hi in pool.ranges) + (1 << 16)
            pipe = pipeline.FramePipeline(native, plan, chunk, capacity, device, n_ranges=pool.n_workers)
            lap("plan_ring_pipeline")
            scored = 0
            for worker, frames, first_frame, slot in pool:
                uploaded = pipe.push_host(frames, compact=True, rng=worker)
                pool.release(slot, uploaded)
                batches = pipe.n_frames // batch_size
                if print_every > 0 and batches > scored and batches % print_every == 0:
                    logging.info(f"Scored batch {batches} ({batches * batch_size} frames).")
                scored = batches
            lap("decode_and_score")
            table, total = pipe.finish_ranges()
            try:
                table.count()
            except engine.ShardOverflow as e:        # a range with more runs than the default join capacity
                table, total = pipe.finish_ranges(capacity=1 << int(e.needed - 1).bit_length())
            lap("finish")
            return table, pipe.n_frames
        except decode.SeekMismatch as e:
            if attempt or pool.n_workers == 1:
                raise
            retry = e
        finally:
            torch.cuda.synchronize()
            pool.close()
            lap("teardown")
        logging.warning(f"{retry}; decoding sequentially instead")
        pool = decode.DecodePool(dec["path"], dec["rows"], h, w, chunk, 1, max(dec["n_frames"], 0), to_eof=dec["to_eof"],
                                 slots_per_worker=3, pin=False)


def main(args):

    if not os.path.isfile(args.input_path):
        raise ValueError(f"{args.input_path} does not exist.")

    if args.cpu:
        raise RuntimeError("this build runs on a CUDA device (sm_100a) only; there is no CPU path")
    device = "cuda:0"

    # the decoder processes are forked first, before this process has touched the CUDA driver at all (not even the device
    # query below): they inherit nothing of it, and they decode while the context is created
    import time
    timings = getattr(args, "timings", None)
    dec = open_decode(args.input_path, args.batch_size, args.frame_limit, getattr(args, "decode_workers", None), timings)

    t0 = time.perf_counter()
    if not torch.cuda.is_available():
        dec["pool"].close()
        raise RuntimeError("this build runs on a CUDA device (sm_100a) only; there is no CPU path")
    logging.info(f"Using {device}")
    if timings is not None:
        timings["driver_init"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    net, params = load_default_net()
    net.eval()
    net.to(device)
    net._native()                       # CUDA context, library load and weight upload happen here, while the workers decode
    logging.info("Loaded default classifier.")
    if timings is not None:
        timings["context_and_weights"] = time.perf_counter() - t0

    with torch.no_grad():
        table, n_scored = score_video(dec, net, args.print_every, device, timings)

        t0 = time.perf_counter()
        seg = Segmentation.from_table(table)
        logging.info(f"Found {len(seg)} initial segments")
        seg.glue_orphans(args.base_threshold, args.blank_threshold)
        logging.info(f"Revised to {len(seg)} segments through orphan combination.")
        seg.combine_adjacent_segments()
        logging.info(f"Revised to {len(seg)} segments through matching adjacent combination.")

        if args.output_path is None:
            out_path = os.path.splitext(args.input_path)[0] + "_segments.csv"
        else:
            out_path = args.output_path

        logging.info(f"Writing {len(seg)} segments to {out_path}")
        seg.write_csv(out_path)
        if timings is not None:
            timings["smooth_and_csv"] = time.perf_counter() - t0


sv_parser = argparse.ArgumentParser("Segment a video into scenes.", fromfile_prefix_chars="@")
sv_parser.add_argument("input_path", type=str, help="Path to video to segment.")
sv_parser.add_argument("--output_path", type=str, default=None, help="Path to output csv")
sv_parser.add_argument("--base-threshold", type=int, default=100,
                       help="Number of frames below which an A22 or EZ segment will be considered an orphan.")
sv_parser.add_argument("--blank-threshold", type=int, default=10,
                       help="Number of frames below which a blank segment will be considered an orphan.")
sv_parser.add_argument("--batch-size", type=int, default=128, help="Batch size for loading frames.")
sv_parser.add_argument("--print-every", type=int, default=50, help="Log message every n batches. 0 to disable.")
sv_parser.add_argument("--frame-limit", type=int, default=None,
                       help="Limit how many frames are processed. Mainly for testing.")
sv_parser.add_argument("--cpu", action="store_true",
                       help="Accepted for compatibility; refused (this build has no CPU path).")
sv_parser.add_argument("--decode-workers", type=int, default=None,
                       help="Decoder processes, each on its own time range of the video (default: half the host cores, at most 8).")

if __name__ == "__main__":

    args = sv_parser.parse_args()

    main(args)
