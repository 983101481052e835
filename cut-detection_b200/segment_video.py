"""segment_video.py -- the reference CLI (reference segment_video.py:20-126), same flags, log lines and CSV bytes,
running on the B200 kernels.

    python segment_video.py <video> [--output_path CSV] [--base-threshold 100] [--blank-threshold 10]
                            [--batch-size 128] [--print-every 50] [--frame-limit N] [--cpu] [@argsfile]

Differences from the reference, all behind the same interface:
  * decoded frames go to the GPU as uint8 (only the source rows the resize reads), K1 does resize/layout/normalise
    there and feeds the classifier directly; logits never leave the device until the run table is built;
  * ``--cpu`` is accepted for compatibility but refused: this build has no CPU path.
"""
from frameID.net import load_default_net
from frameID.data import VideoDataset
from frameID.segmentation import Segmentation

import torch

import argparse
import logging
import os

logging.basicConfig(
    level="INFO",
    format="[%(asctime)s] %(levelname)s [%(name)s.%(funcName)s:%(lineno)d] %(message)s",
)


def main(args):

    if not os.path.isfile(args.input_path):
        raise ValueError(f"{args.input_path} does not exist.")

    if args.cpu or not torch.cuda.is_available():
        raise RuntimeError("this build runs on a CUDA device (sm_100a) only; there is no CPU path")
    device = "cuda:0"
    logging.info(f"Using {device}")

    ds = VideoDataset(args.input_path, resize=256, device=device)

    net, params = load_default_net()
    net.eval()
    net.to(device)
    logging.info("Loaded default classifier.")

    copy_stream = torch.cuda.Stream(device=device)
    yy = []
    with torch.no_grad():
        for i, (plan, frames, compact) in enumerate(ds.frame_batches(args.batch_size)):
            with torch.cuda.stream(copy_stream):
                on_device = frames.to(device, non_blocking=True)
            torch.cuda.current_stream().wait_stream(copy_stream)
            yy.append(net.forward_frames(plan, on_device, compact))
            on_device.record_stream(torch.cuda.current_stream())

            if args.print_every > 0:
                if i % args.print_every == args.print_every - 1:
                    logging.info(f"Scored batch {i+1} ({(i+1) * args.batch_size} frames).")

            # same check as the reference: after the batch is scored, strict '>'
            if args.frame_limit is not None and (i + 1) * args.batch_size > args.frame_limit:
                break

        yy = torch.cat(yy, 0)

        seg = Segmentation(yy)
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

if __name__ == "__main__":

    args = sv_parser.parse_args()

    main(args)
