#!/usr/bin/env python3
"""Split a video into individual frames -- mirror of the reference's training_scripts/split_video.py:1-55 (same arguments, same
messages, same ``frame_{i:07}.jpg`` files), with the resize (split_video.py:47-50, ``cv2.resize(..., INTER_LINEAR)``) done by the
preprocessing kernel K1 on the GPU in batches (``cutdet_preprocess_u8``: bit-exact with OpenCV, so the JPEG bytes are the ones
the reference writes).  Decoding and JPEG encoding stay with OpenCV."""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cutdet import engine  # noqa: E402
from frameID.data import open_video  # noqa: E402

BATCH = 64


def main(argv=None):
    parser = argparse.ArgumentParser("Split a video into individual frames.")
    parser.add_argument("input_path", type=str, help="Path to video to turn into frames.")
    parser.add_argument("output_dir", type=str, help="Path to directory to write images. Will be created if it doesn't exist.")
    parser.add_argument("--resize", type=int, default=0, help="Size of larger dimension.")
    parser.add_argument("--max-frames", type=int, default=-1, help="Number of frames to save.")
    args = parser.parse_args(argv)

    if not os.path.isfile(args.input_path):
        raise ValueError(f"{args.input_path} does not exist.")
    cap, v_properties = open_video(args.input_path)
    frame_limit = v_properties["length"] if args.max_frames < 0 else args.max_frames
    print(f"Processing {frame_limit} frames from {args.input_path}.")
    if not os.path.isdir(args.output_dir):
        os.mkdir(args.output_dir)

    plan = None
    if args.resize > 0:
        new_width = args.resize
        new_height = int(v_properties["height"] * (new_width / v_properties["width"]))
        plan = engine.ResizePlan(v_properties["height"], v_properties["width"], new_height, new_width)

    pending = []        # (frame index, decoded frame)

    def flush():
        if not pending:
            return
        if plan is None:
            out = [f for _, f in pending]
        else:
            dev = torch.from_numpy(np.stack([f for _, f in pending])).cuda()
            out = list(engine.preprocess_u8(plan, dev).cpu().numpy())
        for (i, _), frame in zip(pending, out):
            cv2.imwrite(f"{args.output_dir}/frame_{i:07}.jpg", frame)
        pending.clear()

    for i in range(frame_limit):
        if i % 5000 == 4999:
            print(f"Processing frame {i+1}")
        ret, frame = cap.read()
        if ret:
            pending.append((i, frame))
            if len(pending) == BATCH:
                flush()
    flush()
    print("Done")


if __name__ == "__main__":
    main()
