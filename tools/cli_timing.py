"""Run the drop-in CLI (segment_video.main) in THIS fresh process and print one JSON line with its wall-clock phases.
bench.py's ``cli`` leg calls it as a subprocess: a process that already holds a CUDA context (bench.py itself) would hide what
a user's run pays once -- context creation, forking the decoders, pinning the ring -- and overstate the cost of the forks.

    python tools/cli_timing.py VIDEO OUT.csv [decode_workers|- [worker_nice [decoder_threads|auto|none]]]
"""
import json
import os
import sys
import time

T_PROCESS = time.perf_counter()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cut-detection_b200"))


def main():
    import segment_video as sv          # imports torch
    t_imported = time.perf_counter()
    argv = [sys.argv[1], "--output_path", sys.argv[2], "--print-every", "0"]
    if len(sys.argv) > 3 and sys.argv[3] != "-":
        argv += ["--decode-workers", sys.argv[3]]
    if len(sys.argv) > 4:                                   # scheduling experiments (cutdet.decode's module defaults)
        from cutdet import decode
        decode.WORKER_NICE = int(sys.argv[4])
        if len(sys.argv) > 5:
            decode.DECODER_THREADS = {"auto": "auto", "none": None}.get(sys.argv[5], sys.argv[5])
    ns = sv.sv_parser.parse_args(argv)
    ns.timings = {}
    t0 = time.perf_counter()
    sv.main(ns)
    dt = time.perf_counter() - t0
    print(json.dumps({"seconds": dt, "import_s": round(t_imported - T_PROCESS, 3),
                      "phases_s": {k: round(v, 3) for k, v in ns.timings.items()},
                      "other_s": round(dt - sum(ns.timings.values()), 3)}))


if __name__ == "__main__":
    main()
