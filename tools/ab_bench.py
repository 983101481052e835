#!/usr/bin/env python3
"""Same-box A/B of bench.py under different net options or builds of the library (one gpurun call).

Boxes differ by several per cent and the default 80-step run is power-capped (profiles/README.md, v9), so two variants are only
comparable when they are measured on ONE box, interleaved, a few times each:

    python tools/ab_bench.py --reps 2 --steps 30,80 -- "" "sub_batch=296" "conv1_variant=1" "lib=gpurun_out/libcutdet_old.so"

Each variant is a space-separated list of NAME=VALUE pairs ("" = the defaults): net options (bench.py --net-opt) or
``lib=PATH`` (bench.py --lib: another build of libcutdet_b200.so) or ``lanes=N`` / ``chunk=N`` (bench.py --lanes / --chunk).  Prints one line per run and a summary table
(mean / min / max frames per second per variant and step count); with --out also writes them to a text file for profiles/."""
import argparse
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--steps", default="80", help="comma-separated step counts")
    ap.add_argument("--timeout", type=int, default=120, help="seconds per bench run")
    ap.add_argument("--out", default=None)
    ap.add_argument("variants", nargs="+")
    a = ap.parse_args()
    steps = [int(s) for s in a.steps.split(",")]
    results = {}
    lines = []
    for rep in range(a.reps):
        for v in a.variants:                                     # interleaved: drift over the call hits every variant alike
            for st in steps:
                cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(st), "--no-e2e", "--no-cpu-baseline"]
                for kv in v.split():
                    k, _, val = kv.partition("=")
                    cmd += ["--lib", val] if k == "lib" else [f"--{k}", val] if k in ("lanes", "chunk") else ["--net-opt", kv]
                try:
                    r = subprocess.run(cmd, capture_output=True, text=True, timeout=a.timeout)
                    d = json.loads(r.stdout.strip().splitlines()[-1])
                    ok = bool(d.get("parity", {}).get("runs_equal_plan")) and bool(d.get("parity", {}).get("smoothed_equal_oracle"))
                    line = f"AB rep={rep} variant='{v}' steps={st} frames_per_s={d['value']:.0f} ms_per_step={d['ms_per_step']:.4f} parity={ok} clocks={d.get('clocks')}"
                    if ok:
                        results.setdefault((v, st), []).append(d["value"])
                except Exception as e:                            # a failed variant must not cost the others their measurement
                    line = f"AB rep={rep} variant='{v}' steps={st} FAILED: {type(e).__name__}: {e}"
                print(line, flush=True)
                lines.append(line)
    lines.append("")
    lines.append(f"{'variant':40s} {'steps':>5s} {'mean':>10s} {'min':>10s} {'max':>10s}  n")
    for (v, st), vals in results.items():
        lines.append(f"{(v or '(defaults)'):40s} {st:5d} {statistics.mean(vals):10.0f} {min(vals):10.0f} {max(vals):10.0f}  {len(vals)}")
    print("\n".join(lines[-(len(results) + 1):]))
    if a.out:
        with open(a.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
