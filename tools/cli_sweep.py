"""Scheduling sweep of the CLI's decoder processes on one clip (fresh process per run): workers x nice x ffmpeg threads.
    python tools/cli_sweep.py [n_frames]      -> one JSON line per run"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cut-detection_b200"))


def main():
    import cv2
    from cutdet import synth
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1800
    tmp = tempfile.mkdtemp(prefix="cutdet_sweep_")
    path = os.path.join(tmp, "clip.mp4")
    clip = synth.SyntheticClip(720, 1280, n, seed=0)
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 30, (1280, 720))
    for lo in range(0, n, 256):
        for f in clip.frames_numpy(lo, min(256, n - lo)):
            vw.write(f)
    vw.release()
    print(json.dumps({"cores": os.cpu_count(), "frames": n, "bytes": os.path.getsize(path)}), flush=True)
    cases = [("1", "0", "none")] + [("8", "0", "none"), ("8", "10", "auto"), ("4", "10", "auto"), ("8", "10", "1")] * 3 + [("1", "0", "none")]
    for workers, nice, threads in cases:
        cmd = [sys.executable, os.path.join(ROOT, "tools", "cli_timing.py"), path, os.path.join(tmp, "o.csv"), workers, nice, threads]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        if p.returncode:
            print(json.dumps({"case": [workers, nice, threads], "error": p.stderr[-500:]}), flush=True)
            continue
        r = json.loads(p.stdout.strip().splitlines()[-1])
        print(json.dumps({"workers": workers, "nice": nice, "threads": threads, "fps": round(n / r["seconds"], 1),
                          "seconds": round(r["seconds"], 3), **r["phases_s"], "other": r["other_s"]}), flush=True)


if __name__ == "__main__":
    main()
