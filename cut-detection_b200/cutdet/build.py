"""In-tree build of libcutdet_b200.so (nvcc, sm_100a only).

``python -m cutdet.build`` or ``__graft_entry__.build()``.  The shared object is written next to this file
(``cutdet/_lib/libcutdet_b200.so``): it is git-ignored but travels with the tree to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(HERE)
CSRC = os.path.join(PKG_ROOT, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libcutdet_b200.so")
STAMP_PATH = os.path.join(LIB_DIR, "build_stamp.txt")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return nvcc if os.path.exists(nvcc) else None


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    files.append(os.path.join(os.path.dirname(PKG_ROOT), "include", "cutdet_b200.h"))
    root = os.path.dirname(PKG_ROOT)
    for p in files:
        h.update(os.path.relpath(p, root).encode())      # repo-relative: the stamp holds on any checkout location
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    try:
        return os.path.isfile(LIB_PATH) and open(STAMP_PATH).read().strip() == _digest()
    except OSError:
        return False


def ensure_current() -> str:
    """Path of a library that matches the sources: rebuilds a stale or missing one when nvcc is here, raises otherwise.
    (Called by cutdet._cabi.lib(): tests and the CLI never run a binary older than the sources next to it.)"""
    if is_current():
        return LIB_PATH
    if find_nvcc():
        return build_native(force=True)
    if os.path.isfile(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is older than the sources in {CSRC} and nvcc is not available to rebuild it")
    raise RuntimeError(f"libcutdet_b200.so not found at {LIB_PATH} and nvcc is not available: build it with "
                       "`python -m cutdet.build` (or __graft_entry__.build()); there is no fallback implementation")


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into one shared library.  Returns its path."""
    if not force and is_current():
        return LIB_PATH
    nvcc = find_nvcc()
    if not nvcc:
        raise RuntimeError("nvcc not found: cannot build libcutdet_b200.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            failed.append(f"--- {src}\n{out}")
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(failed))
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    with open(STAMP_PATH, "w") as f:
        f.write(_digest() + "\n")
    return LIB_PATH


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
