#!/usr/bin/env python3
"""profiles/sass_summary.txt: per-kernel counts of the SASS mnemonics that show what a kernel is made of (cuobjdump -sass on
the built library; no GPU needed).  UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = cp.async.bulk.tensor (TMA),
UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, LDGSTS = cp.async, HMMA = legacy mma.sync (there must be none).

    python profiles/make_sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cut-detection_b200", "cutdet", "_lib", "libcutdet_b200.so")
MNEMONICS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "SYNCS", "LDGSTS", "USETMAXREG", "HMMA", "STL", "LDL")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts, order, name = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            raw = m.group(1)
            dem = subprocess.run(["c++filt", raw], capture_output=True, text=True).stdout.strip() or raw
            dem = dem.replace("(anonymous namespace)::", "")
            name = dem.split("(")[0].replace("void ", "").replace("cutdet::", "")
            counts[name] = collections.Counter()
            order.append(name)
            continue
        if name is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            counts[name]["_instructions"] += 1
            for k in MNEMONICS:
                if op.startswith(k):
                    counts[name][k] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: architectures {archs}; columns = occurrences in the kernel's SASS")
    print(f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{k:>10s}" for k in MNEMONICS))
    total = collections.Counter()
    for n in order:
        c = counts[n]
        total.update(c)
        print(f"{n[:70]:70s} {c['_instructions']:7d} " + " ".join(f"{c[k]:10d}" for k in MNEMONICS))
    print(f"{'TOTAL':70s} {total['_instructions']:7d} " + " ".join(f"{total[k]:10d}" for k in MNEMONICS))


if __name__ == "__main__":
    main()
