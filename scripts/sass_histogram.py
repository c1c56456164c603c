#!/usr/bin/env python
"""Static SASS opcode histogram of one kernel of libhamilton.so (no GPU needed): cuobjdump -sass + a count per opcode.

    python scripts/sass_histogram.py [mangled-name-substring] [path/to/lib.so]
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def main():
    pat = sys.argv[1] if len(sys.argv) > 1 else "ode_kernelILb0ELi1"
    lib = sys.argv[2] if len(sys.argv) > 2 else str(ROOT / "tmcmc202601_b200" / "csrc" / "libhamilton.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, hist = None, collections.Counter()
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and pat in cur and "coop" not in cur.replace(pat, ""):
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                hist[m.group(2)] += 1
    tot = sum(hist.values())
    fp64 = sum(hist[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"{pat}: {tot} SASS instructions, FP64 pipe {fp64} ({100.0 * fp64 / max(tot, 1):.1f} %)")
    print("  ".join(f"{k}:{v}" for k, v in hist.most_common(24)))


if __name__ == "__main__":
    main()
