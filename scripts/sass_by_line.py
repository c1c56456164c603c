#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (no GPU needed): nvdisasm -gi on the library's cubin.

    python scripts/sass_by_line.py [lib.so] [mangled-function-name]
"""
import collections
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "tmcmc202601_b200" / "csrc" / "libhamilton.so")
func = sys.argv[2] if len(sys.argv) > 2 else "_ZN3hmx10ode_kernelILb0ELi1EEEvNS_9OdeParamsE"
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, capture_output=True)
    cubin = next(Path(d).glob("*.cubin"))
    txt = subprocess.run(["nvdisasm", "-gi", str(cubin)], capture_output=True, text=True).stdout
lines = txt.split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text." + func + ":")][0]
by = collections.defaultdict(collections.Counter)
cur, fresh = None, True
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.strip().startswith(".section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            fresh = False
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
    if m:
        fresh = True
        by[cur][m.group(3)] += 1
tot = sum(sum(c.values()) for c in by.values())
print(f"{func}: {tot} SASS instructions")
for k in sorted(by, key=lambda k: (k[0], k[1]) if k else ("", 0)):
    c = by[k]
    n = sum(c.values())
    if n >= 4:
        print(f"{k[0]}:{k[1]:<4d} {n:5d}  " + " ".join(f"{o}:{v}" for o, v in c.most_common(6)))
