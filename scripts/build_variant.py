#!/usr/bin/env python
"""Build a kernel variant of libhamilton into scripts/_variants/libhamilton_<tag>.so (development aid):

    python scripts/build_variant.py <tag> -DHMX_SOLVE_DIVFREE [...]
    HMX_LIB=scripts/_variants/libhamilton_<tag>.so python scripts/quick_perf.py 65536
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

tag, extra = sys.argv[1], sys.argv[2:]
out = ROOT / "scripts" / "_variants"
out.mkdir(exist_ok=True)
lib = out / f"libhamilton_{tag}.so"
flags = [f for f in g.NVCC_FLAGS if f not in [e for e in extra if e.startswith("-U")]]
flags = [f for f in flags if not any(e.startswith("-U") and f == "-D" + e[2:] for e in extra)]
cmd = ["nvcc", *flags, *[e for e in extra if not e.startswith("-U")], "-I", str(g.CSRC), "-I", str(ROOT / "include"), "-o", str(lib), str(g.CSRC / "hmx_api.cu")]
res = subprocess.run(cmd, capture_output=True, text=True)
log = res.stdout + res.stderr
(out / f"build_{tag}.log").write_text(log)
if res.returncode:
    sys.stderr.write(log[-3000:])
    raise SystemExit(1)
import re
m = re.search(r"_ZN3hmx10ode_kernelILb0ELi1EE.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers", log, re.S)
print(lib, "ode_kernel<0,1>:", m.groups() if m else "?")
