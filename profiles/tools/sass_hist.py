"""Opcode histogram of one kernel's SASS, optionally restricted to an address range (the main loop).
  python profiles/tools/sass_hist.py <mangled-or-substring> [lo_hex hi_hex]"""
import re
import subprocess
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "gaot_3d_b200", "libgaot_b200.so")
pat = sys.argv[1]
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, hist, n = None, {}, 0
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            a = int(m.group(1), 16)
            if lo <= a < hi:
                toks = m.group(2).split()
                op = toks[1] if toks[0].startswith("@") else toks[0]
                op = ".".join(op.split(".")[:2])
                hist[op] = hist.get(op, 0) + 1
                n += 1
print(n, "instructions")
for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:45]:
    print(f"{v:6d} {k}")
