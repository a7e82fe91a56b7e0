"""Per-kernel SASS evidence for libgaot_b200.so (CPU-only: cuobjdump -sass on the built library):
counts of the Blackwell tensor-core / tensor-memory / TMA mnemonics per kernel -> profiles/<tag>_sass_summary.txt.
  python profiles/tools/sass_summary.py r02
UTCHMMA/UTCQMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP/UBLKRED = bulk copy/reduce,
LDGSTS = cp.async, RED/ATOM = global atomics, HMMA = legacy mma.sync (none expected)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "gaot_3d_b200", "libgaot_b200.so")
MNEM = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "LDGSTS", "HMMA", "RED", "ATOM", "MUFU.EX2", "SHFL"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "rXX"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    rows, name, cnt = [], None, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                rows.append((name, cnt))
            name, cnt = m.group(1), dict.fromkeys(MNEM, 0)
            cnt["_n"] = 0
            continue
        if name and "/*" in line and ";" in line:
            body = line.split("*/", 1)[-1]
            cnt["_n"] += 1
            for k in MNEM:
                if re.search(r"\b" + re.escape(k) + r"\b" if "." not in k else re.escape(k), body):
                    cnt[k] += 1
    if name:
        rows.append((name, cnt))
    dem = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
    out = [f"# SASS summary of gaot_3d_b200/libgaot_b200.so (sm_100a), {len(rows)} kernels; columns: instructions " + " ".join(MNEM)]
    for (n, c), d in sorted(zip(rows, dem), key=lambda t: t[1]):
        d = re.sub(r"\(.*", "", d)[:70]
        out.append(f"{d:70s} {c['_n']:6d} " + " ".join(f"{c[k]:4d}" for k in MNEM))
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print(path, len(rows))


if __name__ == "__main__":
    main()
