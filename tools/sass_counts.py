#!/usr/bin/env python
"""Instruction counts per kernel from the SASS of the in-tree library (no GPU needed):
    python tools/sass_counts.py > profiles/rNN_sass_counts.txt
Shows that the tensor-core kernels issue tcgen05 (UTC*MMA, LDTM / STTM), TMA (UTMALDG, UBLKCP) and no legacy HMMA, and that no kernel
uses floating-point atomics."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "phc_gnn_b200", "libphc_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
COLS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "FFMA", "FFMA2", "ELECT"]
TOT = COLS + ["UTCBAR", "SYNCS", "LDG", "STG", "LDS", "STS", "LD", "ST"]
per, cur, fatom = [], None, 0
it = iter(names)
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = collections.Counter()
        nm = re.sub(r"\(.*", "", next(it)).replace("(anonymous namespace)::", "")
        per.append((nm, cur))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
        if m.group(1) in ("ATOM", "ATOMG", "ATOMS", "RED", "REDG") and (".F32" in m.group(2) or ".F64" in m.group(2) or ".F16" in m.group(2)):
            fatom += 1
tot = collections.Counter()
for _, c in per:
    tot.update(c)
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} (sm_100a), instruction counts per kernel; {len(per)} kernels")
print("# tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP, fma.rn.f32x2 -> FFMA2;")
print(f"# floating-point atomics (ATOM/RED .F32/.F64) in the whole library: {fatom}")
print("# total: " + ", ".join(f"{k} {tot[k]}" for k in TOT))
print()
print(f"{'kernel':92s}" + "".join(f"{c:>9s}" for c in COLS))
for nm, c in per:
    if any(c[k] for k in COLS[:8]):
        print(f"{nm[:92]:92s}" + "".join(f"{c[k]:9d}" for k in COLS))
