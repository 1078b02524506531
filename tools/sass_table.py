#!/usr/bin/env python
"""SASS mnemonic counts per kernel of libgcl_b200.so (cuobjdump -sass), the table kept as profiles/r0N_sass_mnemonics.txt.

  python tools/sass_table.py > profiles/r02_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "graphcast-lite_b200", "gcl_b200", "libgcl_b200.so")
COLS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "FFMA2", "SYNCS", "HMMA", "ATOM", "RED"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    res = {}
    for m, d in zip(names, out):
        d = re.sub(r"^void ", "", d)
        d = d.replace("(anonymous namespace)::", "").replace("gcl::", "")
        depth, cut = 0, len(d)
        for i, ch in enumerate(d):                      # cut the parameter list, keep template arguments
            if ch == "<":
                depth += 1
            elif ch == ">":
                depth -= 1
            elif ch == "(" and depth == 0:
                cut = i
                break
        res[m] = d[:cut]
    return res


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if not m:
            continue
        op = m.group(1)
        for c in COLS:
            if op == c or (c in ("ATOM", "RED") and re.fullmatch(c + r"[GS]?", op)) or (c == "UTCHMMA" and op.startswith("UTCHMMA")):
                counts[cur][c] += 1
    names = demangle(list(counts))
    print("# SASS mnemonic counts per kernel (cuobjdump -sass libgcl_b200.so, sm_100a), round 2.")
    print("# UTC*MMA = tcgen05.mma (with the A operand in shared OR tensor memory), LDTM / STTM = tcgen05.ld / tcgen05.st,")
    print("# UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = cp.async.bulk, LDGSTS = cp.async, FFMA2 = fma.rn.f32x2 (packed")
    print("# fp32 FMA), SYNCS = mbarrier ops.  No HMMA (legacy mma.sync); ATOM / RED lines are integer operations (CSR builder")
    print("# counters, shared-memory bookkeeping of the tcgen05 kernels) -- no floating-point atomics anywhere.")
    print("kernel," + ",".join(COLS))
    for m, c in sorted(counts.items(), key=lambda kv: names[kv[0]]):
        print('"' + names[m] + '",' + ",".join(str(c[k]) for k in COLS))


if __name__ == "__main__":
    sys.exit(main())
