#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show what a kernel is made of (cuobjdump -sass of
rasr_b200/lib/librasr_b200.so; runs in the build container, no GPU needed) -> profiles/sass_summary.md

  UTCHMMA / UTCQMMA  tcgen05.mma        LDTM / STTM  tcgen05.ld / st (TMEM)     UTMALDG / UTMASTG  TMA tensor load / store
  UBLKCP             cp.async.bulk      FFMA2 / FADD2 / FMUL2  packed f32x2     IMMA / HMMA        warp-level mma.sync
  SYNCS              mbarrier ops       LDSM  ldmatrix                           REDUX / SHFL       warp reductions / shuffles
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rasr_b200", "lib", "librasr_b200.so")
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "IMMA", "IDP",
        "HMMA", "LDSM", "LDS", "LDG", "STG", "REDUX", "SHFL", "MUFU", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k in KEYS:
                if op == k or (k in ("LDS", "LDG", "STG", "SHFL", "MUFU", "BAR", "SYNCS", "REDUX", "LDSM") and op.startswith(k)):
                    counts[cur][k] += 1
    names = demangle(order)
    lines = ["# SASS summary of rasr_b200/lib/librasr_b200.so (sm_100a)", "",
             "`python scripts/sass_summary.py` (cuobjdump -sass, counted per kernel; columns with no hit anywhere are "
             "dropped).  UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = "
             "cp.async.bulk, SYNCS = mbarrier, FFMA2 / FADD2 = packed f32x2, IMMA = mma.sync u8, IDP = dp4a.", ""]
    used = [k for k in KEYS if any(counts[f][k] for f in order)]
    lines.append("| kernel | instr | " + " | ".join(used) + " |")
    lines.append("|---|---|" + "---|" * len(used))
    for f in sorted(order, key=lambda f: names[f]):
        short = re.sub(r"\(anonymous namespace\)::", "", names[f])
        short = re.sub(r"\(.*", "", short)[:110]
        lines.append("| `%s` | %d | " % (short, counts[f]["_total"]) + " | ".join(str(counts[f][k] or "") for k in used) + " |")
    text = "\n".join(lines) + "\n"
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_summary.md")
    open(out, "w").write(text)
    print("wrote %s (%d kernels)" % (out, len(order)))


if __name__ == "__main__":
    main()
