#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove what a kernel runs on (B200_PROFILING.md):
tcgen05 MMA (UTCHMMA / UTCQMMA / UTCIMMA / UTCOMMA, .2CTA for cta_group::2), TMEM loads (LDTM), TMA (UTMALDG tensor
loads, UTMASTG tensor stores, UBLKCP bulk copies), warp-level tensor cores (HMMA), packed fp32 (FFMA2), mbarrier (SYNCS),
fp64 (DFMA).

    python tools/sass_summary.py [> profiles/r2_sass_summary.txt]

Reads soft_contrastive_learning_b200/libscl_b200.so with `cuobjdump -sass` (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "soft_contrastive_learning_b200", "libscl_b200.so")
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA",
             "FFMA2", "FFMA", "DFMA", "LDS", "STS", "LDG", "STG", "ATOM", "RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        funcs[cur][base] += 1
        funcs[cur]["_total"] += 1
        if base == "UTCHMMA" and ".2CTA" in op:
            funcs[cur]["UTCHMMA.2CTA"] += 1
    names = list(funcs)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    print(f"SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass; static instruction counts)")
    print("columns: " + " ".join(MNEMONICS) + " | total")
    for n, d in sorted(zip(names, dem), key=lambda t: t[1]):
        c = funcs[n]
        d = re.sub(r"\(.*", "", d)
        if not d.startswith(("void scl::", "scl::")):
            continue
        cols = " ".join(f"{m}={c[m]}" for m in MNEMONICS if c[m])
        print(f"{d[:100]:<100} {cols} | {c['_total']}")


if __name__ == "__main__":
    sys.exit(main())
