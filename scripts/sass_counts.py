#!/usr/bin/env python
"""Static SASS mnemonic counts per kernel of the built library -> markdown (profiles/r02_sass_counts.md).

    python scripts/sass_counts.py clonealign_b200/libclonealign_b200.so [kernel-name-regex] > profiles/r02_sass_counts.md
"""
import collections
import re
import subprocess
import sys

COLS = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "IMMA", "LDSM", "FFMA2", "FADD2", "DFMA", "DADD", "MUFU", "PRMT", "SHFL", "LDS", "STS", "BAR"]


def main():
    lib = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per, tot, cur = collections.OrderedDict(), collections.Counter(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            cur = per.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["instr"] += 1
            cur[op] += 1
            tot[op] += 1
    print(f"# SASS instruction mnemonics per kernel of `{lib.split('/')[-1]}`\n")
    print("`cuobjdump -sass`, static counts per function (not executed counts).  `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTMALDG` = TMA tensor load,")
    print("`UBLKCP` = cp.async.bulk (TMA engine, 1-D), `LDGSTS` = cp.async, `SYNCS` = mbarrier ops, `IMMA` = mma.sync u8 x s8, `LDSM` = ldmatrix, `FFMA2` / `FADD2` = packed fp32 pairs.\n")
    print("| kernel | instr | " + " | ".join(COLS) + " |")
    print("|---|---:|" + "---:|" * len(COLS))
    for name, c in per.items():
        if pat and not pat.search(name):
            continue
        print(f"| `{name}` | {c['instr']} | " + " | ".join(str(c[k]) for k in COLS) + " |")
    print("\nWhole library: " + ", ".join(f"{k} {tot[k]}" for k in COLS))


if __name__ == "__main__":
    main()
