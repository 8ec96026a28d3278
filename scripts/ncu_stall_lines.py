#!/usr/bin/env python
"""Warp-state samples of one kernel from an `ncu --set full --import-source on` capture, per SASS instruction: the top lines with their
main stall reasons, the samples per opcode, and the share of samples in front of the first LDSM (for the producer / consumer Y pass:
per-tile set-up + producer loop) against the stage loop.  Used for profiles/r02_ypass7_stalls.md.

    python scripts/ncu_stall_lines.py gpurun_out/r2ae_prof.ncu-rep k_ypass_k1_v7 [top_n]
"""
import csv
import io
import subprocess
import sys
from collections import Counter


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 14
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kernel}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))   # first launch only
    h, data = rows[hi], [r for r in rows[hi + 1:end] if len(r) > 5 and r[0].startswith("0x")]
    ix = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith("stall_") and "Not" not in k]
    n = lambda r: int(r[ix["# Samples"]])
    total = sum(n(r) for r in data)
    ops = [r[1].strip() for r in data]
    opname = lambda o: (o.split()[1] if o.startswith("@") else o.split()[0]).split(".")[0]
    print(f"# Warp-state samples of `{kernel}` ({rep.split('/')[-1]}, first captured launch): {total} samples\n")
    first = next((i for i, o in enumerate(ops) if "LDSM" in o), None)
    if first is not None:
        last = max(i for i, o in enumerate(ops) if "IMMA" in o)
        a, b = max(first - 30, 0), min(last + 40, len(data))
        print(f"In front of the stage loop (set-up of a tile + producer loop): {sum(n(r) for r in data[:a])} samples; stage loop of the "
              f"consumers: {sum(n(r) for r in data[a:b])}; behind it (tile epilogue): {sum(n(r) for r in data[b:])}.\n")
    print("| samples | executed | instruction | main stall reasons |\n|---:|---:|---|---|")
    for r in sorted(data, key=lambda r: -n(r))[:top_n]:
        s = sorted(((k, int(r[ix[k]])) for k in stalls if int(r[ix[k]]) > 0), key=lambda kv: -kv[1])[:3]
        print(f"| {n(r)} | {r[ix['Instructions Executed']]} | `{r[1].strip()[:64]}` | {', '.join(f'{k[6:]} {v}' for k, v in s)} |")
    c, e = Counter(), Counter()
    for r, o in zip(data, ops):
        c[opname(o)] += n(r)
        e[opname(o)] += int(r[ix["Instructions Executed"]])
    print("\n| opcode | samples | warp instructions executed |\n|---|---:|---:|")
    for k, v in c.most_common(12):
        print(f"| {k} | {v} | {e[k]} |")


if __name__ == "__main__":
    main()
