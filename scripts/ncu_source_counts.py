#!/usr/bin/env python
"""Per-CUDA-source-line warp-instruction counts of one kernel from an `ncu --set full --import-source on` capture
(compile with -lineinfo).  This is the view the round-1 redesign of the Y pass / per-cell epilogue was made from
(profiles/r01_source_counts.md).

    python scripts/ncu_source_counts.py gpurun_out/r02_prof.ncu-rep k_cell_fused [units_per_launch] [top_n]

`units_per_launch` (e.g. the number of cells) turns the counts into instructions per unit.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else None
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          f"regex:{kernel}"], capture_output=True, text=True).stdout
    cur, agg = None, {}
    for r in csv.reader(io.StringIO(out)):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 8 or r[0] in ("Line No", "Function Name") or r[0] == "":
            continue
        try:
            key = (cur, int(r[0]), r[1].strip()[:100])
            agg[key] = agg.get(key, 0) + int(r[7])
        except ValueError:
            pass
    tot = sum(agg.values())
    print(f"## `{kernel}` - {tot:.4g} warp instructions" + (f" = {tot / units:.0f} per unit" if units else ""))
    print("\n| warp instr. | share | source line |\n|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        n = f"{v / units:.1f}" if units else f"{v:.3g}"
        print(f"| {n} | {100 * v / max(tot, 1):.1f} % | `{k[0]}:{k[1]}` `{k[2].replace('|', '/')}` |")


if __name__ == "__main__":
    main()
