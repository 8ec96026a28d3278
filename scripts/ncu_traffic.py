#!/usr/bin/env python
""".ncu-rep (ncu --set full) -> DRAM bytes per launch of every kernel, as JSON for bench.py's `roofline.traffic`.

    python scripts/ncu_traffic.py gpurun_out/r2e_prof.ncu-rep c3 u8 > profiles/r02_traffic.json

Output: {config: {y_store: {label: {"kernel": name, "bytes": dram__bytes_read.sum + dram__bytes_write.sum, "time_us": ...}}}}
with bench.py's launch labels (ypass, cell_epilogue, ...) next to the kernel names."""
import csv
import io
import json
import subprocess
import sys

LABELS = [("k_ypass", "ypass"), ("k_cell_fused", "cell_epilogue"), ("k_prologue", "prologue"), ("k_gene_fused", "gene_grads"),
          ("k_adam_all", "adam"), ("k_interp_nodes2<1", "lse_fwd_nodes"), ("k_interp_nodes2<0", "lse_bwd_nodes"),
          ("k_interp_coeffs", "coeffs"), ("k_colpart_add", "colpart_add")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, cfg, store = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
    out = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        label = next((lab for pat, lab in LABELS if pat in name), None)
        if label is None or label in out:
            continue
        b = sum(float(r[col[k]]) * UNIT[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out[label] = {"kernel": name.split("(")[0], "bytes": b, "time_us": float(r[col["gpu__time_duration.sum"]])}
    print(json.dumps({cfg: {store: out}}, indent=1))


if __name__ == "__main__":
    main()
