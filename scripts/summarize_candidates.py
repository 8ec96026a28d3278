#!/usr/bin/env python
"""`bench.py --selfcheck` output (one JSON line per candidate, e.g. gpurun_out/r02_candidates.jsonl) or a bench JSON line
(reads config.selfcheck.candidates) -> markdown table for profiles/.

    python scripts/summarize_candidates.py gpurun_out/r02_candidates.jsonl > profiles/r02_candidates.md
"""
import json
import sys


def rows_of(path):
    rows = []
    for ln in open(path):
        ln = ln.strip()
        if not ln.startswith("{"):
            continue
        d = json.loads(ln)
        if "candidate" in d:
            rows.append(d)
        elif "config" in d and d["config"].get("selfcheck"):
            rows += d["config"]["selfcheck"].get("candidates", [])
    return rows


def main():
    rows = rows_of(sys.argv[1])
    base = next((r for r in rows if r["candidate"][0] == "tensor" and r.get("ms_per_step")), None)
    print("| path | variants | verdict | ms / step | it/s | vs tcgen05 | ELBO dev. | worst gradient dev. | calls agree |")
    print("|---|---|---|---:|---:|---:|---:|---:|---:|")
    for r in rows:
        p, v = r["candidate"]
        ms = r.get("ms_per_step")
        d = r.get("deviation_vs_tensor_path") or {}
        g = max([x for k, x in d.items() if k.startswith("grad_")], default=None)
        print("| {} | {} | {} | {} | {} | {} | {} | {} | {} |".format(
            p, v or "—", "ok" if r.get("ok") else ("FAILED: " + r.get("error", "mismatch")[:60]),
            f"{ms:.3f}" if ms else "—", f"{1e3 / ms:.0f}" if ms else "—",
            f"{base['ms_per_step'] / ms:.2f}x" if (ms and base) else "—",
            f"{d['elbo']:.1e}" if "elbo" in d else "—", f"{g:.1e}" if g is not None else "—",
            f"{d['calls_agree']:.4f}" if "calls_agree" in d else "—"))


if __name__ == "__main__":
    main()
