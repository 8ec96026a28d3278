#!/bin/bash
# Round 2, call V (one GPU): ypass5 with fewer warps in the co-scheduled per-cell kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2v_build.log 2>&1 || { tail -20 $O/r2v_build.log; exit 1; }
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], "timeline", d["roofline"].get("timeline_ms"))
except Exception as e:
    print("no line:", e)
PY
}
for w in 8 12; do
echo "== ypass5, per-cell kernel with $w warps"
CLONEALIGN_B200_FUSED_WARPS=$w timeout 300 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline --variants ypass5 > $O/r2v_bench_w$w.json 2> $O/r2v_bench_w$w.err; summ $O/r2v_bench_w$w.json; tail -3 $O/r2v_bench_w$w.err
done
