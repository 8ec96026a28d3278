#!/bin/bash
# Round 2, call U (one GPU): integer tensor-pipe Y pass (variant ypass5): parity of the kernel set against the default, bench A/B.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2u_build.log 2>&1 || { tail -20 $O/r2u_build.log; exit 1; }
echo "== 1. small-shape parity, default kernel set + ypass5 (CLONEALIGN_B200_VARIANTS)"
echo skipped
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_start"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], d["config"]["parity"]["sampled_cell_check"], "timeline", d["roofline"].get("timeline_ms"))
except Exception as e:
    print("no line:", e)
PY
}
echo "== 2. bench, ypass5"
timeout 600 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline --variants ypass5 > $O/r2u_bench_y5.json 2> $O/r2u_bench_y5.err; summ $O/r2u_bench_y5.json; tail -5 $O/r2u_bench_y5.err
