#!/bin/bash
# Round 2, call T (one GPU): gamma-logit Adam inside k_cell_fused2 (A/B with CLONEALIGN_B200_NO_CELL_ADAM=1), strict suite.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2t_build.log 2>&1 || { tail -20 $O/r2t_build.log; exit 1; }
echo "== 1. GPU suite (strict)"
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2t_tests.log 2>&1; echo "rc=$?"; tail -6 $O/r2t_tests.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], d["config"]["parity"]["sampled_cell_check"]["ok"], "timeline", d["roofline"].get("timeline_ms"))
except Exception as e:
    print("no line:", e)
PY
}
echo "== 2. bench, default"
timeout 600 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline > $O/r2t_bench.json 2> $O/r2t_bench.err; summ $O/r2t_bench.json; tail -3 $O/r2t_bench.err
echo "== 2b. bench, gamma-logit Adam in k_adam_all"
CLONEALIGN_B200_NO_CELL_ADAM=1 timeout 300 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline > $O/r2t_bench_nocelladam.json 2> $O/r2t_bench_nocelladam.err; summ $O/r2t_bench_nocelladam.json; tail -3 $O/r2t_bench_nocelladam.err
