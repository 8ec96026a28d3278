#!/bin/bash
# Round 2, call W (one GPU): ypass5 with 16 / 8 warps per CTA, ncu of the pass.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2w_build.log 2>&1 || { tail -20 $O/r2w_build.log; exit 1; }
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], "timeline", d["roofline"].get("timeline_ms"))
except Exception as e:
    print("no line:", e)
PY
}
for w in 16 8; do
echo "== ypass5, $w warps per Y-pass CTA"
CLONEALIGN_B200_Y5_WARPS=$w timeout 300 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline --variants ypass5 > $O/r2w_bench_y$w.json 2> $O/r2w_bench_y$w.err; summ $O/r2w_bench_y$w.json; tail -3 $O/r2w_bench_y$w.err
done
echo "== ncu of the default step with ypass5 (16 warps)"
CLONEALIGN_B200_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ypass -s 20 -c 2 -o $O/r2w_prof -f \
  python bench.py --steps 2 --warmup 3 --quick --no-e2e --no-cpu-baseline --variants ypass5 > $O/r2w_ncu_full.log 2>&1
tail -2 $O/r2w_ncu_full.log | cut -c1-200
