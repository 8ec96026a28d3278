#!/bin/bash
# Round 2, call P (one GPU): cell2 set: derivative columns from the interpolants (node sums over S*C columns, one coefficient table, Rx [N][S*C]).
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2p_build.log 2>&1 || { tail -20 $O/r2p_build.log; exit 1; }
echo "== 1. GPU suite (strict)"
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2p_tests.log 2>&1; echo "rc=$?"; tail -12 $O/r2p_tests.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], d["config"]["parity"]["sampled_cell_check"]["ok"], "timeline", d["roofline"].get("timeline_ms"))
except Exception as e:
    print("no line:", e)
PY
}
echo "== 2. bench, default (cell2)"
timeout 600 python bench.py --steps 30 --warmup 5 > $O/r2p_bench.json 2> $O/r2p_bench.err; summ $O/r2p_bench.json; tail -3 $O/r2p_bench.err
echo "== 2b. bench, previous per-cell kernel"
CLONEALIGN_B200_NO_CELL2=1 timeout 300 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline > $O/r2p_bench_nocell2.json 2> $O/r2p_bench_nocell2.err; summ $O/r2p_bench_nocell2.json; tail -3 $O/r2p_bench_nocell2.err
echo "== 3. ncu launch list + full capture, default set"
CLONEALIGN_B200_NO_GRAPH=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 150 -c 40 --csv --log-file $O/r2p_launches.csv \
  python bench.py --steps 2 --warmup 3 --quick --no-e2e --no-cpu-baseline > $O/r2p_ncu_launch.log 2>&1
CLONEALIGN_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 150 -c 12 -o $O/r2p_prof -f \
  python bench.py --steps 2 --warmup 3 --quick --no-e2e --no-cpu-baseline > $O/r2p_ncu_full.log 2>&1
tail -2 $O/r2p_ncu_full.log | cut -c1-300
ls -la $O | grep r2p
