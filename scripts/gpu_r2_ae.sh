#!/bin/bash
# Round 2, call AE (one GPU): final evidence of the round with the tensor-copy integer Y pass under path = auto: strict suite, default bench line, reference arm, ncu launch list + full
# capture of the default train step (graph replay off for the capture).
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2ae_build.log 2>&1 || { tail -20 $O/r2ae_build.log; exit 1; }
echo "== 0. smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== 1. GPU suite (strict)"
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2ae_tests.log 2>&1; echo "rc=$?"; tail -4 $O/r2ae_tests.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], "frac", round(d["roofline"]["frac"],3), "step_hbm", round(d["step_hbm"]["frac"],3), d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], "e2e", d["e2e"] and round(d["e2e"]["value"],1), "late", d.get("late_training") and round(d["late_training"]["value"],1))
except Exception as e:
    print("no line:", e)
PY
}
echo "== 2. bench, default"
timeout 600 python bench.py > $O/r2ae_bench.json 2> $O/r2ae_bench.err; summ $O/r2ae_bench.json; tail -3 $O/r2ae_bench.err
echo "== 3. reference arm"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2ae_bench_reference.json 2> $O/r2ae_bench_reference.err; tail -c 500 $O/r2ae_bench_reference.json
echo "== 4. ncu launch list + full capture, default set"
CLONEALIGN_B200_NO_GRAPH=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 150 -c 40 --csv --log-file $O/r2ae_launches.csv \
  python bench.py --steps 2 --warmup 3 --quick --no-e2e --no-cpu-baseline > $O/r2ae_ncu_launch.log 2>&1
CLONEALIGN_B200_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 150 -c 11 -o $O/r2ae_prof -f \
  python bench.py --steps 2 --warmup 3 --quick --no-e2e --no-cpu-baseline > $O/r2ae_ncu_full.log 2>&1
tail -2 $O/r2ae_ncu_full.log | cut -c1-200
echo "== 5. the other configurations (quick lines)"
for c in c2 c4 c5; do
  timeout 200 python bench.py --config $c --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline > $O/r2ae_$c.json 2> $O/r2ae_$c.err; summ $O/r2ae_$c.json 2>/dev/null | cut -c1-200
done
