#!/bin/bash
# Round 2, call C (one GPU): ypass4 smoke in its own process, strict parity suite, benches, ncu of the default set.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2c_build.log 2>&1 || { tail -20 $O/r2c_build.log; exit 1; }
echo "== 0. ypass4 in its own process"
timeout 300 python -m pytest tests/test_interp_gpu.py -m gpu -q -p no:cacheprovider -k "ypass4_is_bit_identical" > $O/r2c_y4.log 2>&1; Y4=$?; echo "rc=$Y4"; tail -4 $O/r2c_y4.log
DESEL=""
if [ $Y4 -ne 0 ]; then DESEL="-k not(ypass4)"; fi
echo "== 1. GPU suite (strict) $DESEL"
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider $DESEL > $O/r2c_tests.log 2>&1; echo "rc=$?"; tail -15 $O/r2c_tests.log
echo "== 2. bench, default path"
timeout 600 python bench.py --steps 30 --warmup 5 > $O/r2c_bench.json 2> $O/r2c_bench.err; tail -c 1500 $O/r2c_bench.json; tail -3 $O/r2c_bench.err
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], d["config"]["parity"]["sampled_cell_check"]["ok"])
except Exception as e:
    print("no line:", e)
PY
}
summ $O/r2c_bench.json
if [ $Y4 -eq 0 ]; then
for V in "ypass4,epi2,lean,defer" "ypass4,epi2,lean,defer,cosched"; do
  echo "== 3. bench --variants $V"
  timeout 300 python bench.py --path interp --variants "$V" --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline > "$O/r2c_bench_$V.json" 2> "$O/r2c_bench_$V.err"
  summ "$O/r2c_bench_$V.json"; tail -2 "$O/r2c_bench_$V.err"
done
fi
echo "== 4. ncu launch list + full capture, default set"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 150 -c 40 --csv --log-file $O/r2c_launches.csv \
  python bench.py --steps 2 --warmup 3 --quick --no-e2e --no-cpu-baseline > $O/r2c_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 150 -c 11 -o $O/r2c_prof -f \
  python bench.py --steps 2 --warmup 3 --quick --no-e2e --no-cpu-baseline > $O/r2c_ncu_full.log 2>&1
tail -2 $O/r2c_ncu_full.log | cut -c1-300
if [ $Y4 -eq 0 ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ypass_k1_v4' -s 8 -c 1 -o $O/r2c_prof_y4 -f \
  python bench.py --path interp --variants ypass4,epi2,lean,defer --steps 2 --warmup 3 --quick --no-e2e --no-cpu-baseline > $O/r2c_ncu_y4.log 2>&1
fi
ls -la $O | grep r2c
