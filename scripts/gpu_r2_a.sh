#!/bin/bash
# Round 2, call A (one GPU): strict parity suite with interp as the AUTO path, bench line, ncu launch list + --set full
# capture of every kernel of the default train step at c3.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2a_build.log 2>&1 || { tail -20 $O/r2a_build.log; exit 1; }
echo "== 1. GPU suite (strict)"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > $O/r2a_tests.log 2>&1; echo "rc=$?"; tail -8 $O/r2a_tests.log
echo "== 2. bench, default path"
timeout 500 python bench.py --path auto --steps 30 --warmup 5 > $O/r2a_bench.json 2> $O/r2a_bench.err; tail -c 2500 $O/r2a_bench.json; tail -3 $O/r2a_bench.err
echo "== 3. ncu launch list + full capture (default path)"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 60 -c 60 --csv --log-file $O/r2a_launches.csv \
  python bench.py --path auto --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2a_ncu_launch.log 2>&1
tail -3 $O/r2a_launches.csv | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 60 -c 12 -o $O/r2a_prof -f \
  python bench.py --path auto --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2a_ncu_full.log 2>&1
tail -2 $O/r2a_ncu_full.log | cut -c1-300
ls -la $O | grep r2a
