#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_expgemm_tc" -s 3 -c 2 \
   -o gpurun_out/prof_tc -f python bench.py --config c3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
