#!/bin/bash
# usage: KREGEX=... SKIP=.. COUNT=.. EXTRA="bench args" bash scripts/gpu_ncu_k.sh  -> gpurun_out/prof_k.ncu-rep
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX}" -s ${SKIP:-2} -c ${COUNT:-1} \
   -o gpurun_out/prof_k -f python bench.py --config c3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline ${EXTRA:-} > gpurun_out/ncu_k.log 2>&1
tail -2 gpurun_out/ncu_k.log | cut -c1-300
