#!/bin/bash
# ncu captures: (1) launch list with per-launch device time, (2) --set full on the heavy kernels.
set -u
mkdir -p gpurun_out
CFG=${CFG:-c3}
EXTRA=${EXTRA:-}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${CFG}.csv \
   python bench.py --config $CFG --steps 2 --warmup 1 --no-e2e --no-cpu-baseline $EXTRA > gpurun_out/ncu_launch_${CFG}.log 2>&1
tail -2 gpurun_out/ncu_launch_${CFG}.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_expgemm_tc|k_ypass_k1|k_gene_grads|k_cell_epilogue" -s 14 -c 6 \
   -o gpurun_out/prof_${CFG} -f python bench.py --config $CFG --steps 2 --warmup 1 --no-e2e --no-cpu-baseline $EXTRA > gpurun_out/ncu_full_${CFG}.log 2>&1
tail -3 gpurun_out/ncu_full_${CFG}.log
ls -la gpurun_out/*.ncu-rep
