#!/bin/bash
# round evidence: default bench line, reference arm, ncu launch list + full capture of the heavy kernels
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python bench.py > gpurun_out/BENCH_default.json 2> gpurun_out/BENCH_default.err
tail -c 600 gpurun_out/BENCH_default.json; echo
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/BENCH_reference.json 2> gpurun_out/BENCH_reference.err
tail -c 700 gpurun_out/BENCH_reference.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/launches_c3.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_expgemm_tc|k_ypass_k1|k_gene_grads|k_cell_epilogue" -s 11 -c 5 \
   -o gpurun_out/prof_c3 -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
