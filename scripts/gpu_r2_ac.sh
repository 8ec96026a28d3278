#!/bin/bash
# Round 2, call AC (one GPU): full ncu capture of the tensor-copy integer Y pass (k_ypass_k1_v7), one launch.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2ac_build.log 2>&1 || { tail -20 $O/r2ac_build.log; exit 1; }
CLONEALIGN_B200_Y7_PLAIN=1 CLONEALIGN_B200_Y5_SPEC=2 CLONEALIGN_B200_NO_GRAPH=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_ypass_k1_v7 -s 4 -c 1 -o $O/r2ac_prof -f \
  python bench.py --steps 2 --warmup 3 --quick --no-e2e --no-cpu-baseline --variants ypass5 > $O/r2ac_ncu_full.log 2>&1
tail -2 $O/r2ac_ncu_full.log | cut -c1-200
ls -la $O/r2ac_prof.ncu-rep
