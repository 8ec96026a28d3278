#!/bin/bash
# full GPU test suite + c3 benches (f32 and auto storage)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/t_all.log 2>&1
grep -E "^E  .*(mismatch|assert)|passed|failed|^FAILED" gpurun_out/t_all.log | cut -c1-600 | head -30
EXPS=${EXPS:-'f32|CLONEALIGN_B200_NO_OVERLAP=1|--y-store f32
auto|CLONEALIGN_B200_NO_OVERLAP=1|--y-store auto
auto_overlap|A=1|--y-store auto'}
EXPS="$EXPS" bash scripts/gpu_exp.sh
