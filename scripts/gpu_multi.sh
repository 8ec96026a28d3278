#!/bin/bash
# multi-GPU bench via torchrun (NG GPUs).  EXTRA passes bench flags, e.g. the one-shot NVLink all-reduce (kernels_p2p.cuh):
#   gpurun --gpus 2 --timeout 300 -- 'NG=2 EXTRA="--path interp --variants ypass2,epi2,lean,p2p" bash scripts/gpu_multi.sh'
# (validate at NG=2 with a short timeout before any larger N: a peer that never signals leaves the kernel spinning)
set -u
NG=${NG:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout ${INNER_TIMEOUT:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus $NG --steps ${STEPS:-20} --warmup 5 --no-cpu-baseline ${EXTRA:-} > gpurun_out/bench_multi_$NG.json 2> gpurun_out/bench_multi_$NG.err
tail -c 3000 gpurun_out/bench_multi_$NG.json; echo; tail -5 gpurun_out/bench_multi_$NG.err
