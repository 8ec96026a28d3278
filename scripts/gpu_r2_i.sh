#!/bin/bash
# Round 2, call I (EIGHT GPUs): BASELINE config 5 -- 32 run_clonealign restarts of 200k x 20k x 16 spread over 8 GPUs of one
# process (replicas vs batched Y pass) -- then the N = 8 scaling point of config 3 again.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2i_build.log 2>&1 || { tail -20 $O/r2i_build.log; exit 1; }
echo "== c5: 32 restarts over 8 GPUs"
timeout 420 python bench.py --config c5 --restarts 32 --steps 20 --gpus 8 > $O/r2i_restarts_c5_8gpu.json 2> $O/r2i_restarts_c5_8gpu.err; tail -c 1500 $O/r2i_restarts_c5_8gpu.json; tail -3 $O/r2i_restarts_c5_8gpu.err
echo "== bench N=8"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 8 --steps 30 --warmup 5 --quick --no-cpu-baseline > "$O/r2i_bench_8.json" 2> "$O/r2i_bench_8.err"
python - "$O/r2i_bench_8.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N", d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"])
except Exception as e:
    print("no line:", e)
PY
tail -3 $O/r2i_bench_8.err
ls -la $O | grep r2i
