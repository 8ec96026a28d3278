#!/bin/bash
# Round 2, call AF (N GPUs): cell-sharded bench line of config 3 and the multi-GPU tests with the tensor-copy integer Y pass under path = auto.
set -u
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
python -c "import __graft_entry__ as g; g.build()" > $O/r2af_build.log 2>&1 || { tail -20 $O/r2af_build.log; exit 1; }
if [ "$N" = 2 ] && [ -z "${SKIP_TESTS:-}" ]; then
  timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider > $O/r2af_tests.log 2>&1; echo "rc=$?"; tail -3 $O/r2af_tests.log | cut -c1-300
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > $O/r2af_bench_${N}gpu.json 2> $O/r2af_bench_${N}gpu.err
python - $O/r2af_bench_${N}gpu.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],4), d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], "e2e", d["e2e"] and round(d["e2e"]["value"],1))
except Exception as e:
    print("no line:", e)
PY
tail -3 $O/r2af_bench_${N}gpu.err | cut -c1-300
