#!/bin/bash
# Round 2, call R (TWO GPUs): cell2 kernel set, NCCL and peer-memory all-reduce
# at N = 2, multi-GPU tests.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2r_build.log 2>&1 || { tail -20 $O/r2r_build.log; exit 1; }
timeout 420 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider > $O/r2r_tests.log 2>&1; echo "rc=$?"; tail -4 $O/r2r_tests.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N", d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"])
except Exception as e:
    print("no line:", e)
PY
}
NG=${NG:-2}
for V in "" "p2p"; do
  echo "== bench N=$NG variants '$V'"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $NG --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline --variants "$V" > "$O/r2r_bench_${NG}_$V.json" 2> "$O/r2r_bench_${NG}_$V.err"
  summ "$O/r2r_bench_${NG}_$V.json"; tail -2 "$O/r2r_bench_${NG}_$V.err"
done
