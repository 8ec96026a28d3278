#!/bin/bash
# Round 2, call F (TWO GPUs): multi-GPU tests (MultiSession, run_clonealign over devices), sharding parity + timing of
# the bench at N = 2 with NCCL and with the peer-memory all-reduce kernel.  Short inner timeouts (a hang is charged 2x).
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2f_build.log 2>&1 || { tail -20 $O/r2f_build.log; exit 1; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== 1. multi-GPU tests"
timeout 420 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider > $O/r2f_tests.log 2>&1; echo "rc=$?"; tail -12 $O/r2f_tests.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N", d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_start"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], d["config"]["parity"]["sampled_cell_check"]["ok"], "e2e", d["e2e"] and (round(d["e2e"]["value"],1), d["e2e"]["seconds"]))
except Exception as e:
    print("no line:", e)
PY
}
NG=${NG:-2}
for V in "" "p2p"; do
  echo "== bench N=$NG variants '$V'"
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $NG --steps 30 --warmup 5 --quick --no-cpu-baseline --variants "$V" > "$O/r2f_bench_${NG}_$V.json" 2> "$O/r2f_bench_${NG}_$V.err"
  summ "$O/r2f_bench_${NG}_$V.json"; tail -3 "$O/r2f_bench_${NG}_$V.err"
done
ls -la $O | grep r2f
