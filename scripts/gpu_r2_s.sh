#!/bin/bash
# Round 2, call S (EIGHT GPUs, cell2 kernel set): scaling points N = 8 and N = 4 of the bench (cells sharded, one all-reduce per step),
# sharding parity (ELBO / hard-call hash must equal the N = 1 line).  Short inner timeouts: a hang is charged 8x.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2s_build.log 2>&1 || { tail -20 $O/r2s_build.log; exit 1; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("N", d["n_gpus"], round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_start"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], d["config"]["parity"]["sampled_cell_check"]["ok"], "e2e", d["e2e"] and (round(d["e2e"]["value"],1), d["e2e"]["seconds"]))
except Exception as e:
    print("no line:", e)
PY
}
for NG in ${NGS:-8 4}; do
  echo "== bench N=$NG"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $NG --steps 30 --warmup 5 --quick --no-cpu-baseline ${EXTRA:-} > "$O/r2s_bench_${NG}.json" 2> "$O/r2s_bench_${NG}.err"
  summ "$O/r2s_bench_${NG}.json"; tail -3 "$O/r2s_bench_${NG}.err"
done
ls -la $O | grep r2s
