#!/bin/bash
# Round 2, call Q (one GPU): cell2 set with the node-sum kernel at 4 blocks per SM; the other BASELINE configurations (c2, c4 with V = 2000, c5 one replica),
# run_clonealign restarts on one device (replicas vs batched Y pass), default bench + reference arm.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2q_build.log 2>&1 || { tail -20 $O/r2q_build.log; exit 1; }
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    if "modes" in d:
        print(d["value"], d["unit"], json.dumps(d["modes"]))
    else:
        print(d["config"]["workload"], "|", round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], "frac", round(d["roofline"]["frac"],3), "step_hbm", round(d["step_hbm"]["frac"],3), d["config"]["parity"]["sampled_cell_check"], "panels", d["config"]["panels"], "e2e", d["e2e"] and round(d["e2e"]["value"],1))
except Exception as e:
    print("no line:", e)
PY
}
for CFG in c2 c4 c5; do
  echo "== bench --config $CFG"
  timeout 500 python bench.py --config $CFG --steps 30 --warmup 5 --no-cpu-baseline > $O/r2q_bench_$CFG.json 2> $O/r2q_bench_$CFG.err; summ $O/r2q_bench_$CFG.json; tail -3 $O/r2q_bench_$CFG.err
done
echo "== c5 restarts on one device (4 restarts, 20 iterations each)"
timeout 600 python bench.py --config c5 --restarts 4 --steps 20 --gpus 1 > $O/r2q_restarts_c5_1gpu.json 2> $O/r2q_restarts_c5_1gpu.err; summ $O/r2q_restarts_c5_1gpu.json; tail -3 $O/r2q_restarts_c5_1gpu.err
echo "== default bench + reference arm"
timeout 600 python bench.py --steps 30 --warmup 5 > $O/r2q_bench_c3.json 2> $O/r2q_bench_c3.err; summ $O/r2q_bench_c3.json; tail -3 $O/r2q_bench_c3.err
ls -la $O | grep r2q
