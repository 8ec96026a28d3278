#!/bin/bash
# Round 2, call B (one GPU): strict parity suite (new node-sum kernels, ypass4 / cosched variants), bench of the default
# kernel set and of the ypass4 candidates, ncu of the co-scheduled set.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2b_build.log 2>&1 || { tail -20 $O/r2b_build.log; exit 1; }
echo "== 1. GPU suite (strict)"
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r2b_tests.log 2>&1; echo "rc=$?"; tail -15 $O/r2b_tests.log
echo "== 2. bench, default path"
timeout 600 python bench.py --steps 30 --warmup 5 > $O/r2b_bench.json 2> $O/r2b_bench.err; tail -c 3500 $O/r2b_bench.json; tail -3 $O/r2b_bench.err
for V in "ypass4,epi2,lean,defer" "ypass4,epi2,lean,defer,cosched" "ypass3,epi2,lean,defer,overlap"; do
  echo "== 3. bench --variants $V"
  timeout 300 python bench.py --path interp --variants "$V" --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline > "$O/r2b_bench_$V.json" 2> "$O/r2b_bench_$V.err"
  python - "$O/r2b_bench_$V.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], d["config"]["parity"]["sampled_cell_check"])
except Exception as e:
    print("no line:", e)
PY
  tail -2 "$O/r2b_bench_$V.err"
done
echo "== 4. ncu full capture, co-scheduled set"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ypass_k1_v4|k_interp_nodes2|k_cell_fused|k_interp_coeffs2|k_gene_fused|k_prologue|k_adam_all' -s 40 -c 10 -o $O/r2b_prof -f \
  python bench.py --path interp --variants ypass4,epi2,lean,defer --steps 2 --warmup 3 --quick --no-e2e --no-cpu-baseline > $O/r2b_ncu_full.log 2>&1
tail -2 $O/r2b_ncu_full.log | cut -c1-300
ls -la $O | grep r2b
