#!/bin/bash
# Round 2, call D (one GPU): bulk-copy Y pass (ypass4) + co-scheduling experiments, coeffs2 fix.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2d_build.log 2>&1 || { tail -20 $O/r2d_build.log; exit 1; }
echo "== 0. ypass4 in its own process"
timeout 300 python -m pytest tests/test_interp_gpu.py -m gpu -q -p no:cacheprovider -k "ypass4" > $O/r2d_y4.log 2>&1; Y4=$?; echo "rc=$Y4"; tail -4 $O/r2d_y4.log
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], d["config"]["parity"]["sampled_cell_check"]["ok"])
except Exception as e:
    print("no line:", e)
PY
}
run() { # label, env, variants
  echo "== bench [$1] env: $2 variants: $3"
  env $2 timeout 300 python bench.py --path interp --variants "$3" --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline > "$O/r2d_$1.json" 2> "$O/r2d_$1.err"
  summ "$O/r2d_$1.json"; tail -2 "$O/r2d_$1.err"
}
run default "X=1" "ypass3,epi2,lean,defer"
if [ $Y4 -eq 0 ]; then
run y4_64 "CLONEALIGN_B200_Y4_MINB=4" "ypass4,epi2,lean,defer"
run y4_80 "CLONEALIGN_B200_Y4_MINB=3" "ypass4,epi2,lean,defer"
run co_64_w16 "CLONEALIGN_B200_Y4_MINB=4" "ypass4,epi2,lean,defer,cosched"
run co_80_w12 "CLONEALIGN_B200_Y4_MINB=3" "ypass4,epi2,lean,defer,cosched"
run co_64_w12 "CLONEALIGN_B200_Y4_MINB=4 CLONEALIGN_B200_FUSED_WARPS=12" "ypass4,epi2,lean,defer,cosched"
run co_64_w8 "CLONEALIGN_B200_Y4_MINB=4 CLONEALIGN_B200_FUSED_WARPS=8" "ypass4,epi2,lean,defer,cosched"
echo "== ncu ypass4 (64 regs)"
CLONEALIGN_B200_Y4_MINB=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ypass_k1_v4|k_interp_coeffs2' -s 8 -c 3 -o $O/r2d_prof_y4 -f \
  python bench.py --path interp --variants ypass4,epi2,lean,defer --steps 2 --warmup 3 --quick --no-e2e --no-cpu-baseline > $O/r2d_ncu_y4.log 2>&1
fi
ls -la $O | grep r2d
