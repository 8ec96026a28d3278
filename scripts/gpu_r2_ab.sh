#!/bin/bash
# Round 2, call AB (one GPU): integer Y pass with stages as 2-D tensor copies (k_ypass_k1_v7) against the row-copy passes.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2ab_build.log 2>&1 || { tail -20 $O/r2ab_build.log; exit 1; }
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_start"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], "timeline", d["roofline"].get("timeline_ms"))
except Exception as e:
    print("no line:", e)
PY
}
echo "== small-shape parity with the tensor-copy pass"
CLONEALIGN_B200_Y5_SPEC=2 CLONEALIGN_B200_VARIANTS=ypass5 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_interp_gpu.py -m gpu -q -x -p no:cacheprovider -k "auto or ypass5" > $O/r2ab_tests.log 2>&1; echo "rc=$?"; tail -3 $O/r2ab_tests.log | cut -c1-300
echo "== ypass5, tensor copies (v7)"
CLONEALIGN_B200_Y5_SPEC=2 timeout 120 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline --variants ypass5 > $O/r2ab_bench_v7.json 2> $O/r2ab_bench_v7.err; summ $O/r2ab_bench_v7.json; tail -3 $O/r2ab_bench_v7.err
echo "== ypass5, tensor copies (v7), strided tile walk"
CLONEALIGN_B200_Y7_PLAIN=1 CLONEALIGN_B200_Y5_SPEC=2 timeout 120 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline --variants ypass5 > $O/r2ab_bench_v7p.json 2> $O/r2ab_bench_v7p.err; summ $O/r2ab_bench_v7p.json; tail -3 $O/r2ab_bench_v7p.err
