#!/bin/bash
# Round 2, call W (one GPU): ypass5: block-barrier pass (k_ypass_k1_v5) against the mbarrier-only producer / consumer pass (k_ypass_k1_v6).
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2w_build.log 2>&1 || { tail -20 $O/r2w_build.log; exit 1; }
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_start"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], "timeline", d["roofline"].get("timeline_ms"))
except Exception as e:
    print("no line:", e)
PY
}
echo "== small-shape parity with the producer / consumer pass"
CLONEALIGN_B200_Y5_SPEC=1 CLONEALIGN_B200_VARIANTS=ypass5 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "auto" > $O/r2w_tests.log 2>&1; echo "rc=$?"; tail -3 $O/r2w_tests.log | cut -c1-300
echo "== ypass5, producer / consumer (v6)"
CLONEALIGN_B200_Y5_SPEC=1 timeout 120 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline --variants ypass5 > $O/r2w_bench_v6.json 2> $O/r2w_bench_v6.err; summ $O/r2w_bench_v6.json; tail -3 $O/r2w_bench_v6.err
echo "== ypass5, block barrier (v5, 8 warps)"
timeout 120 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline --variants ypass5 > $O/r2w_bench_v5.json 2> $O/r2w_bench_v5.err; summ $O/r2w_bench_v5.json; tail -3 $O/r2w_bench_v5.err
