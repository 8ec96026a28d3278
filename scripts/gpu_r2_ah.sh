#!/bin/bash
# Round 2, call AH (one GPU): tensor-copy Y pass after a change: forced-on tests, parity subset, quick default bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2ah_build.log 2>&1 || { tail -20 $O/r2ah_build.log; exit 1; }
CLONEALIGN_B200_Y7=1 timeout 400 python -m pytest tests/test_interp_gpu.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "tensor_copy or auto or ypass5" > $O/r2ah_tests.log 2>&1; echo "rc=$?"; tail -3 $O/r2ah_tests.log | cut -c1-400
for i in 1 2; do
timeout 200 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline > $O/r2ah_bench_$i.json 2> $O/r2ah_bench_$i.err
python - $O/r2ah_bench_$i.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"]["ypass"], d["roofline"]["timeline_ms"]["ypass"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"])
PY
done
