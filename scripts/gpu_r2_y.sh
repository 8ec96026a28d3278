#!/bin/bash
# Round 2, call Y (one GPU): direct u8 ingest (no fp32 staging matrix): strict suite + default bench (e2e).
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2y_build.log 2>&1 || { tail -20 $O/r2y_build.log; exit 1; }
echo "== 1. GPU suite (strict)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > $O/r2y_tests.log 2>&1; echo "rc=$?"; tail -4 $O/r2y_tests.log
echo "== 2. bench, default"
timeout 600 python bench.py > $O/r2y_bench.json 2> $O/r2y_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2y_bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],4), "e2e", d["e2e"]["value"], d["e2e"]["seconds"], "f32 host", d["e2e_f32_host"]["value"], d["e2e_f32_host"]["seconds"])
PY
tail -3 $O/r2y_bench.err
