#!/bin/bash
# Round 2, call AI (one GPU): the default bench line (e2e = warm-up fit + median of three fits).
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2ai_build.log 2>&1 || { tail -20 $O/r2ai_build.log; exit 1; }
timeout 600 python bench.py > $O/r2ai_bench.json 2> $O/r2ai_bench.err; tail -2 $O/r2ai_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ai_bench.json').read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["frac"], d["step_hbm"]["frac"], d["e2e"]["value"], d["e2e"]["seconds"], d["e2e"]["first_fit_seconds_total"], d["e2e"]["seconds_total_all"], d["gpu_launches"])
PY
