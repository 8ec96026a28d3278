#!/bin/bash
# Round 2, call X (one GPU): where to start the Y pass: first in the step (cosched) or next to the per-cell kernel (defer + overlap).
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2x_build.log 2>&1 || { tail -20 $O/r2x_build.log; exit 1; }
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"], d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"], "timeline", d["roofline"].get("timeline_ms"))
except Exception as e:
    print("no line:", e)
PY
}
for V in "ypass4,epi2,lean,defer,overlap,cell2" "ypass4,epi2,lean,defer,cell2"; do
echo "== interp + $V"
timeout 300 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline --path interp --variants "$V" > "$O/r2x_bench_$V.json" 2> "$O/r2x_bench_$V.err"; summ "$O/r2x_bench_$V.json"; tail -3 "$O/r2x_bench_$V.err"
done
echo "== GPU suite (strict), ypass5 variants included"
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2x_tests.log 2>&1; echo "rc=$?"; tail -6 $O/r2x_tests.log
