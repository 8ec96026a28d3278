#!/bin/bash
# Round 2, call AD (one GPU): tensor-copy integer Y pass under path = auto (CLONEALIGN_B200_Y7=1) against the default on every bench config.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2ad_build.log 2>&1 || { tail -20 $O/r2ad_build.log; exit 1; }
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["all_kernels_ms"].get("ypass"), d["config"]["parity"]["elbo_after"], d["config"]["parity"]["hard_calls_sha256"])
except Exception as e:
    print("no line:", e)
PY
}
for c in c2 c4 c5 c3; do
  for y in 0 1; do
    echo "== $c Y7=$y"
    CLONEALIGN_B200_Y7=$y timeout 200 python bench.py --config $c --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline > $O/r2ad_${c}_$y.json 2> $O/r2ad_${c}_$y.err; summ $O/r2ad_${c}_$y.json; tail -2 $O/r2ad_${c}_$y.err | cut -c1-200
  done
done
echo "== GPU suite with the tensor-copy pass under auto"
CLONEALIGN_B200_Y7=1 timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2ad_tests.log 2>&1; echo "rc=$?"; tail -4 $O/r2ad_tests.log | cut -c1-300
