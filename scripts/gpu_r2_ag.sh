#!/bin/bash
# Round 2, call AG (one GPU): the forced tensor-copy Y-pass tests + the default bench line against the committed traffic file.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2ag_build.log 2>&1 || { tail -20 $O/r2ag_build.log; exit 1; }
timeout 400 python -m pytest tests/test_interp_gpu.py -m gpu -q -p no:cacheprovider -k "tensor_copy" > $O/r2ag_tests.log 2>&1; echo "rc=$?"; tail -5 $O/r2ag_tests.log | cut -c1-400
timeout 600 python bench.py > $O/r2ag_bench.json 2> $O/r2ag_bench.err; tail -c 300 $O/r2ag_bench.json; tail -2 $O/r2ag_bench.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2ag_bench.json').read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],4), d["roofline"]["frac"], d["roofline"]["traffic"], d["roofline"]["traffic_source"], d["e2e"]["value"], d["dtype"][:60])
PY
