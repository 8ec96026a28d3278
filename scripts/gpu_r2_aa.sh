#!/bin/bash
# Round 2, call AA (one GPU): what holds the Y pass up -- its duration next to each of the other launches of the step alone.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r2aa_build.log 2>&1 || { tail -20 $O/r2aa_build.log; exit 1; }
CA_BENCH_ABLATE=1 timeout 200 python bench.py --steps 30 --warmup 5 --quick --no-e2e --no-cpu-baseline > $O/r2aa_bench.json 2> $O/r2aa_bench.err
grep ablation_ms $O/r2aa_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())['ablation_ms']
for k,v in d.items(): print(k, {n:x for n,x in v.items()})
"
