#!/bin/bash
# quick iteration: tensor-path tests + c3 bench (no e2e / cpu baseline)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
timeout 900 python -m pytest tests -m gpu -q -k "tensor or full_size or device_rng or seed or storage" -p no:cacheprovider > gpurun_out/t_tensor.log 2>&1
grep -E "^E  .*(mismatch|assert)|passed|failed" gpurun_out/t_tensor.log | cut -c1-900 | head -30
timeout 900 python bench.py --config c3 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline ${BENCH_EXTRA:-} > gpurun_out/bench_c3_quick.json 2> gpurun_out/bench_c3_quick.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_c3_quick.json').read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], "clocks", d["clocks"])
    print("kernels", d["roofline"]["all_kernels_ms"])
    print("per_kernel", {k:(round(v["achieved"],1), round(v["frac"],3)) for k,v in d["roofline"]["per_kernel"].items()})
    print("step_hbm", d["step_hbm"])
except Exception as e:
    print("bench parse failed", e); print(open('gpurun_out/bench_c3_quick.err').read()[-2000:])
PY
