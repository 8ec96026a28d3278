#!/bin/bash
# experiment runner: each line of EXPS is "label|env assignments|bench extra args"
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || tail -20 gpurun_out/build.log
run() {
  label=$1; envs=$2; extra=$3
  env $envs timeout 600 python bench.py --config ${CFG:-c3} --steps 20 --warmup 5 --no-e2e --no-cpu-baseline $extra > gpurun_out/exp_$label.json 2> gpurun_out/exp_$label.err
  python - "$label" <<'PY'
import json,sys
lab=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/exp_{lab}.json').read().strip().splitlines()[-1])
    k=d["roofline"]["all_kernels_ms"]
    print(f"{lab:28s} {d['value']:8.1f} it/s {d['ms_per_step']:7.3f} ms | fwd {k.get('lse_fwd',0):.3f} bwd {k.get('lse_bwd',0):.3f} ypass {k.get('ypass',0):.3f} epi {k.get('cell_epilogue',0):.3f} gene {k.get('gene_grads',0):.3f} | sm {d['clocks']['sm_mhz']} pw {d['clocks'].get('power_w_max')}")
except Exception as e:
    print(lab, "FAILED", e); print(open(f'gpurun_out/exp_{lab}.err').read()[-1500:])
PY
}
while IFS='|' read -r label envs extra; do
  [ -z "$label" ] && continue
  run "$label" "$envs" "$extra"
done <<< "$EXPS"
