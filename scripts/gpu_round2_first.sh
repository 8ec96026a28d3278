#!/bin/bash
# FIRST GPU call of round 2 (one GPU, ~6-8 minutes): validate on hardware everything that round 1 could only verify on
# the CPU emulation (tests/test_emul_parity.py) -- the interp path, the kernel variants (ypass2 / epi2 / lean), the
# device PCA, correlations and CSR ingest -- then measure every candidate and capture ncu evidence for the winner.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round2_first.sh'
# Logs land in gpurun_out/r02_*.  Each stage runs in its own process (a faulting kernel cannot poison the next one).
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/r02_build.log 2>&1 || { tail -20 $O/r02_build.log; exit 1; }
echo "== 1. default-path parity suite (must stay green)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider > $O/r02_t_default.log 2>&1; tail -3 $O/r02_t_default.log
echo "== 2. new kernels (xfail-tolerant file, run strictly here: --runxfail turns XFAIL/XPASS into real verdicts)"
timeout 900 python -m pytest tests/test_zz_interp_gpu.py -m gpu -q --runxfail -p no:cacheprovider > $O/r02_t_new.log 2>&1; tail -15 $O/r02_t_new.log
echo "== 3. memcheck on a small case, every variant set"
timeout 900 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 python -m pytest tests/test_zz_interp_gpu.py -m gpu -q --runxfail \
  -k "ragged and 130 or correlations or pca or sparse" -p no:cacheprovider > $O/r02_sanitizer.log 2>&1
echo "rc=$?"; tail -3 $O/r02_sanitizer.log
echo "== 4. candidate table at c3 (self-check child: verdict + ms/step per candidate)"
timeout 400 python bench.py --selfcheck --config c3 > $O/r02_candidates.jsonl 2> $O/r02_candidates.err; cat $O/r02_candidates.jsonl
python scripts/summarize_candidates.py $O/r02_candidates.jsonl > $O/r02_candidates.md 2>/dev/null; cat $O/r02_candidates.md
echo "== 5. default bench (picks the fastest passing candidate) + reference arm"
timeout 600 python bench.py > $O/r02_bench.json 2> $O/r02_bench.err; tail -c 3000 $O/r02_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
echo "== 6. ncu: launch list of 2 steps and one --set full capture of every kernel of the winning configuration"
PV=$(python - <<'PY'
import json
try:
    c = json.load(open("gpurun_out/r02_bench.json"))["config"]
    print(c["path"].replace("tcgen05", "auto"), c.get("variants", ""))
except Exception:
    print("auto", "")
PY
)
set -- $PV; P=${1:-auto}; V=${2:-}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 120 --csv --log-file $O/r02_launches.csv \
  python bench.py --path "$P" --variants "$V" --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ -s 40 -c 24 -o $O/r02_prof -f \
  python bench.py --path "$P" --variants "$V" --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02_ncu_full.log 2>&1
ls -la $O | tail -15
