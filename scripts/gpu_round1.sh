#!/bin/bash
# First GPU pass: parity tests per path (separate processes so a faulting kernel cannot poison the rest),
# sanitizer on a small case, then short benches.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== cudacore tests" 
CLONEALIGN_B200_PATH=cudacore timeout 900 python -m pytest tests -m gpu -q -k "not tensor and not full_size" -p no:cacheprovider > gpurun_out/t_cudacore.log 2>&1
tail -5 gpurun_out/t_cudacore.log
echo "== tensor tests"
timeout 900 python -m pytest tests -m gpu -q -k "tensor" -p no:cacheprovider > gpurun_out/t_tensor.log 2>&1
tail -5 gpurun_out/t_tensor.log
echo "== sanitizer (cudacore, small)"
CLONEALIGN_B200_PATH=cudacore timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -k "ragged and cudacore and 130" -p no:cacheprovider > gpurun_out/sanitizer_cudacore.log 2>&1
tail -3 gpurun_out/sanitizer_cudacore.log
echo "== sanitizer (tensor, small)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -k "ragged and tensor and 130" -p no:cacheprovider > gpurun_out/sanitizer_tensor.log 2>&1
tail -3 gpurun_out/sanitizer_tensor.log
echo "== bench c2 cudacore"
timeout 600 python bench.py --config c2 --path cudacore --steps 20 --no-cpu-baseline > gpurun_out/bench_c2_cudacore.json 2> gpurun_out/bench_c2_cudacore.err
tail -c 1500 gpurun_out/bench_c2_cudacore.json
echo "== bench c3 cudacore"
timeout 900 python bench.py --config c3 --path cudacore --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c3_cudacore.json 2> gpurun_out/bench_c3_cudacore.err
tail -c 1500 gpurun_out/bench_c3_cudacore.json
echo "== bench c3 tensor"
timeout 900 python bench.py --config c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3_tensor.json 2> gpurun_out/bench_c3_tensor.err
tail -c 2500 gpurun_out/bench_c3_tensor.json; tail -5 gpurun_out/bench_c3_tensor.err
echo "== full size test"
timeout 900 python -m pytest tests -m gpu -q -k "full_size" -p no:cacheprovider > gpurun_out/t_fullsize.log 2>&1
tail -5 gpurun_out/t_fullsize.log
echo "== smoke"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
