#!/bin/bash
# One fat GPU call: smoke, GPU parity tests, short bench, sanitizer. Everything lands in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
if ! grep -q "smoke OK" gpurun_out/smoke.log; then echo "smoke failed: stopping"; exit 1; fi
echo "== pytest gpu"; timeout 420 python -m pytest tests -q -m gpu -x -p no:cacheprovider --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== bench small"; timeout 240 python bench.py --frames-per-gpu 256 --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
if [ "$1" == "sanitize" ]; then
echo "== sanitizer"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -15 gpurun_out/sanitizer.log
fi
