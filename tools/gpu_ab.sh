#!/bin/bash
# smoke + GPU parity tests on the default configuration, then the A/B sweep of the kernel-variant switches.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "== smoke" ; timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
if ! grep -q "smoke OK" gpurun_out/smoke.log; then echo "smoke failed: stopping"; exit 1; fi
echo "== pytest gpu"; timeout 420 python -m pytest tests -q -m gpu -x -p no:cacheprovider --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== A/B"; timeout 300 python tools/gpu_ab.py 1024 5 > gpurun_out/ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/ab.log | cut -c1-600
