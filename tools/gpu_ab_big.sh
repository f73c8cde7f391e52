cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for tag in default bu2; do
  if [ "$tag" = default ]; then unset PWPP_LIB; else export PWPP_LIB=$PWD/tools/_build/libpwpp_b200_$tag.so; fi
  timeout 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --sensor dense1m --frames-per-gpu 32 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/ab.json')); print('$tag', round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
done
PWPP_LIB=$PWD/tools/_build/libpwpp_b200_bu2.so timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py -m gpu -x -q -p no:cacheprovider -k "dense or edge" 2>&1 | tail -3
