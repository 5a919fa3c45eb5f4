#!/bin/bash
# One GPU call: every K2 variant library under lib/var/ through the bit-exact table/curve tests and a short C2 bench,
# then the whole -m gpu suite on the default library.  Logs under gpurun_out/.
mkdir -p gpurun_out
V=$PWD/sprfittingpaper2023.jl_b200/lib/var
for v in v0_base v1_skip v2_unroll v3_skip_unroll v4_skip_unroll_cta5 v5_skip_unroll_cta6; do
  echo "=== $v"
  if [ "$v" != v0_base ]; then
    SPRB200_LIB=$V/$v.so timeout 300 python -m pytest tests/test_gpu_mirror.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -2
  fi
  SPRB200_LIB=$V/$v.so timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({k:d.get(k) for k in ('value','ms_per_step')}), json.dumps(d.get('roofline',{}).get('kernel_ms_per_step')))"
done
echo "=== full suite (default library)"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_final.log
