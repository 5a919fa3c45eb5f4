#!/bin/bash
# Confirmation of the committed library on one B200: the bit-exact suites, smoke(), a short C2 bench line.
mkdir -p gpurun_out
timeout 40 python -m pytest tests -m gpu -x -q --ignore=tests/test_gpu_parity.py > gpurun_out/r2_confirm_pytest.log 2>&1; echo "suite rc=$?"; tail -n 2 gpurun_out/r2_confirm_pytest.log
timeout 15 python __graft_entry__.py --smoke > gpurun_out/r2_confirm_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r2_confirm_smoke.log
timeout 30 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2_confirm_bench.json 2> gpurun_out/r2_confirm_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2_confirm_bench.json
