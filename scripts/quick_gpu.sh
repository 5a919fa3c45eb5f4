#!/bin/bash
# quick GPU check after a kernel change: mirror/api tests, regime probe, C2 bench without the CPU leg
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python scripts/perf_probe.py > gpurun_out/perf_probe.log 2>&1; cat gpurun_out/perf_probe.log
timeout 300 python scripts/c2_breakdown.py > gpurun_out/c2_breakdown.log 2>&1; tail -3 gpurun_out/c2_breakdown.log
