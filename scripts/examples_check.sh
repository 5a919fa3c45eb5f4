#!/bin/bash
# the two example scripts on a GPU: a 2x2x2x2 surrogate, then metadata -> 2 slices -> merge on a 3x2x2x2 grid
mkdir -p gpurun_out/ex
{
python examples/make_surrogate.py gpurun_out/ex/t.jld --nsims 20 &&
python examples/make_surrogate_slices.py metadata gpurun_out/ex 3 2 2 2 &&
python examples/make_surrogate_slices.py slice gpurun_out/ex 1 2 &&
python examples/make_surrogate_slices.py slice gpurun_out/ex 2 2 &&
python examples/make_surrogate_slices.py merge gpurun_out/ex 2
} > gpurun_out/r2_examples_check.log 2>&1
echo "rc=$?"; cat gpurun_out/r2_examples_check.log
