#!/bin/bash
# run on the GPU box: full gpu test suite, output to gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "exit=$?" >> gpurun_out/pytest_gpu.log
tail -45 gpurun_out/pytest_gpu.log
