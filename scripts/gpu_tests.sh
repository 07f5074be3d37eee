#!/bin/bash
# the -m gpu suite on a GPU box:  gpurun -- 'bash scripts/gpu_tests.sh'
mkdir -p gpurun_out
timeout 1500 python -W ignore -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/gpu_tests.log
