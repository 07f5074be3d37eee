#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_x7.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_em.py --deselect tests/test_gpu_model.py --deselect tests/test_gpu_layout.py 2>&1 | grep -v "^frame\|omitting" | tail -60 > $L
cat $L | cut -c1-400
