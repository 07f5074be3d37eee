#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_x8.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^frame\|omitting" | tail -30 > $L
cat $L | cut -c1-300
