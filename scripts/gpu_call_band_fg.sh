#!/bin/bash
# 2-GPU call: row bands with and without a first guess at world size 2 (GPUs 0,1), and, at the same time on
# GPU 1 alone, the single-GPU suite (regression check of the solve_dev refactor).
mkdir -p gpurun_out/band
( CUDA_VISIBLE_DEVICES=1 timeout 170 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_band.py \
    > gpurun_out/band/pytest_single.log 2>&1 ) &
timeout 170 python -m pytest tests/test_gpu_band.py -m gpu -q -k "2-shape0" > gpurun_out/band/pytest_band2.log 2>&1
wait
tail -n 6 gpurun_out/band/pytest_band2.log gpurun_out/band/pytest_single.log
