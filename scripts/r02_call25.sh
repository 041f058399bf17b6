#!/bin/bash
# round 2, GPU call 25 (1 GPU, the last 100 s of the budget): the headline-size and merged-kernel test files on the final library
O=gpurun_out/r02c25
mkdir -p $O
( time timeout 88 python -m pytest tests/test_gpu_headline.py tests/test_gpu_fused.py -m gpu -v -x -p no:cacheprovider ) > $O/pytest_headline_fused.log 2>&1
grep -c PASSED $O/pytest_headline_fused.log; grep -E "FAILED|ERROR|passed|failed" $O/pytest_headline_fused.log | tail -5
