#!/bin/bash
# One short GPU call: srsal fixtures from the reference's own kernel, our outputs for the same cases,
# timings of the two stages, then the GPU test suite (new tests first).
mkdir -p gpurun_out/golden gpurun_out/post
timeout 120 python tests/golden/make_golden.py gpurun_out/golden --srsal-only > gpurun_out/post/make_golden.log 2>&1
timeout 120 python scripts/time_post.py gpurun_out/post > gpurun_out/post/time_post.log 2>&1
cp gpurun_out/golden/srsal_*.npz tests/golden/ 2>/dev/null
timeout 200 python -m pytest tests/test_gpu_post.py tests/test_cli.py -m gpu -q > gpurun_out/post/pytest_post.log 2>&1
timeout 200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_post.py --deselect tests/test_cli.py > gpurun_out/post/pytest_rest.log 2>&1
tail -5 gpurun_out/post/*.log
