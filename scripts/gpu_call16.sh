#!/bin/bash
# GPU call 16 (1 GPU): fixtures of the polar / Mercator ingest from the reference, full parity suite, batch64 bench, build-kernel capture
set -x
mkdir -p gpurun_out
python tests/golden/make_golden.py gpurun_out/golden --ingest-only > gpurun_out/make_golden_ingest2.log 2>&1
cp gpurun_out/golden/gridnav_*.npz tests/golden/ 2>/dev/null
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
python bench.py --workload batch64 --steps 3 --warmup 2 > gpurun_out/bench_batch64_s4.json 2> gpurun_out/bench_batch64_s4.err
python bench.py --workload batch64 --steps 3 --warmup 2 --streams 1 > gpurun_out/bench_batch64_s1.json 2> gpurun_out/bench_batch64_s1.err
python bench.py --workload batch64 --steps 3 --warmup 2 --streams 8 > gpurun_out/bench_batch64_s8.json 2> gpurun_out/bench_batch64_s8.err
export OCTANE_NO_GRAPHS=1
B="python bench.py --workload conus --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_build -s 54 -c 4 -f -o gpurun_out/prof_build_v6 $B > gpurun_out/ncu_build.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/make_golden_ingest2.log | tail -5; cut -c1-300 gpurun_out/bench_batch64_s*.json; tail -3 gpurun_out/bench_batch64_s4.err
