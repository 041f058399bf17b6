#!/bin/bash
# GPU call: parity tests, default bench (full disk, 1 GPU), CONUS bench + ncu launch list + ncu --set full of the PCG kernels
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_fulldisk.json 2> gpurun_out/bench_fulldisk.err
python bench.py --workload conus --steps 3 --warmup 3 > gpurun_out/bench_conus.json 2> gpurun_out/bench_conus.err
python bench.py --workload meso --steps 5 --warmup 3 --ref-cuda > gpurun_out/bench_meso.json 2> gpurun_out/bench_meso.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
export OCTANE_NO_GRAPHS=1
B="python bench.py --workload conus --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_conus.csv $B > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pcg_pass1_tma -s 830 -c 2 -f -o gpurun_out/prof_pass1_tma $B > gpurun_out/ncu_p1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pcg_pass2 -s 830 -c 2 -f -o gpurun_out/prof_pass2 $B > gpurun_out/ncu_p2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:^k_build$' -s 28 -c 1 -f -o gpurun_out/prof_build $B > gpurun_out/ncu_build.log 2>&1
ls -la gpurun_out
