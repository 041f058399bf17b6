#!/bin/bash
# GPU call 5 (1 GPU): bit-identity of the new build / pass-1 kernels against the previous
# library build, parity tests, A/B benches (old vs new library; build occupancy 2 vs 3)
set -x
mkdir -p gpurun_out
OCTANE_B200_LIB=$PWD/build/liboctane_b200_old.so python scripts/gpu_hash.py > gpurun_out/hash_old.txt 2>&1
python scripts/gpu_hash.py > gpurun_out/hash_new.txt 2>&1
OCTANE_BUILD_OCC=2 python scripts/gpu_hash.py > gpurun_out/hash_new_occ2.txt 2>&1
diff gpurun_out/hash_old.txt gpurun_out/hash_new.txt > gpurun_out/hash_diff.txt 2>&1; echo "diff rc=$?" >> gpurun_out/hash_diff.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
python bench.py --workload conus --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_conus_v4.json 2> gpurun_out/bench_conus_v4.err
OCTANE_BUILD_OCC=2 python bench.py --workload conus --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_conus_v4_occ2.json 2> gpurun_out/bench_conus_v4_occ2.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fulldisk_v4.json 2> gpurun_out/bench_fulldisk_v4.err
python bench.py --workload meso --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_meso_v4.json 2> gpurun_out/bench_meso_v4.err
export OCTANE_NO_GRAPHS=1
B="python bench.py --workload conus --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_build' -s 27 -c 4 -f -o gpurun_out/prof_build_v4 $B > gpurun_out/ncu_build.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pcg_pass1_tma -s 830 -c 1 -f -o gpurun_out/prof_pass1_tma_v4 $B > gpurun_out/ncu_p1.log 2>&1
ls -la gpurun_out
