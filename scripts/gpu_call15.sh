#!/bin/bash
# GPU call 15 (1 GPU): full parity suite, smoke, default bench (+ reference arm), ncu launch list and full captures of the final kernels
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_fulldisk_v6.json 2> gpurun_out/bench_fulldisk_v6.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_v6.json 2> gpurun_out/bench_reference_v6.err
python bench.py --workload conus --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_conus_v6.json 2> gpurun_out/bench_conus_v6.err
export OCTANE_NO_GRAPHS=1
B="python bench.py --workload conus --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_conus_v6.csv $B > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pcg_pass1_tma -s 830 -c 1 -f -o gpurun_out/prof_pass1_tma_v6 $B > gpurun_out/ncu_p1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pcg_pass2 -s 830 -c 1 -f -o gpurun_out/prof_pass2_v6 $B > gpurun_out/ncu_p2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k_build<' -s 30 -c 3 -f -o gpurun_out/prof_build_v6 $B > gpurun_out/ncu_build.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_fulldisk_v6.json | cut -c1-400
