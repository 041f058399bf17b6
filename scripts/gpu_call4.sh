#!/bin/bash
# GPU call 4: strip-walk bandwidth microbenchmark (why is pass 1 slower on the full disk?),
# parity tests, benches of the current kernels, ncu launch list + full captures (CONUS)
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
{
for args in "10000 6000 0" "21696 2712 0" "21696 10848 0" "21696 21696 0" "21696 21696 1" "21696 21696 2" \
            "21696 21696 0 11 6 24" "21696 21696 0 11 6 256" "21696 21696 0 11 0" "21696 21696 1 11 0" "21696 21696 0 4 4" "21696 21696 1 4 4"; do
  timeout 120 ./build/strip_bw $args
done
} > gpurun_out/strip_bw.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --workload conus --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_conus_v3.json 2> gpurun_out/bench_conus_v3.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_fulldisk_v3.json 2> gpurun_out/bench_fulldisk_v3.err
python bench.py --workload meso --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_meso_v3.json 2> gpurun_out/bench_meso_v3.err
export OCTANE_NO_GRAPHS=1
B="python bench.py --workload conus --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_conus.csv $B > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pcg_pass1_tma -s 830 -c 2 -f -o gpurun_out/prof_pass1_tma $B > gpurun_out/ncu_p1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pcg_pass2 -s 830 -c 2 -f -o gpurun_out/prof_pass2 $B > gpurun_out/ncu_p2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:^k_build$' -s 28 -c 1 -f -o gpurun_out/prof_build $B > gpurun_out/ncu_build.log 2>&1
ls -la gpurun_out
