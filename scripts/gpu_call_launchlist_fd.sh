#!/bin/bash
# ncu launch list of one whole step of the DEFAULT bench workload (full disk 21696^2, one GPU): our kernels only
# (-k regex:^k_), the 2300 launches of the first step.  Times under ncu are cold-cache and serialised: shares only.
mkdir -p gpurun_out
timeout 88 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_' -c 2300 --csv \
    --log-file gpurun_out/launches_fulldisk.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fd.log 2>&1
echo "exit $?"; tail -c 600 gpurun_out/ncu_fd.log; wc -l gpurun_out/launches_fulldisk.csv
