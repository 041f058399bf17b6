#!/bin/bash
# GPU call 8 (1 GPU): ncu of pass 1 on the tapered vs untapered full disk (iteration 0 and a steady iteration)
set -x
mkdir -p gpurun_out
export OCTANE_NO_GRAPHS=1
B="python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
SEC="--section SpeedOfLight --section WarpStateStats --section InstructionStats --section MemoryWorkloadAnalysis --section SchedulerStats --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy"
timeout 420 ncu $SEC --clock-control none -k regex:k_pcg_pass1_tma -s 810 -c 1 -f -o gpurun_out/prof_p1_fd_taper_it0 $B > gpurun_out/ncu_a.log 2>&1
timeout 420 ncu $SEC --clock-control none -k regex:k_pcg_pass1_tma -s 810 -c 1 -f -o gpurun_out/prof_p1_fd_notaper_it0 $B --taper 0 > gpurun_out/ncu_b.log 2>&1
timeout 420 ncu $SEC --clock-control none -k regex:k_pcg_pass1_tma -s 830 -c 1 -f -o gpurun_out/prof_p1_fd_taper_it20 $B > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out | grep prof_p1
