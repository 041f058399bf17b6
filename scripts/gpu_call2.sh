#!/bin/bash
# why is pass 1 slower per byte on the full disk than on CONUS?
set -x
mkdir -p gpurun_out
export OCTANE_NO_GRAPHS=1
B="python bench.py --workload fulldisk --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_pcg_pass1 -s 830 -c 2 -f -o gpurun_out/prof_pass1_fulldisk $B > gpurun_out/ncu_p1_fd.log 2>&1
unset OCTANE_NO_GRAPHS
for sz in 20000x3000 10000x6000 5000x12000 10016x6000 21696x2712 21696x5424; do
  python bench.py --size $sz --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/size_$sz.json 2> gpurun_out/size_$sz.err
done
OCTANE_NO_TMA=1 python bench.py --size 21696x5424 --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/size_21696x5424_notma.json 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
