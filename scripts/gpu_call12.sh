#!/bin/bash
# GPU call 12 (2 GPUs): row-band parity (banded == single GPU) and the N=2 bench over NCCL
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
nvidia-smi topo -m >> gpurun_out/gpus2.txt 2>&1
python -m pytest tests/test_gpu_band.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_band2.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_fulldisk_n2.json 2> gpurun_out/bench_fulldisk_n2.err
$T bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --workload conus > gpurun_out/bench_conus_n2.json 2> gpurun_out/bench_conus_n2.err
$T bench.py --gpus 2 --steps 5 --warmup 2 --no-cpu-baseline --workload meso > gpurun_out/bench_meso_n2.json 2> gpurun_out/bench_meso_n2.err
tail -3 gpurun_out/*_n2.err
cat gpurun_out/bench_fulldisk_n2.json gpurun_out/bench_conus_n2.json gpurun_out/bench_meso_n2.json
