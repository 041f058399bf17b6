#!/bin/bash
# GPU call 13 (2 GPUs): per-iteration exchanges through peer memory (CUDA IPC) vs NCCL
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_band.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_band2_p2p.log
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_fulldisk_n2_p2p.json 2> gpurun_out/bench_fulldisk_n2_p2p.err
$T bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --workload conus > gpurun_out/bench_conus_n2_p2p.json 2> gpurun_out/bench_conus_n2_p2p.err
$T bench.py --gpus 2 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --workload meso > gpurun_out/bench_meso_n2_p2p.json 2> gpurun_out/bench_meso_n2_p2p.err
OCTANE_COMM=nccl $T bench.py --gpus 2 --steps 5 --warmup 2 --no-cpu-baseline --no-e2e --workload meso > gpurun_out/bench_meso_n2_nccl.json 2> gpurun_out/bench_meso_n2_nccl.err
for f in gpurun_out/bench_*_n2_p2p.err; do echo == $f; tail -n 5 $f; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*_n2_*.json')):
    try:
        d=json.load(open(f)); print(f, d['ms_per_step'], d['gpu_launches'], d['roofline']['pass1'], d['roofline']['pass2'])
    except Exception as e: print(f,'ERR',e)
PY
cat gpurun_out/pytest_band2_p2p.log | tail -15
