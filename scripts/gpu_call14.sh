#!/bin/bash
# GPU call 14 (8 GPUs): band parity at world 2/4/8, full-disk bench at N = 8 and 4 (peer-memory path), N = 8 over NCCL, CONUS N = 4
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
timeout 900 python -m pytest tests/test_gpu_band.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_band8.log
run() { n=$1; name=$2; shift 2; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "exit $name $?" >> gpurun_out/exitcodes8.txt; }
: > gpurun_out/exitcodes8.txt
run 8 fulldisk_n8 --steps 3 --warmup 3 --no-cpu-baseline
run 4 fulldisk_n4 --steps 3 --warmup 3 --no-cpu-baseline
OCTANE_COMM=nccl run 8 fulldisk_n8_nccl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
run 4 conus_n4 --steps 3 --warmup 3 --no-cpu-baseline --workload conus
run 8 conus_n8 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --workload conus
cat gpurun_out/exitcodes8.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*_n[48]*.json')):
    try:
        line=[l for l in open(f).read().splitlines() if l.startswith('{')][0]
        d=json.loads(line); print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['gpu_launches'], d['roofline']['pass1'], d['roofline']['pass2'], (d.get('e2e') or {}).get('ms_per_step'))
    except Exception as e: print(f,'ERR',e)
PY
tail -5 gpurun_out/pytest_band8.log
