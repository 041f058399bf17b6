#!/bin/bash
# round 2, GPU call 11 (2 GPUs): pipelined dispatcher on one GPU and on bands; N = 2 bench with e2e
O=gpurun_out/r02c11
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -rs -k "pipelined or pairs_in_flight or graph_and_plain" ) > $O/pytest_stream.log 2>&1
tail -n 5 $O/pytest_stream.log
( time timeout 900 python -m pytest tests/test_gpu_band.py -m gpu -q -rs -k "2-shape" ) > $O/pytest_band.log 2>&1
tail -n 5 $O/pytest_band.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload conus --steps 4 --warmup 3 \
    > $O/bench_conus_n2.json 2> $O/bench_conus_n2.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --workload conus --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_conus_n1.json 2> $O/bench_conus_n1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c11/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 2), "Mpix/s", round(d["value"], 1), "e2e", d.get("e2e"), d.get("check"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 5 $O/*.err
