#!/bin/bash
# round 2, GPU call 10 (8 GPUs): banded == single GPU at world 4 and 8 with the merged-reduction kernels; N = 8 and N = 4
# full-disk bench lines with the in-run result check against the stored single-GPU digest
O=gpurun_out/r02c15
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
( time timeout 900 python -m pytest tests/test_gpu_band.py -m gpu -q -rs -k "4-shape or 8-shape" ) > $O/pytest_band.log 2>&1
tail -n 5 $O/pytest_band.log
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --rank-stats \
      > $O/bench_fulldisk_n$n.json 2> $O/bench_fulldisk_n$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --workload batch64 --steps 3 --warmup 3 --streams 8 \
      > $O/bench_batch64_n8.json 2> $O/bench_batch64_n8.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c15/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 2), "Mpix/s", round(d["value"], 1), d.get("roofline", {}).get("fused"), "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("check"), {k: round(v, 1) for k, v in d.get("stage_ms", {}).items() if isinstance(v, float)})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 3 $O/*.err
