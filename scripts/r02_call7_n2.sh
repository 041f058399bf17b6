#!/bin/bash
# round 2, GPU call 7 (2 GPUs): banded == single GPU with the merged-reduction kernels (peer-memory halo rows of r, q and
# six sums per iteration inside the kernel), N = 2 bench lines
O=gpurun_out/r02c7
mkdir -p $O
( time timeout 1200 python -m pytest tests/test_gpu_band.py -m gpu -q -rs ) > $O/pytest_band.log 2>&1
tail -n 6 $O/pytest_band.log
for wl in conus fulldisk; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $wl --steps 3 --warmup 3 \
      > $O/bench_${wl}_n2.json 2> $O/bench_${wl}_n2.err
  tail -c 600 $O/bench_${wl}_n2.json; echo
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c7/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 1), "Mpix/s", round(d["value"], 1), d["roofline"].get("fused"), "e2e", d.get("e2e", {}).get("ms_per_step"), d.get("check"))
    except Exception as e:
        print(f, "ERR", e)
PY
