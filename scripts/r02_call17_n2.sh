#!/bin/bash
# round 2, GPU call 17 (2 GPUs): (a) the band tests at world 2 with the new re-plan-failure case (one rank runs out
# of device memory: every rank must get an error, the context must solve again); (b) where the pipelined end-to-end
# loop loses its overlap at N > 1: the full-disk bench at N = 2 with the copy-in, the copy-out or both switched off
# (OCTANE_STREAM_SKIP, timing experiment only) and the stage times of the last pipelined pair beside the device-only ones
O=gpurun_out/r02c17
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
nvidia-smi topo -m > $O/topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_gpu_band.py -m gpu -q -rs -k "2-shape" ) > $O/pytest_band.log 2>&1
tail -n 5 $O/pytest_band.log
for skip in 0 3 1 2; do
  OCTANE_STREAM_SKIP=$skip timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2960$skip bench.py --gpus 2 --steps 3 --warmup 3 \
      > $O/bench_fulldisk_n2_skip$skip.json 2> $O/bench_fulldisk_n2_skip$skip.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c17/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e", {})
        print(f, "dev ms", round(d["ms_per_step"], 1), "e2e ms", round(e.get("ms_per_step", 0), 1), "lat", round(e.get("latency_ms_per_pair", 0), 1),
              "stages e2e", {k: round(v, 1) for k, v in e.get("stage_ms_last_pair", {}).items()},
              "stages dev", {k: round(v, 1) for k, v in d.get("stage_ms", {}).items() if isinstance(v, float)})
    except Exception as ex:
        print(f, "ERR", ex)
PY
tail -n 3 $O/*.err
