#!/bin/bash
# round 2, GPU call 22 (8 GPUs, bench only -- what is left of the round's GPU budget): the full disk on eight row bands
# with the deferred copy-out, device-timed value and pipelined end-to-end figure
O=gpurun_out/r02c22
mkdir -p $O
timeout 85 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 8 --steps 5 --warmup 3 --rank-stats \
      > $O/bench_fulldisk_n8.json 2> $O/bench_fulldisk_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02c22/bench_fulldisk_n8.json").read().strip().splitlines()[-1])
    e = d.get("e2e", {})
    print("dev ms", round(d["ms_per_step"], 1), "Mpix/s", round(d["value"], 1), "e2e ms", round(e.get("ms_per_step", 0), 1), round(e.get("value", 0), 1), "lat", round(e.get("latency_ms_per_pair", 0), 1), d["check"])
except Exception as ex:
    print("ERR", ex)
PY
tail -n 3 $O/*.err | grep -v "^\*\|OMP_NUM"
