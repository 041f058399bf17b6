#!/bin/bash
# round 2, GPU call 20 (2 GPUs): banded pipelined dispatcher with the copy-out held back to the next pair's finest level
# and without per-store fences: band tests at world 2 (incl. pipelined == device path, re-plan failure), full-disk bench
O=gpurun_out/r02c20
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_band.py -m gpu -q -rs -k "2-shape" ) > $O/pytest_band.log 2>&1
tail -n 4 $O/pytest_band.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 2 --steps 5 --warmup 3 --rank-stats \
      > $O/bench_fulldisk_n2.json 2> $O/bench_fulldisk_n2.err
OCTANE_STREAM_DEFER=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 3 --warmup 3 \
      > $O/bench_fulldisk_n2_nodefer.json 2> $O/bench_fulldisk_n2_nodefer.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c20/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e", {})
        print(f.split("/")[-1], "dev ms", round(d["ms_per_step"], 1), "e2e ms", round(e.get("ms_per_step", 0), 1), "lat", round(e.get("latency_ms_per_pair", 0), 1),
              "host wait", round(e.get("host_ms_in_wait", 0), 1), "check", d["check"].get("max_abs_du"), d["check"].get("within_gates"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
tail -n 3 $O/*.err | grep -v "^\*\|OMP_NUM"
