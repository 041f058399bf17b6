#!/bin/bash
# round 2, GPU call 19 (2 GPUs): does a system-scope fence cost more while a device-to-host copy is in flight?
# Call 18 ruled out hardware-queue sharing and a blocking submit (1.1 ms of host time per pair).  Variants of the library:
#   default       per-store + per-block + release fences at system scope (what ships)
#   nostorefence  without the per-store fences (the per-block fence after the block's barrier covers them)
#   gpuscope      every peer fence / release / acquire at GPU scope -- NOT valid between two GPUs, timing experiment only
O=gpurun_out/r02c19
mkdir -p $O
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 3 --warmup 3 \
      > $O/bench_fulldisk_n2_$name.json 2> $O/bench_fulldisk_n2_$name.err
}
run default X=1
run nostorefence OCTANE_B200_LIB=$PWD/octane_b200/lib/variants/liboctane_b200_nostorefence.so
run gpuscope OCTANE_B200_LIB=$PWD/octane_b200/lib/variants/liboctane_b200_gpuscope.so
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c19/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e", {})
        print(f.split("n2_")[1], "dev ms", round(d["ms_per_step"], 1), "e2e ms", round(e.get("ms_per_step", 0), 1), "lat", round(e.get("latency_ms_per_pair", 0), 1),
              "fused ms", round(d["roofline"]["fused"]["avg_ms"], 4), "check", d["check"].get("max_abs_du"), d["check"].get("within_gates"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
tail -n 3 $O/*.err | grep -v "^\*\|OMP_NUM"
