#!/bin/bash
# round 2, GPU call 18 (2 GPUs): why the pipelined end-to-end loop hides no copy at N > 1 (call 17: without the copy-out
# the loop runs at the device time, with it at the one-pair-at-a-time latency).  Suspect: with NCCL's streams in the
# process the 8 default hardware queues are shared and the copy stream falls in line with the solve stream.
#   a) default, with the host time spent inside submit / wait       b) CUDA_DEVICE_MAX_CONNECTIONS=32
#   c) copy streams at the highest priority (OCTANE_COPY_PRIO=1)     d) both
O=gpurun_out/r02c18
mkdir -p $O
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 \
      > $O/bench_fulldisk_n2_$name.json 2> $O/bench_fulldisk_n2_$name.err
}
run a_default X=1
run b_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run c_prio OCTANE_COPY_PRIO=1
run d_both CUDA_DEVICE_MAX_CONNECTIONS=32 OCTANE_COPY_PRIO=1
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pipelined or fresh" ) > $O/pytest_stream.log 2>&1
tail -n 4 $O/pytest_stream.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c18/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e", {})
        print(f.split("n2_")[1], "dev ms", round(d["ms_per_step"], 1), "e2e ms", round(e.get("ms_per_step", 0), 1), "lat", round(e.get("latency_ms_per_pair", 0), 1),
              "host submit", round(e.get("host_ms_in_submit", 0), 1), "host wait", round(e.get("host_ms_in_wait", 0), 1), "build", round(d["stage_ms"]["build"], 1))
    except Exception as ex:
        print(f, "ERR", ex)
PY
tail -n 3 $O/*.err | grep -v "^\*\|OMP_NUM" 
