#!/bin/bash
# round 2, GPU call 14 (2 GPUs): merged-reduction kernel v6 + branch-free reciprocal: fused / headline tests, bands at
# world 2 (incl. pipelined dispatcher and EHALO recovery), fresh single-GPU digests, N = 2 full-disk line
O=gpurun_out/r02c14
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_headline.py -m gpu -q -x -rs ) > $O/pytest_fused.log 2>&1
tail -n 3 $O/pytest_fused.log
( time timeout 900 python -m pytest tests/test_gpu_band.py -m gpu -q -rs -k "2-shape" ) > $O/pytest_band.log 2>&1
tail -n 12 $O/pytest_band.log | cut -c1-300
for wl in conus fulldisk; do
  CUDA_VISIBLE_DEVICES=0 timeout 500 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --write-digest > $O/bench_${wl}_n1.json 2> $O/bench_${wl}_n1.err
done
cp tests/golden/digest_*.npz $O/
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 4 --warmup 3 \
    > $O/bench_fulldisk_n2.json 2> $O/bench_fulldisk_n2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c14/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 2), "Mpix/s", round(d["value"], 1), d["roofline"].get("fused"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), d.get("check"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 3 $O/*.err | cut -c1-300
