#!/bin/bash
# round 2, GPU call 9: sanity of the new tiling search + cooperative kernel restricted to small levels; N = 1 digests
O=gpurun_out/r02c9
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -m gpu -q -x -rs ) > $O/pytest.log 2>&1
tail -n 3 $O/pytest.log
for wl in conus fulldisk; do
  timeout 500 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --write-digest > $O/bench_${wl}.json 2> $O/bench_${wl}.err
done
cp tests/golden/digest_*.npz $O/
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c9/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 2), "Mpix/s", round(d["value"], 1), d["roofline"].get("fused"), d.get("check"))
    except Exception as e:
        print(f, "ERR", e)
PY
