#!/bin/bash
# round 2, GPU call 13: merged-reduction kernel v6 (q = A p computed, never stored: 60 B/px): tests, benches, one capture
O=gpurun_out/r02c13
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_headline.py tests/test_gpu_parity.py -m gpu -q -x -rs ) > $O/pytest.log 2>&1
tail -n 3 $O/pytest.log
for wl in conus fulldisk; do
  timeout 500 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_${wl}.json 2> $O/bench_${wl}.err
done
timeout 300 python bench.py --workload meso --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_meso.json 2> $O/bench_meso.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c13/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, "ms/step", round(d["ms_per_step"], 1), "Mpix/s", round(d["value"], 1), r.get("fused"), d.get("check"), {k: round(v, 1) for k, v in d["stage_ms"].items() if isinstance(v, float)})
    except Exception as e:
        print(f, "ERR", e)
PY
OCTANE_NO_GRAPHS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_pcg_fused -s 945 -c 2 \
    -o $O/ncu_fused_conus -f python bench.py --workload conus --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_fused.log 2>&1
tail -n 2 $O/ncu_fused.log | cut -c1-200
