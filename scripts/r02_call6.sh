#!/bin/bash
# round 2, GPU call 6: fused kernel with per-mode ring depth (15 and 14 consumer warps), batch64 with more contexts in flight
O=gpurun_out/r02c6
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_headline.py -m gpu -q -x -rs ) > $O/pytest_fused.log 2>&1
tail -n 3 $O/pytest_fused.log
for wl in conus fulldisk; do
  timeout 500 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_${wl}_w15.json 2> $O/bench_${wl}_w15.err
  OCTANE_B200_LIB=$PWD/octane_b200/lib/variants/liboctane_b200_w14.so timeout 500 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_${wl}_w14.json 2> $O/bench_${wl}_w14.err
done
timeout 300 python bench.py --workload meso --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_meso.json 2> $O/bench_meso.err
for s in 8 16 32; do
  timeout 300 python bench.py --workload batch64 --steps 3 --warmup 3 --streams $s > $O/bench_batch64_s$s.json 2> $O/bench_batch64_s$s.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c6/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, "ms/step", round(d["ms_per_step"], 1), "Mpix/s", round(d["value"], 1), r.get("fused"), "e2e", d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
OCTANE_NO_GRAPHS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_pcg_fused -s 900 -c 2 \
    -o $O/ncu_fused_conus -f python bench.py --workload conus --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_fused.log 2>&1
tail -n 2 $O/ncu_fused.log
