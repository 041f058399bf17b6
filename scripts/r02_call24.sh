#!/bin/bash
# round 2, GPU call 24 (1 GPU, what is left of the budget): k_build at 5 and 6 resident blocks per SM (48 / 40 registers,
# 100-190 B of spills per thread) against the shipped 4, CONUS, stage_ms.build of one profiled step
O=gpurun_out/r02c24
mkdir -p $O
for occ in 5 6; do
  OCTANE_B200_LIB=$PWD/octane_b200/lib/variants/liboctane_b200_bocc$occ.so timeout 60 python bench.py --workload conus --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_conus_bocc$occ.json 2> $O/bench_conus_bocc$occ.err
done
timeout 60 python bench.py --workload conus --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_conus_bocc4.json 2> $O/bench_conus_bocc4.err
python - <<'PY' | tee gpurun_out/r02c24/build_occ_ab2.txt
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c24/bench_conus_bocc*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bocc")[1][0], "blocks/SM: build", round(d["stage_ms"]["build"], 2), "ms, pyramid", round(d["stage_ms"]["pyramid"], 2), "ms, pair", round(d["ms_per_step"], 1), "ms, check", d["check"].get("max_abs_du"))
    except Exception as ex:
        print(f, "ERR", ex)
PY
