#!/bin/bash
# round 2, GPU call 16: the final kernels -- whole GPU suite, the default bench as the driver runs it, the reference arm,
# k_build occupancy A/B, the ncu launch list of one full-disk step (our kernels only) and a full capture of two
# finest-level launches of the dominant kernel
O=gpurun_out/r02c16
mkdir -p $O
rm -f gpurun_out/parity_report.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q -rs ) > $O/pytest_gpu.log 2>&1
tail -n 6 $O/pytest_gpu.log
cp gpurun_out/parity_report.jsonl $O/parity_report.jsonl
( time timeout 300 python __graft_entry__.py smoke ) > $O/smoke.log 2>&1
tail -n 2 $O/smoke.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > $O/bench_default.json 2> $O/bench_default.err
tail -c 1500 $O/bench_default.json
( time timeout 300 python bench.py --impl reference --steps 5 --warmup 3 ) > $O/bench_reference.json 2> $O/bench_reference.err
for occ in 2 4; do
  OCTANE_B200_LIB=$PWD/octane_b200/lib/variants/liboctane_b200_bocc$occ.so timeout 300 python bench.py --workload conus --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_conus_bocc$occ.json 2> $O/bench_conus_bocc$occ.err
done
timeout 300 python bench.py --workload conus --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_conus_bocc3.json 2> $O/bench_conus_bocc3.err
python - <<'PY' > gpurun_out/r02c16/build_occ_ab.txt
import json, glob
print("# k_build occupancy target (blocks of 256 threads per SM), CONUS 10000 x 6000, stage_ms.build of one profiled step")
for f in sorted(glob.glob("gpurun_out/r02c16/bench_conus_bocc*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bocc")[1][0], "blocks/SM: build", round(d["stage_ms"]["build"], 2), "ms, pair", round(d["ms_per_step"], 1), "ms")
    except Exception as e:
        print(f, "ERR", e)
PY
cat $O/build_occ_ab.txt
OCTANE_NO_GRAPHS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 1400 --csv --log-file $O/launches_fulldisk.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1
tail -n 2 $O/ncu_launches.log | cut -c1-300
OCTANE_NO_GRAPHS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_pcg_fused -s 945 -c 2 \
    -o $O/ncu_fused_fulldisk -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_fused.log 2>&1
tail -n 2 $O/ncu_fused.log | cut -c1-300
ls -la $O
