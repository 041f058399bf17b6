#!/bin/bash
# round 2, GPU call 8: cooperative whole-solve kernel on the small levels (k_pcg_coop): full suite, small-scene benches, batch64
O=gpurun_out/r02c8
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q -rs -x ) > $O/pytest_all.log 2>&1
tail -n 4 $O/pytest_all.log
( time timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "flow_matches_oracle and 96x80" ) > $O/racecheck_coop.log 2>&1
tail -n 4 $O/racecheck_coop.log
timeout 300 python bench.py --workload meso --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_meso.json 2> $O/bench_meso.err
timeout 300 python bench.py --workload meso500 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_meso500.json 2> $O/bench_meso500.err
for s in 4 8 16; do
  timeout 300 python bench.py --workload batch64 --steps 3 --warmup 3 --streams $s > $O/bench_batch64_s$s.json 2> $O/bench_batch64_s$s.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02c8/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 2), "Mpix/s", round(d["value"], 1), "launches", d.get("gpu_launches"), "e2e", d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 3 $O/*.err
