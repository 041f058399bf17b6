#!/bin/bash
# round 2, GPU call 1: headline fixtures from the reference's sm_100 build, the new headline parity tests against
# them, the enlarged smoke, and the A/B of the two pass-1 levers prepared in round 1 (OCTANE_P1_PX, OCTANE_CONST_WN).
O=gpurun_out/r02c1
mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt
( time timeout 900 python tests/golden/make_golden_headline.py gpurun_out/golden_headline ) > $O/make_headline.log 2>&1
export OCTANE_GOLDEN_EXTRA=$PWD/gpurun_out/golden_headline
( time timeout 900 python -m pytest tests/test_gpu_headline.py -m gpu -q -rs ) > $O/pytest_headline.log 2>&1
( time timeout 300 python __graft_entry__.py smoke ) > $O/smoke.log 2>&1
ab() {   # tag, then VAR=value pairs
    tag=$1; shift
    for wl in fulldisk conus; do
        env "$@" timeout 400 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu-baseline \
            > $O/bench_${wl}_$tag.json 2> $O/bench_${wl}_$tag.err
    done
    env "$@" timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "meso_2000 or graph_and_plain or flow_matches" \
        > $O/pytest_$tag.log 2>&1
}
ab px2 OCTANE_P1_PX=2
ab px1 OCTANE_P1_PX=1
ab px2_cwn OCTANE_P1_PX=2 OCTANE_CONST_WN=1
ab px1_cwn OCTANE_P1_PX=1 OCTANE_CONST_WN=1
ab px4 OCTANE_P1_PX=4
python - <<'PY' > gpurun_out/r02c1/summary.txt 2>&1
import glob, json
for f in sorted(glob.glob("gpurun_out/r02c1/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 1), "pass1", d["roofline"]["pass1"], "pass2", d["roofline"]["pass2"], "stage", {k: round(v, 1) for k, v in d["stage_ms"].items() if isinstance(v, float)})
    except Exception as e:
        print(f, "ERR", e)
PY
cat $O/summary.txt
tail -n 3 $O/pytest_*.log $O/make_headline.log $O/smoke.log
