#!/bin/bash
# round 2, GPU call 23 (1 GPU, the last 4 GPU-minutes): k_gradient (4 columns per thread, rolling rows, division-free /12)
# and k_zoom_in (column interpolation shared through shared memory) -- the full-disk bench first (its in-run check
# compares the flow with the digest stored by the old kernels: 0.0 = bit-identical pipeline), then the stage / fixture tests
O=gpurun_out/r02c23
mkdir -p $O
timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_fulldisk.json 2> $O/bench_fulldisk.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02c23/bench_fulldisk.json").read().strip().splitlines()[-1])
    print("ms/step", round(d["ms_per_step"], 1), "stages", {k: round(v, 1) for k, v in d["stage_ms"].items() if isinstance(v, float)}, "check", d["check"].get("max_abs_du"), d["check"].get("max_abs_dv"), "e2e", round(d["e2e"]["ms_per_step"], 1))
except Exception as ex:
    print("ERR", ex)
PY
tail -n 3 $O/bench_fulldisk.err
( time timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stage_ or flow_matches_reference_fixture or flow_matches_oracle or dispatcher_with" ) > $O/pytest_subset.log 2>&1
tail -n 5 $O/pytest_subset.log
