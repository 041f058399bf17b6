#!/bin/bash
# round 2, GPU call 21 (1 GPU): the final library -- whole GPU suite, smoke, the default bench exactly as the driver runs it
O=gpurun_out/r02c21
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -q -rs -x ) > $O/pytest_gpu.log 2>&1
tail -n 6 $O/pytest_gpu.log
cp tests/parity_report.jsonl $O/ 2>/dev/null || cp gpurun_out/parity_report.jsonl $O/ 2>/dev/null
( time python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1
tail -n 3 $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c21/bench_default.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("ms/step", round(d["ms_per_step"], 1), "Mpix/s", round(d["value"], 1), "frac", round(d["roofline"]["frac"], 3), "e2e", round(e["ms_per_step"], 1), round(e["value"], 1),
      "lat", round(e["latency_ms_per_pair"], 1), "stages", {k: round(v, 1) for k, v in d["stage_ms"].items() if isinstance(v, float)}, d["check"]["max_abs_du"], d["clocks"])
PY
tail -n 3 $O/bench_default.err
