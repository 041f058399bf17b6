#!/bin/bash
# GPU call 9 (1 GPU): tapered full disk is slower in the bench but not under ncu -- data or allocator state?
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$B --noise-floor 2.0 > gpurun_out/u_taper_noise.json 2> gpurun_out/u_taper_noise.err
$B --empty-cache > gpurun_out/u_taper_emptycache.json 2> gpurun_out/u_taper_emptycache.err
$B --taper 0 --empty-cache > gpurun_out/u_notaper_emptycache.json 2> gpurun_out/u_notaper_emptycache.err
$B > gpurun_out/u_taper.json 2> gpurun_out/u_taper.err
python - <<'PY' > gpurun_out/u_summary.txt
import json,glob
for f in sorted(glob.glob('gpurun_out/u_*.json')):
    try:
        d=json.load(open(f))
        print(f, d['ms_per_step'], d['roofline']['pass1'], d['roofline']['pass2'], d['stage_ms']['build'], d['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/u_summary.txt
