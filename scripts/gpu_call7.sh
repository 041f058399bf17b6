#!/bin/bash
# GPU call 7 (1 GPU): pass 1 is slower on the tapered full disk than on an untapered scene of the same size -- why?
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
OCTANE_DUMP_EVENTS=gpurun_out/events_fd_taper.txt $B > gpurun_out/t_fd_taper.json 2> gpurun_out/t_fd_taper.err
OCTANE_DUMP_EVENTS=gpurun_out/events_fd_notaper.txt $B --taper 0 > gpurun_out/t_fd_notaper.json 2> gpurun_out/t_fd_notaper.err
OCTANE_B200_LIB=$PWD/build/liboctane_b200_ftz.so OCTANE_DUMP_EVENTS=gpurun_out/events_fd_taper_ftz.txt $B > gpurun_out/t_fd_taper_ftz.json 2> gpurun_out/t_fd_taper_ftz.err
$B --workload conus --taper 1 > gpurun_out/t_conus_taper.json 2> gpurun_out/t_conus_taper.err
python - <<'PY' > gpurun_out/t_summary.txt
import json,glob
for f in sorted(glob.glob('gpurun_out/t_*.json')):
    try:
        d=json.load(open(f))
        print(f, d['ms_per_step'], d['roofline']['pass1'], d['roofline']['pass2'], d['stage_ms']['build'], d['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/t_summary.txt
