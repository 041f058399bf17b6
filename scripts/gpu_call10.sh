#!/bin/bash
# GPU call 10 (1 GPU): pass 1 with 16 consumer warps (2 px/thread) vs 8 (4 px/thread), +/- FTZ; ingest goldens; parity
set -x
mkdir -p gpurun_out
python tests/golden/make_golden.py gpurun_out/golden --ingest-only > gpurun_out/make_golden_ingest.log 2>&1
cp gpurun_out/golden/ingest_*.npz gpurun_out/golden/uv2pix_*.npz tests/golden/ 2>/dev/null
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$B > gpurun_out/v_fd_px2.json 2> gpurun_out/v_fd_px2.err
OCTANE_P1_PX=4 $B > gpurun_out/v_fd_px4.json 2> gpurun_out/v_fd_px4.err
OCTANE_B200_LIB=$PWD/build/liboctane_b200_ftz.so $B > gpurun_out/v_fd_px2_ftz.json 2> gpurun_out/v_fd_px2_ftz.err
$B --taper 0 > gpurun_out/v_fdnotaper_px2.json 2> gpurun_out/v_fdnotaper_px2.err
$B --workload conus > gpurun_out/v_conus_px2.json 2> gpurun_out/v_conus_px2.err
OCTANE_P1_PX=4 $B --workload conus > gpurun_out/v_conus_px4.json 2> gpurun_out/v_conus_px4.err
$B --workload meso > gpurun_out/v_meso_px2.json 2> gpurun_out/v_meso_px2.err
python - <<'PY' > gpurun_out/v_summary.txt
import json,glob
for f in sorted(glob.glob('gpurun_out/v_*.json')):
    try:
        d=json.load(open(f))
        print(f, d['ms_per_step'], d['roofline']['pass1'], d['roofline']['pass2'], d['stage_ms']['build'], d['clocks'])
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/v_summary.txt
