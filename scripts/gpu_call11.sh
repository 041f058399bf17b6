#!/bin/bash
# GPU call 11 (1 GPU): parity incl. CLI end to end; pass-1 A/B: ring depth 4/5 x pixels per thread 2/4 (all with FTZ)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
NS4=$PWD/build/liboctane_b200_ns4.so
for wl in fulldisk conus; do
  OCTANE_P1_PX=2 $B --workload $wl > gpurun_out/w_${wl}_ns5_px2.json 2> gpurun_out/w_${wl}_ns5_px2.err
  OCTANE_P1_PX=4 $B --workload $wl > gpurun_out/w_${wl}_ns5_px4.json 2> gpurun_out/w_${wl}_ns5_px4.err
  OCTANE_B200_LIB=$NS4 OCTANE_P1_PX=2 $B --workload $wl > gpurun_out/w_${wl}_ns4_px2.json 2> gpurun_out/w_${wl}_ns4_px2.err
  OCTANE_B200_LIB=$NS4 OCTANE_P1_PX=4 $B --workload $wl > gpurun_out/w_${wl}_ns4_px4.json 2> gpurun_out/w_${wl}_ns4_px4.err
done
python - <<'PY' > gpurun_out/w_summary.txt
import json,glob
for f in sorted(glob.glob('gpurun_out/w_*.json')):
    try:
        d=json.load(open(f))
        print(f, d['ms_per_step'], d['roofline']['pass1'], d['roofline']['pass2'], d['stage_ms']['build'], d['clocks']['sm_mhz'])
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/w_summary.txt
