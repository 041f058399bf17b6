#!/bin/bash
# GPU call 6 (1 GPU): what makes pass 1 slower per byte on the full disk?  Shape / knob matrix.
set -x
mkdir -p gpurun_out
{ ./build/strip_bw 21696 21696 3; ./build/strip_bw 21696 21696 0; } > gpurun_out/strip_bw2.txt 2>&1
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
probe() { name=$1; size=$2; shift 2; env "$@" $B --size $size > gpurun_out/p1_$name.json 2> gpurun_out/p1_$name.err; }
probe fd 21696x21696 X=1
probe fd_half 21696x10848 X=1
probe fd_8th 21696x2712 X=1
probe w20480 20480x21696 X=1
probe conus_tall 10000x40000 X=1
probe fd_rs82 21696x21696 OCTANE_P1_RS=82
probe fd_rs256 21696x21696 OCTANE_P1_RS=256
probe fd_sw512 21696x21696 OCTANE_P1_SW=512
probe fd_sw1024 21696x21696 OCTANE_P1_SW=1024
probe fd_old 21696x21696 OCTANE_B200_LIB=$PWD/build/liboctane_b200_old.so
probe conus_old 10000x6000 OCTANE_B200_LIB=$PWD/build/liboctane_b200_old.so
probe conus 10000x6000 X=1
python - <<'PY' > gpurun_out/p1_summary.txt
import json,glob
for f in sorted(glob.glob('gpurun_out/p1_*.json')):
    try:
        d=json.load(open(f))
        print(f, d['ms_per_step'], d['roofline']['pass1'], d['roofline']['pass2'], d['stage_ms']['build'])
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/p1_summary.txt
