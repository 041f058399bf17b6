#!/bin/bash
# Experiment prepared at the end of round 1, NOT RUN YET (the round's GPU budget was spent): pass 1 with one pixel
# per consumer thread = 31 consumer warps per SM instead of 16 (OCTANE_P1_PX=1; pcg_tma.cu, Consumers<1>).
# Going from 8 to 16 warps was worth +13 % on the full disk (profiles/r01_pass1_ab_stages_px.txt); the kernel is bound
# by the latency of each warp's dependent chain per row, so more warps in flight is the cheapest lever left.
# The default kernels' SASS is unchanged by the extra instantiations.
#   1. parity of the variant (large scenes use the TMA kernel) against the oracle / properties
#   2. A/B of pass 1 on CONUS and the full disk: roofline.pass1 GB/s and ms_per_step of each line
# Second prepared experiment, also NOT RUN YET: OCTANE_CONST_WN=1 -- the solves of the first GNC stage (a third of
# them) do not read the W and N coupling planes, which the build fills with -1 everywhere in that stage
# (8 of pass 1's 68 B/px; k_pcg_pass1_tma<..., CWN = true>).  Same arithmetic, so results must be bit-identical to
# the default build's: the pytest lines below run the fixture / oracle parity tests under each switch.
# Run on one GPU:  gpurun --timeout 1500 -- 'bash scripts/exp_pass1_px.sh'
mkdir -p gpurun_out/px
run() {   # tag, then VAR=value pairs
    tag=$1; shift
    env "$@" timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
        -k "meso_2000 or large_scene or graph_and_plain or flow_matches" > gpurun_out/px/pytest_$tag.log 2>&1
    for wl in conus fulldisk; do
        env "$@" timeout 300 python bench.py --workload $wl --steps 2 --warmup 2 --no-e2e --no-cpu-baseline \
            > gpurun_out/px/bench_${wl}_$tag.json 2> gpurun_out/px/bench_${wl}_$tag.err
    done
}
run px2 OCTANE_P1_PX=2
run px1 OCTANE_P1_PX=1
run px2_cwn OCTANE_P1_PX=2 OCTANE_CONST_WN=1
run px1_cwn OCTANE_P1_PX=1 OCTANE_CONST_WN=1
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/px/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", round(d["ms_per_step"], 1), "pass1", d["roofline"]["pass1"], "pass2", d["roofline"]["pass2"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 2 gpurun_out/px/pytest_*.log
