#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --workload conus --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_conus_v3.json 2> gpurun_out/bench_conus_v3.err
OCTANE_PLANE_PAD_MB=1700 python bench.py --workload conus --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_conus_v3_pad.json 2> gpurun_out/bench_conus_v3_pad.err
python bench.py --workload meso --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_meso_v3.json 2> gpurun_out/bench_meso_v3.err
python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_fulldisk_v3.json 2> gpurun_out/bench_fulldisk_v3.err
