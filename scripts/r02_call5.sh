#!/bin/bash
# round 2, GPU call 2: first run of the merged-reduction PCG kernel (pcg_fused.cu): stage tests against both
# oracle recurrences, the whole GPU suite, solver 0 / 1 A/B on CONUS and the full disk, one ncu capture.
O=gpurun_out/r02c5
mkdir -p $O
( time timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -rs ) > $O/pytest_fused.log 2>&1
tail -n 5 $O/pytest_fused.log


( time timeout 1500 python -m pytest tests -m gpu -q -rs --deselect tests/test_gpu_fused.py ) > $O/pytest_all.log 2>&1
tail -n 5 $O/pytest_all.log
( time timeout 300 python __graft_entry__.py smoke ) > $O/smoke.log 2>&1
tail -n 2 $O/smoke.log
for solver in 1; do
  for wl in conus fulldisk; do
    timeout 500 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --solver $solver \
        > $O/bench_${wl}_s$solver.json 2> $O/bench_${wl}_s$solver.err
    tail -c 1500 $O/bench_${wl}_s$solver.json | head -c 1500; echo
  done
done
OCTANE_NO_GRAPHS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_pcg_fused -s 900 -c 2 \
    -o $O/ncu_fused_conus -f python bench.py --workload conus --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_fused.log 2>&1
tail -n 3 $O/ncu_fused.log
ls -la $O
