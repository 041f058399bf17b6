#!/bin/bash
# round 2, GPU call 12: the whole GPU suite, the default bench as the driver runs it, the reference arm, the ncu launch
# list of one full-disk step and a full capture of two finest-level launches of the dominant kernel
O=gpurun_out/r02c12
mkdir -p $O
rm -f gpurun_out/parity_report.jsonl
( time timeout 1800 python -m pytest tests -m gpu -q -rs ) > $O/pytest_gpu.log 2>&1
tail -n 6 $O/pytest_gpu.log
cp gpurun_out/parity_report.jsonl $O/parity_report.jsonl
( time timeout 300 python __graft_entry__.py smoke ) > $O/smoke.log 2>&1
tail -n 2 $O/smoke.log
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > $O/bench_default.json 2> $O/bench_default.err
tail -c 3000 $O/bench_default.json
( time timeout 300 python bench.py --impl reference --steps 5 --warmup 3 ) > $O/bench_reference.json 2> $O/bench_reference.err
OCTANE_NO_GRAPHS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $O/launches_fulldisk.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_launches.log 2>&1
tail -n 2 $O/ncu_launches.log | cut -c1-300
OCTANE_NO_GRAPHS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_pcg_fused -s 945 -c 2 \
    -o $O/ncu_fused_fulldisk -f python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > $O/ncu_fused.log 2>&1
tail -n 2 $O/ncu_fused.log | cut -c1-300
ls -la $O
