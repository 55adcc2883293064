#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/dnuts_launches.csv python scripts/bench_dense_nuts.py 16384 1024 8 0.5 1 > gpurun_out/dnuts_prof.log 2>&1
python scripts/ncu_summary.py gpurun_out/dnuts_launches.csv gpurun_out/dnuts_summary.md - "dense NUTS 16384x1024" | head -40; tail -3 gpurun_out/dnuts_prof.log
