#!/usr/bin/env bash
# ncu launch list of one config-2 transition + full captures of the fused product kernel (run through gpurun)
out=gpurun_out
mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $out/launches_dense.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/bench_under_ncu.json 2> $out/prof_err.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_f16x3 -s 60 -c 3 -o $out/prof_gemm_fused -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline >> $out/prof_err.log 2>&1
ls -la $out/*.ncu-rep $out/launches_dense.csv
