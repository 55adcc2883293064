#!/usr/bin/env bash
# ncu launch list + full capture of the hand-written product kernel (run through gpurun)
out=gpurun_out
mkdir -p $out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_gemm.csv \
  python scripts/gemm_check.py prof > $out/prof_run.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_gemm_f16x3 -s 8 -c 3 -o $out/prof_gemm_r2 -f \
  python scripts/gemm_check.py prof >> $out/prof_run.log 2>&1
ls -la $out
