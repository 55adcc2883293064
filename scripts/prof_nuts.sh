#!/usr/bin/env bash
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1
echo "pytest rc=$? : $(grep -E 'passed|failed' $out/pytest_gpu.log | tail -1)"
grep -E "^FAILED|^ERROR|worst elementwise" $out/pytest_gpu.log | head -30
python scripts/bench_nuts.py 65536 128 40 > $out/nuts40.jsonl 2>&1; tail -3 $out/nuts40.jsonl | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_nuts.csv python scripts/bench_nuts.py 65536 128 3 > /dev/null 2>&1
