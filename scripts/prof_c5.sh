#!/usr/bin/env bash
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q -k "big_rows or hier or config5 or nuts" > $out/pytest_c5.log 2>&1
echo "pytest rc=$? : $(grep -E 'passed|failed' $out/pytest_c5.log | tail -1)"; grep -E "^FAILED|^E  " $out/pytest_c5.log | head
python scripts/bench_c5.py 32768 3 2>&1 | tail -1
python scripts/bench_nuts.py 65536 128 40 2>&1 | tail -1 | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_big_hmc -c 1 -o $out/prof_c5 -f python scripts/bench_c5.py 2048 1 > /dev/null 2>&1
ls -la $out/prof_c5.ncu-rep
