#!/usr/bin/env bash
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1
echo "pytest rc=$? : $(grep -E 'passed|failed' $out/pytest_gpu.log | tail -1)"; grep -E "^FAILED|^E  |worst" $out/pytest_gpu.log | head -20
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q -s -k "chees" 2>&1 | grep -E "ChEES|passed|failed" | head
python scripts/bench_c5.py 32768 3 2>&1 | tail -1
