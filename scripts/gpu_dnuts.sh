#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_round2.py -q -k "dense_path_nuts" -s 2>&1 | grep -v "^  " | tail -12
timeout 300 python scripts/bench_dense_nuts.py 16384 1024 8 0.5 6
timeout 300 python scripts/bench_dense_nuts.py 65536 1024 8 0.5 4
timeout 300 python scripts/bench_dense_nuts.py 65536 256 8 0.5 6
