#!/bin/bash
# decoupled NUTS sampler: parity with the step-wise path, then timing against the step-synchronous loop
cd "$(dirname "$0")/.."

timeout 300 python -m pytest tests/test_gpu_meads.py -q 2>&1 | grep -v "^  " | tail -25
timeout 200 python scripts/nuts_decoupled.py 65536 128 64
BJX_NUTS_DECOUPLED=0 timeout 200 python scripts/nuts_decoupled.py 65536 128 64
timeout 200 python scripts/nuts_decoupled.py 65536 128 16
