#!/usr/bin/env bash
for v in 0 3; do for d in 16 80 20 18 22; do
  BJX_GEMM_VARIANT=$v BJX_GEMM_DEBUG=$d timeout 100 python scripts/gemm_check.py loop 2>&1 | tail -4
done; done
