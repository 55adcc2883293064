#!/usr/bin/env bash
# Where does the fused product spend its time?  BJX_GEMM_DEBUG bits: 1 no Cin loads, 2 no plane stores, 4 no Y stores, 8 no
# epilogue math, 16 cycle counters of the single-thread roles; BJX_GEMM_VARIANT: shared-memory plan (bjx_gemm.cu).
for v in 0 2 3; do for d in 16 17 18 20 22 31; do
  BJX_GEMM_VARIANT=$v BJX_GEMM_DEBUG=$d timeout 100 python scripts/gemm_check.py loop 2>&1 | tail -4
done; done
