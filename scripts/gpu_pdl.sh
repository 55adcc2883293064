#!/usr/bin/env bash
for pdl in 1 0 1 0; do
  BJX_GEMM_PDL=$pdl timeout 100 python scripts/gemm_check.py loop 2>&1 | tail -1 | sed "s/^/PDL=$pdl /"
done
timeout 300 python scripts/gemm_check.py quick 2>&1 | tail -2
BJX_GEMM_PDL=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PDL=1 bench ms/step %.2f' % d['ms_per_step'])"
BJX_GEMM_PDL=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PDL=0 bench ms/step %.2f' % d['ms_per_step'])"
