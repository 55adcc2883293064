#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -q -k "native_nuts" 2>&1 | grep -v "^  " | tail -5
timeout 300 python bench.py --workload nuts_funnel_65536x128 --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2>> gpurun_out/bench_c3.err
python -c "
import json; l=json.load(open('gpurun_out/bench_c3.json')); print('C3 block', l['value'], l['ms_per_step'], l['config']['ms_per_transition'], l['config']['mean_tree_size'], l['e2e'], l.get('cpu_baseline',{}).get('value'), l['clocks'])"
timeout 200 python scripts/nuts_decoupled.py 65536 128 64
tail -3 gpurun_out/bench_c3.err
