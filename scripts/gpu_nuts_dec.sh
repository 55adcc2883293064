#!/bin/bash
# decoupled NUTS sampler + GHMC/MEADS: parity tests, bench line, timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -q -k "native_nuts" 2>&1 | grep -v "^  " | tail -15
timeout 300 python -m pytest tests/test_gpu_meads.py -q 2>&1 | grep -v "^  " | tail -15
timeout 300 python bench.py --workload nuts_funnel_65536x128 --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python -c "
import json; l=json.load(open('gpurun_out/bench_c3.json')); print('C3', l['value'], l['ms_per_step'], l['config']['ms_per_transition'], l['config']['mean_tree_size'], l['e2e'], l.get('cpu_baseline',{}).get('value'))"
BJX_BENCH_NUTS_BLOCK=1 BJX_BENCH_NO_CLOCKS=1 timeout 300 python bench.py --workload nuts_funnel_65536x128 --steps 20 --warmup 40 --no-cpu-baseline > gpurun_out/bench_c3_stepwise.json 2>> gpurun_out/bench_c3.err
python -c "
import json; l=json.load(open('gpurun_out/bench_c3_stepwise.json')); print('C3 stepwise', l['value'], l['ms_per_step'], l['config']['ms_per_transition'], l['config']['mean_tree_size'])"
timeout 200 python scripts/bench_meads.py 4096 128 4 200
timeout 200 python scripts/bench_meads.py 32768 512 4 50
tail -3 gpurun_out/bench_c3.err
