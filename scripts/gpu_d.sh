#!/usr/bin/env bash
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1
echo "pytest rc=$? : $(grep -E 'passed|failed' $out/pytest_gpu.log | tail -1)"; grep -E "^FAILED|^E  |worst" $out/pytest_gpu.log | head -20
for nc in "" 1; do
  BJX_BENCH_NO_CLOCKS=$nc timeout 300 python bench.py --workload nuts_funnel_65536x128 --steps 20 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c3 no_clocks=$nc', 'ms/step %.3f' % d['ms_per_step'], d['clocks'])"
done
timeout 600 python bench.py --workload nuts_window_adaptation_512 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4', 'ms/step %.1f value %.3e' % (d['ms_per_step'], d['value']), d['config']['ms_per_transition'])"
