#!/usr/bin/env bash
# Round-2 validation (through gpurun): GPU tests, smoke, default bench + the C3 / C4 workloads.
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $out/pytest_gpu.log 2>&1
echo "pytest rc=$? : $(grep -E 'passed|failed' $out/pytest_gpu.log | tail -1)"
grep -E "^FAILED|^ERROR|worst elementwise" $out/pytest_gpu.log | head -60
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -s > $out/pytest_round2.log 2>&1
grep -E "config-2 shape|of chains identical|passed|failed" $out/pytest_round2.log | head -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > $out/bench_default.json 2> $out/bench_err.log; echo "bench default rc=$?"
timeout 600 python bench.py --workload nuts_funnel_65536x128 --steps 20 > $out/bench_c3.json 2>> $out/bench_err.log; echo "bench c3 rc=$?"
timeout 900 python bench.py --workload nuts_window_adaptation_512 --steps 2 --warmup 3 > $out/bench_c4.json 2>> $out/bench_err.log; echo "bench c4 rc=$?"
tail -5 $out/bench_err.log
python - <<'PY'
import json
for f in ("bench_default", "bench_c3", "bench_c4"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.3e ms/step %.3f e2e %.3e roofline %.1f %s frac %.3f clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["unit"], d["roofline"]["frac"], d["clocks"]))
        print("   config", {k: v for k, v in d["config"].items() if k in ("mean_tree_size", "ms_per_transition", "mean_acceptance")}, "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
