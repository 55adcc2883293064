#!/usr/bin/env bash
# Round-2 closing validation (one gpurun call): every GPU test, smoke(), the default bench line, the config-5 bench line and
# its ncu captures (launch list + --set full of the two-chains-per-CTA kernel).  Outputs under gpurun_out/.
set -u
out=gpurun_out
mkdir -p "$out"
timeout 900 python -m pytest tests -m gpu -q -x > "$out/pytest_gpu.log" 2>&1
echo "pytest rc=$? : $(grep -E 'passed|failed' "$out/pytest_gpu.log" | tail -1)"
grep -E '^(FAILED|ERROR)' "$out/pytest_gpu.log" | head -10
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > "$out/bench_default.json" 2> "$out/bench_err.log"
python bench.py --workload hmc_hier_logit_32768x10000_L20 --steps 5 --warmup 3 > "$out/bench_c5.json" 2>> "$out/bench_err.log"
python - <<'PY'
import json
for f in ("bench_default", "bench_c5"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print("%s: value %.3e  ms/step %.2f  e2e %.3e  roofline %.1f %s frac %.3f  cpu %.3e  clocks %s" % (
            f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["unit"],
            d["roofline"]["frac"], d.get("cpu_baseline", {}).get("value", float("nan")), d["clocks"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
python scripts/bench_user_target.py > "$out/user_target.json" 2>> "$out/bench_err.log"; tail -c 900 "$out/user_target.json"; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$out/launches_c5.csv" \
  python bench.py --workload hmc_hier_logit_32768x10000_L20 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_big2_hmc_hier -s 1 -c 1 -o "$out/prof_c5" -f \
  python bench.py --workload hmc_hier_logit_32768x10000_L20 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la "$out"/*.ncu-rep "$out"/launches_c5.csv 2>&1 | tail -3
