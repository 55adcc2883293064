#!/usr/bin/env bash
# One-call GPU validation (run through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_validate.sh [quick|full|profile]'
# quick   : pytest -m gpu, smoke(), default bench line                      (~2.5 min of box time)
# full    : quick + reference arm + diag workload + config-3 NUTS bench      (~4 min)
# profile : full + ncu launch list of the default bench + ncu --set full of the GEMM and the split kernel
# Everything lands in gpurun_out/ (merged back by gpurun); nothing here reads /root/reference.
set -u
mode="${1:-quick}"
out=gpurun_out
mkdir -p "$out"
timeout 1200 python -m pytest tests -m gpu -q > "$out/pytest_gpu.log" 2>&1
echo "pytest rc=$? : $(grep -E 'passed|failed' "$out/pytest_gpu.log" | tail -1)"
grep -E '^FAILED' "$out/pytest_gpu.log" | head -10
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > "$out/bench_default.json" 2> "$out/bench_err.log"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("default: value %.3e  ms/step %.2f  e2e %.3e  roofline %.1f %s frac %.3f  clocks %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["unit"],
    d["roofline"]["frac"], d["clocks"]))
PY
[ "$mode" = quick ] && exit 0
python bench.py --impl reference --steps 2 --warmup 1 > "$out/bench_ref.json" 2>> "$out/bench_err.log"
python bench.py --workload hmc_diag_gaussian_65536x1024_L50 > "$out/bench_diag.json" 2>> "$out/bench_err.log"
python scripts/bench_nuts.py 65536 128 40 > "$out/nuts40.jsonl" 2>&1
python bench.py --workload nuts_funnel_65536x128 --steps 20 > "$out/bench_c3.json" 2>> "$out/bench_err.log"
python bench.py --workload nuts_window_adaptation_512 --steps 2 --warmup 1 > "$out/bench_c4.json" 2>> "$out/bench_err.log"
python - <<'PY'
import json
r = json.load(open("gpurun_out/bench_ref.json")); d = json.load(open("gpurun_out/bench_diag.json"))
print("reference arm: %.3e (%s)" % (r["value"], r["cpu_baseline"]["sample"]))
print("diag: value %.3e  ms/step %.3f  e2e %.3e  HBM frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
print("nuts:", open("gpurun_out/nuts40.jsonl").read().strip().splitlines()[-1][:160])
PY
[ "$mode" = full ] && exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file "$out/launches_dense.csv" \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_f16x3 -s 60 -c 3 -o "$out/prof_gemm" -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la "$out"/*.ncu-rep "$out"/launches_dense.csv
