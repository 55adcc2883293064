#!/usr/bin/env bash
# Round-2 profile refresh (through gpurun): config 5 and NUTS ncu captures, bench lines.
out=gpurun_out
mkdir -p $out
python scripts/bench_c5.py 32768 3 2>&1 | tail -1 > $out/c5_1gpu.json; cat $out/c5_1gpu.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_big_hmc -c 1 -o $out/prof_c5 -f python scripts/bench_c5.py 2048 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_nuts.csv python scripts/bench_nuts.py 65536 128 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_nuts_doubling -s 8 -c 3 -o $out/prof_nuts -f python scripts/bench_nuts.py 65536 128 2 > /dev/null 2>&1
python scripts/bench_nuts.py 65536 128 40 2>&1 | tail -1 > $out/nuts40.json; cat $out/nuts40.json | cut -c1-200
timeout 600 python bench.py --workload nuts_funnel_65536x128 --steps 20 > $out/bench_c3.json 2> $out/bench_err.log; echo "c3 rc=$?"
timeout 600 python bench.py > $out/bench_default.json 2>> $out/bench_err.log; echo "default rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2>> $out/bench_err.log; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("bench_default", "bench_c3", "bench_ref"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), d.get("roofline", {}).get("frac"), d.get("cpu_baseline", {}))
    except Exception as e:
        print(f, "unreadable", e)
PY
ls -la $out/*.ncu-rep
