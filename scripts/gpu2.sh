#!/usr/bin/env bash
# two-GPU checks (gpurun --gpus 2): G=1 vs G=2 invariance of the sharded warm-up, C4 / C2 scaling lines
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "bit_identical" -s > $out/pytest_2gpu.log 2>&1
echo "2-GPU test rc=$? : $(grep -E 'passed|failed|skipped' $out/pytest_2gpu.log | tail -1)"
grep -E "Error|error|assert" $out/pytest_2gpu.log | head -10
timeout 900 python bench.py --workload nuts_window_adaptation_512 --steps 2 --warmup 1 --no-cpu-baseline > $out/bench_c4_1gpu.json 2> $out/bench2_err.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload nuts_window_adaptation_512 --steps 2 --warmup 1 > $out/bench_c4_2gpu.json 2>> $out/bench2_err.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 > $out/bench_dense_2gpu.json 2>> $out/bench2_err.log
tail -3 $out/bench2_err.log
python - <<'PY'
import json
for f in ("bench_c4_1gpu", "bench_c4_2gpu", "bench_dense_2gpu"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), "ms/transition", d["config"].get("ms_per_transition"), d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
