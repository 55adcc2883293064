#!/usr/bin/env bash
# 8-GPU weak scaling of BASELINE config 4 (gpurun --gpus 8): NUTS + shared window adaptation, one NCCL all-gather per step
out=gpurun_out
mkdir -p $out
for n in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540 + n)) bench.py --gpus $n --workload nuts_window_adaptation_512 --steps 2 --warmup 1 > $out/bench_c4_${n}gpu.json 2>> $out/bench8_err.log
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_c4_${n}gpu.json").read().strip().splitlines()[-1])
    print("c4 n_gpus", d["n_gpus"], "value %.3e ms/step %.1f ms/warm-up step %.3f" % (d["value"], d["ms_per_step"], d["config"]["ms_per_transition"]), d["clocks"].get("per_rank_sm_mhz"))
except Exception as e:
    print("c4 ${n} unreadable", e)
PY
done
grep -i "nranks\|NCCL INFO comm" $out/bench8_err.log | head -3
tail -3 $out/bench8_err.log
