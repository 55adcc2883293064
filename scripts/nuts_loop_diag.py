"""Why does bench.py's C3 loop take 2.0 ms per transition when scripts/bench_nuts.py takes 1.2?  Variants of the loop."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blackjax_b200 as bj
DEV = "cuda:0"
C, D = 65536, 128
tgt = bj.targets.Funnel(D)
imm = torch.ones(D, device=DEV)
q = 0.1 * bj.random.normal(bj.random.split(bj.random.key(7, DEV), C), (D,))
keys = bj.random.split(bj.random.key(0, DEV), 200)


def run(name, sync, accumulate, offset_kw, n=20, warm=3):
    kern = bj.nuts.build_kernel(inplace=True, max_tree_depth=10, **offset_kw)
    st = bj.nuts.init(q.clone(), tgt)
    lf = torch.zeros(1, dtype=torch.int64, device=DEV)
    for t in range(warm):
        st, info = kern(keys[t], st, tgt, 0.1, imm, 10)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for t in range(warm, warm + n):
        st, info = kern(keys[t], st, tgt, 0.1, imm, 10)
        if accumulate:
            lf += info.num_integration_steps.sum()
        if sync:
            torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1) / n:.3f} ms/transition (wall {(time.perf_counter() - w0) * 1e3 / n:.3f}) mean tree {float(info.num_integration_steps.float().mean()):.1f}", flush=True)


run("sync per step", True, False, {})
run("no sync", False, False, {})
run("no sync + accumulate", False, True, {})
run("no sync + accumulate + chain_offset", False, True, {"chain_offset": 0})
run("no sync, 40 steps after 20 warm", False, False, {}, n=40, warm=20)
