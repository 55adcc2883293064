"""Cost of the plug-in route: the diagonal Gaussian as a built-in target and as a user-defined target (the same arithmetic
compiled in another translation unit / shared library, launched through bjx_plugin_launch) at 65536 x 1024, HMC L = 50;
and the reference's regression posterior (tests/mcmc/test_sampling.py:103-111, N = 1000 observations) under NUTS.
    python scripts/bench_user_target.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blackjax_b200 as bj
from blackjax_b200 import plugin

dev = torch.device("cuda", 0)
C, D, L = 65536, 1024, 50
s = np.logspace(-0.5, 0.5, D)
inv_var = (1.0 / s ** 2).astype(np.float32)
targets = {"builtin": bj.targets.DiagGaussian(s),
           "plugin": bj.targets.UserTarget(D, plugin.read_example("diag_gaussian"), inv_var, name="diag_gaussian",
                                           dense_metric=False, general_integrators=False)}
imm = torch.from_numpy((s ** 2).astype(np.float32)).to(dev)
q0 = torch.randn(C, D, device=dev) * torch.from_numpy(s.astype(np.float32)).to(dev)
keys = bj.random.split(bj.random.key(0, dev), 16)
res = {}
for name, tgt in targets.items():
    kern = bj.hmc.build_kernel(inplace=True)
    st = bj.hmc.init(q0.clone(), tgt)
    for t in range(3):
        st, info = kern(keys[t], st, tgt, 0.1, imm, L)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(3, 13):
        st, info = kern(keys[t], st, tgt, 0.1, imm, L)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    res[name] = {"ms_per_transition": ms, "leapfrogs_per_s": C * L / (ms * 1e-3), "final_position_sum": float(st.position.double().sum())}
res["bit_identical"] = res["builtin"]["final_position_sum"] == res["plugin"]["final_position_sum"]

# the reference's regression posterior: 65536 chains, NUTS, adapted by a short shared-free warm-up on 1024 chains
rs = np.random.default_rng(42)
x = rs.standard_normal((1000, 1)).astype(np.float32)
y = (3 * x[:, 0] + rs.standard_normal(1000)).astype(np.float32)
tgt = bj.targets.LinearRegression(x, y)
Cr = 65536
q0 = torch.tensor([[0.0, 3.0]], device=dev).repeat(Cr, 1).contiguous() + 0.02 * torch.randn(Cr, 2, device=dev)
imm = torch.full((2,), 1e-3, device=dev)
st = bj.nuts.init(q0, tgt)
_, _, acc, nint = bj.sample_nuts_native(bj.random.key(1, dev), st, tgt, 0.5, imm, 4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
st2, _, acc, nint = bj.sample_nuts_native(bj.random.key(2, dev), st, tgt, 0.5, imm, 16, keep_history=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
n_leap = int(nint.sum())
res["linear_regression_nuts"] = {"chains": Cr, "observations": 1000, "transitions": 16, "ms_per_transition": ms / 16,
                                 "leapfrogs_per_s": n_leap / (ms * 1e-3), "mean_tree": n_leap / (16 * Cr),
                                 "gradient_terms_per_s": n_leap * 1000 / (ms * 1e-3), "mean_acceptance": float(acc.mean())}
print(json.dumps(res))
