"""Teacher-forced step-by-step comparison of per-chain window adaptation (device vs oracle)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import numpy as np, torch
import blackjax_b200 as bj
from blackjax_b200 import targets as T
from blackjax_b200._lib import check, lib, ptr
from blackjax_b200._engine import get_engine
from oracle import adaptation as oa, hmc as ohmc, prng as oprng, targets as ot
F = np.float32
DEV = "cuda:0"
D, C, T_ = 8, 16, 60
scale = np.logspace(-0.5, 0.5, D)
tgt, otgt = T.DiagGaussian(scale), ot.DiagGaussian(scale)
rs = np.random.default_rng(9)
q = rs.standard_normal((C, D)).astype(F)
ckeys = oprng.split(oprng.key(77), C)
keys_np = oprng.split(ckeys, T_)      # [C,T,2]
tk = lambda k: torch.from_numpy(np.ascontiguousarray(k).view(np.int32)).to(DEV).view(torch.uint32)
tf = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)
npy = lambda t: t.detach().cpu().numpy()
kern = bj.hmc.build_kernel()
state = bj.hmc.init(tf(q), tgt)
eng = get_engine(state.position, tgt)
da_state = torch.empty(C, 5, device=DEV)
eps = torch.full((C,), 1.0, device=DEV)
check(lib().bjx_da_init(eng.h, ptr(da_state), ptr(eps), ptr(eps)), eng.h)
imm = torch.ones(C, D, device=DEV)
w_mean = torch.zeros(C, D, device=DEV); w_m2 = torch.zeros(C, D, device=DEV); w_n = 0
odas = [oa.da_init(1.0) for _ in range(C)]
owfs = [oa.welford_init(D) for _ in range(C)]
sched = bj.build_schedule(T_)
for t, (stage, wend) in enumerate(sched):
    kt = tk(keys_np[:, t])
    # oracle from the DEVICE's current inputs (teacher forced)
    ost = ohmc.HMCState(npy(state.position), npy(state.logdensity), npy(state.logdensity_grad))
    oeps, oimm = npy(eps).copy(), npy(imm).copy()
    onew, oinfo = ohmc.hmc_kernel(keys_np[:, t], ost, otgt, oeps, oa._PerChainDiag(oimm), 8)
    state, info = kern(kt, state, tgt, eps, imm, 8)
    torch.cuda.synchronize()
    dpos = np.max(np.abs(npy(state.position) - onew.position))
    dacc = np.max(np.abs(npy(info.acceptance_rate) - oinfo.acceptance_rate))
    flips = int((npy(info.is_accepted) != oinfo.is_accepted).sum())
    # DA teacher-forced: oracle DA from device DA state
    st_np = npy(da_state).copy()
    odas = [oa.DAState(F(r[0]), F(r[1]), int(r[2]), F(r[3]), F(r[4])) for r in st_np]
    acc_np = npy(info.acceptance_rate)
    if stage == 1:
        w_n += 1
        check(lib().bjx_welford_update(eng.h, ptr(state.position), ptr(w_mean), ptr(w_m2), w_n), eng.h)
    check(lib().bjx_da_update(eng.h, ptr(da_state), ptr(info.acceptance_rate), 0.8, ptr(eps)), eng.h)
    odas = [oa.da_update(s, a, 0.8) for s, a in zip(odas, acc_np)]
    deps = np.max(np.abs(npy(eps) / np.array([np.exp(s.log_step_size) for s in odas]) - 1))
    msg = f"t={t:2d} stage={stage} end={int(wend)} dpos={dpos:.2e} dacc={dacc:.2e} flips={flips} rel_deps={deps:.2e} eps0={float(eps[0]):.5f}"
    if wend:
        new_imm = torch.empty_like(imm)
        m2_np, mean_np = npy(w_m2).copy(), npy(w_mean).copy()
        check(lib().bjx_welford_final(eng.h, ptr(w_mean), ptr(w_m2), w_n, ptr(new_imm)), eng.h)
        oimm_new = np.stack([oa.welford_final(oa.Welford(mean_np[c], m2_np[c], w_n)) for c in range(C)])
        msg += f" imm_rel={np.max(np.abs(npy(new_imm)/oimm_new-1)):.2e}"
        imm = new_imm; w_n = 0
        check(lib().bjx_da_reset(eng.h, ptr(da_state), ptr(eps)), eng.h)
    print(msg)
fin = torch.empty(C, device=DEV)
check(lib().bjx_da_final(eng.h, ptr(da_state), ptr(fin)), eng.h)
print("final eps dev", npy(fin)[:4])
# now the library's own run() and the oracle's run
warm = bj.window_adaptation(bj.hmc, tgt, num_integration_steps=8)
(st2, params), _ = warm.run(tk(ckeys), tf(q), T_)
okern = lambda k, s, t, e, m, **kw: ohmc.hmc_kernel(k, s, t, e, m, 8)
ost, oeps, oimm, ohist = oa.window_adaptation_run(okern, otgt, ckeys, q, T_)
print("run() eps  ", npy(params["step_size"])[:4])
print("oracle eps ", oeps[:4])
print("oracle hist t=0..5 chain0", ohist[:6, 0], " t=53..59", ohist[53:, 0])
