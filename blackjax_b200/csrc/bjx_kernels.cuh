// HMC / NUTS kernels: one warp owns one chain (see bjx_row.cuh).  Every kernel replaces an
// XLA-fused jaxpr region of the reference; citations are blackjax file:line.
#pragma once
#include "bjx_row.cuh"

namespace bjx {

constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;

struct InfoPtrs {  // device view of bjx_info
  float* acceptance_rate;
  uint8_t* is_accepted;
  uint8_t* is_divergent;
  uint8_t* is_turning;
  float* energy;
  int32_t* num_integration_steps;
  int32_t* num_trajectory_expansions;
  float* momentum;
  float* proposal_position;
  float* proposal_momentum;
};

// NaN-propagating min(x, 1)  (jnp.clip(x, max=1): proposal.py:155,225)
__device__ __forceinline__ float clip_max1(float x) { return x > 1.0f ? 1.0f : x; }

#define BJX_WARP_PROLOGUE()                                                   \
  const int lane = threadIdx.x & 31;                                          \
  const int chain = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);         \
  extern __shared__ float bjx_smem[];                                         \
  float* sm = bjx_smem + (size_t)(threadIdx.x >> 5) * P.D;                    \
  if (chain >= P.C) return;                                                   \
  const size_t roff = (size_t)chain * P.D;

// ---- hmc.init (hmc.py:90-92) ---------------------------------------------------------------------
template <class R, int TK, bool DM>
__global__ void __launch_bounds__(kThreads) k_init_state(Params P, const float* __restrict__ q_in,
                                                         float* __restrict__ logp_out, float* __restrict__ g_out) {
  BJX_WARP_PROLOGUE();
  Ctx<R, TK, true> c;  // metric not needed: DM=true variant holds no mass registers
  c.init(P, chain, lane, sm);
  float q[R::NS], g[R::NS], logp;
  R::load(q, q_in + roff, P.D, lane);
  c.value_and_grad(P, q, g, logp);
  R::store(g, g_out + roff, P.D, lane);
  if (lane == 0) logp_out[chain] = logp;
}

// ---- metric.sample_momentum (metrics.py:260-261) ----------------------------------------------------
template <class R, int TK, bool DM>
__global__ void __launch_bounds__(kThreads) k_sample_momentum(Params P, const uint32_t* __restrict__ keys,
                                                              float* __restrict__ p_out) {
  BJX_WARP_PROLOGUE();
  Ctx<R, TK_FUNNEL, DM> c;  // target-independent
  c.init(P, chain, lane, sm);
  float p[R::NS];
  Key k{keys[2 * chain], keys[2 * chain + 1]};
  c.sample_momentum(P, chain, k, p);
  R::store(p, p_out + roff, P.D, lane);
}

// ---- static_integration of velocity_verlet (trajectory.py:136-167, integrators.py:62-152) ----------
// n_steps == 1 is the HBM-roofline kernel: reads q,p,g and writes q,p,g = 24*D bytes per chain.
template <class R, int TK, bool DM>
__global__ void __launch_bounds__(kThreads) k_leapfrog(Params P, float* __restrict__ q_io, float* __restrict__ p_io,
                                                       float* __restrict__ logp_io, float* __restrict__ g_io,
                                                       int n_steps) {
  BJX_WARP_PROLOGUE();
  Ctx<R, TK, DM> c;
  float q[R::NS], p[R::NS], g[R::NS];
  R::load(q, q_io + roff, P.D, lane);
  R::load(p, p_io + roff, P.D, lane);
  R::load(g, g_io + roff, P.D, lane);
  c.init(P, chain, lane, sm);
  const float eps = P.eps_dev ? P.eps_dev[chain] : P.eps;
  float logp = 0.f;
  for (int i = 0; i < n_steps; ++i) c.leapfrog(P, q, p, g, logp, eps);
  R::store(q, q_io + roff, P.D, lane);
  R::store(p, p_io + roff, P.D, lane);
  R::store(g, g_io + roff, P.D, lane);
  if (lane == 0 && n_steps > 0) logp_io[chain] = logp;
}

// ---- hmc_energy (trajectory.py:730-750) ---------------------------------------------------------------
template <class R, int TK, bool DM>
__global__ void __launch_bounds__(kThreads) k_energy(Params P, const float* __restrict__ p_in,
                                                     const float* __restrict__ logp_in, float* __restrict__ e_out) {
  BJX_WARP_PROLOGUE();
  Ctx<R, TK_FUNNEL, DM> c;
  c.init(P, chain, lane, sm);
  float p[R::NS];
  R::load(p, p_in + roff, P.D, lane);
  const float k = c.kinetic(P, p);
  if (lane == 0) e_out[chain] = -logp_in[chain] + k;
}

// ---- metrics.is_turning on explicit momenta (metrics.py:272-304) -----------------------------------------
template <class R, int TK, bool DM>
__global__ void __launch_bounds__(kThreads) k_is_turning(Params P, const float* __restrict__ pl_in,
                                                         const float* __restrict__ pr_in,
                                                         const float* __restrict__ ps_in, uint8_t* __restrict__ out) {
  BJX_WARP_PROLOGUE();
  Ctx<R, TK_FUNNEL, DM> c;
  c.init(P, chain, lane, sm);
  float pl[R::NS], pr[R::NS], ps[R::NS];
  R::load(pl, pl_in + roff, P.D, lane);
  R::load(pr, pr_in + roff, P.D, lane);
  R::load(ps, ps_in + roff, P.D, lane);
  const bool t = c.is_turning(P, pl, pr, ps);
  if (lane == 0) out[chain] = t ? 1 : 0;
}

// ---- one whole HMC transition (hmc.py:279-312) with the chain row resident in registers -------------------
// key split -> momentum draw -> L leapfrogs -> energies -> Metropolis accept -> select.
// HBM traffic per transition: read q,g (8D) + conditional write q,g (8D); the L leapfrogs never
// touch HBM.
template <class R, int TK, bool DM>
__global__ void __launch_bounds__(kThreads) k_hmc_transition(Params P, const uint32_t* __restrict__ keys,
                                                             const float* q_in, const float* logp_in, const float* g_in,
                                                             float* q_out, float* logp_out, float* g_out, int L,
                                                             InfoPtrs info) {
  BJX_WARP_PROLOGUE();
  Ctx<R, TK, DM> c;
  float q[R::NS], p[R::NS], g[R::NS];
  R::load(q, q_in + roff, P.D, lane);
  R::load(g, g_in + roff, P.D, lane);
  c.init(P, chain, lane, sm);
  const Key rng{keys[2 * chain], keys[2 * chain + 1]};
  const Key key_momentum = fold_in(rng, 0u);    // jax.random.split(rng_key, 2)  hmc.py:299
  const Key key_integrator = fold_in(rng, 1u);
  c.sample_momentum(P, chain, key_momentum, p);  // hmc.py:302
  if (info.momentum) R::store(p, info.momentum + roff, P.D, lane);
  const float logp0 = logp_in[chain];
  const float e0 = -logp0 + c.kinetic(P, p);    // hmc.py:159
  const float eps = P.eps_dev ? P.eps_dev[chain] : P.eps;
  float logp = logp0;
  for (int i = 0; i < L; ++i) c.leapfrog(P, q, p, g, logp, eps);  // trajectory.py:165
#pragma unroll
  for (int s = 0; s < R::NS; ++s) p[s] = -1.0f * p[s];              // flip_momentum hmc.py:158
  const float e1 = -logp + c.kinetic(P, p);                         // hmc.py:160
  const float delta = safe_energy_diff(e0, e1);                     // hmc.py:161
  const bool is_div = (-delta) > P.div_thr;                         // hmc.py:162
  const float p_acc = clip_max1(expf(delta));                       // proposal.py:225
  const float u = uniform01(key_integrator);                        // proposal.py:226 (same key, no split)
  const bool acc = u < p_acc;
  if (info.proposal_position) R::store(q, info.proposal_position + roff, P.D, lane);
  if (info.proposal_momentum) R::store(p, info.proposal_momentum + roff, P.D, lane);
  if (acc) {
    R::store(q, q_out + roff, P.D, lane);
    R::store(g, g_out + roff, P.D, lane);
    if (lane == 0) logp_out[chain] = logp;
  } else if (q_out != q_in) {  // out-of-place call: carry the old state over
    R::load(q, q_in + roff, P.D, lane);
    R::load(g, g_in + roff, P.D, lane);
    R::store(q, q_out + roff, P.D, lane);
    R::store(g, g_out + roff, P.D, lane);
    if (lane == 0) logp_out[chain] = logp0;
  }
  if (lane == 0) {
    if (info.acceptance_rate) info.acceptance_rate[chain] = p_acc;
    if (info.is_accepted) info.is_accepted[chain] = acc;
    if (info.is_divergent) info.is_divergent[chain] = is_div;
    if (info.energy) info.energy[chain] = e1;
    if (info.num_integration_steps) info.num_integration_steps[chain] = L;
  }
}

// =====================================================================================================
// NUTS (nuts.py:223-321, trajectory.py:273-393,616-725, termination.py:31-106)
//
// Host C++ drives the doubling loop; per-chain tree state lives in this workspace.  A sub-tree is
// integrated IN PLACE on the tree endpoint it extends (the last leaf always becomes the new
// endpoint after the merge, trajectory.py:376-385,697-704), so a leaf costs one read and one write
// of (q,p,g) plus the momentum-sum / checkpoint traffic.
// =====================================================================================================
struct NutsWs {
  float *left_q, *left_p, *left_g, *right_q, *right_p, *right_g;  // [C,D] trajectory endpoints
  float *psum, *sub_psum;                                         // [C,D] momentum sums
  float *sub_prop_q, *sub_prop_g;                                 // [C,D] sub-tree proposal
  float *ckpt_p, *ckpt_sum;                                       // [C,depth,D] U-turn checkpoints
  float *left_logp, *right_logp, *h0;
  float *prop_energy, *prop_weight, *prop_slpa;
  float *sub_logp, *sub_energy, *sub_weight, *sub_slpa;
  int *n_states, *sub_n, *step;
  uint8_t *is_div, *is_turn, *sub_div, *sub_term, *run, *active;
  int8_t* dir;
  uint32_t *key_int, *traj_key, *prop_key;                        // [C,2]
  int* counters;                                                  // [depth+2 + extra] device counters
  int max_depth;                                                  // checkpoint capacity
};

// trajectory.py:642-655 for the doubling about to start (run by all lanes, lane 0 writes)
__device__ __forceinline__ void nuts_begin(const NutsWs& ws, int chain, int lane, int max_doublings, int* counter) {
  const int step = ws.step[chain];
  const bool run = (step < max_doublings) && !ws.is_div[chain] && !ws.is_turn[chain];  // trajectory.py:622-630
  if (lane == 0) {
    ws.run[chain] = run;
    ws.active[chain] = run;
    if (run) {
      const Key ki{ws.key_int[2 * chain], ws.key_int[2 * chain + 1]};
      const Key sub = fold_in(ki, (uint32_t)step);           // :645
      const Key dk = fold_in(sub, 0u);                       // split(subkey, 3)  :646
      const Key tk = fold_in(sub, 1u);
      const Key pk = fold_in(sub, 2u);
      ws.dir[chain] = (uniform01(dk) < 0.5f) ? 1 : -1;       // :650
      ws.traj_key[2 * chain] = tk.a; ws.traj_key[2 * chain + 1] = tk.b;
      ws.prop_key[2 * chain] = pk.a; ws.prop_key[2 * chain + 1] = pk.b;
      ws.sub_div[chain] = 0;
      ws.sub_term[chain] = 0;
      ws.sub_n[chain] = 0;
      atomicAdd(counter, 1);
    }
  }
}

// nuts.py:133-136,278-294: key split, momentum draw, initial proposal/trajectory, then begin(depth 0)
template <class R, int TK, bool DM>
__global__ void __launch_bounds__(kThreads) k_nuts_init(Params P, NutsWs ws, const uint32_t* __restrict__ keys,
                                                        const float* q_in, const float* logp_in, const float* g_in,
                                                        float* q_out, float* logp_out, float* g_out,
                                                        const float* mom_override, const uint32_t* keyint_override,
                                                        float* mom_out, int max_doublings) {
  BJX_WARP_PROLOGUE();
  Ctx<R, TK_FUNNEL, DM> c;
  c.init(P, chain, lane, sm);
  float q[R::NS], p[R::NS], g[R::NS];
  R::load(q, q_in + roff, P.D, lane);
  R::load(g, g_in + roff, P.D, lane);
  Key key_integrator;
  if (keyint_override) {
    key_integrator = Key{keyint_override[2 * chain], keyint_override[2 * chain + 1]};
    R::load(p, mom_override + roff, P.D, lane);
  } else {
    const Key rng{keys[2 * chain], keys[2 * chain + 1]};
    key_integrator = fold_in(rng, 1u);
    c.sample_momentum(P, chain, fold_in(rng, 0u), p);
  }
  if (mom_out) R::store(p, mom_out + roff, P.D, lane);
  const float logp0 = logp_in[chain];
  const float h0 = -logp0 + c.kinetic(P, p);
  R::store(q, ws.left_q + roff, P.D, lane);
  R::store(p, ws.left_p + roff, P.D, lane);
  R::store(g, ws.left_g + roff, P.D, lane);
  R::store(q, ws.right_q + roff, P.D, lane);
  R::store(p, ws.right_p + roff, P.D, lane);
  R::store(g, ws.right_g + roff, P.D, lane);
  R::store(p, ws.psum + roff, P.D, lane);
  if (q_out != q_in) {  // the running proposal lives in the output buffers
    R::store(q, q_out + roff, P.D, lane);
    R::store(g, g_out + roff, P.D, lane);
  }
  if (lane == 0) {
    if (q_out != q_in) logp_out[chain] = logp0;
    ws.left_logp[chain] = logp0;
    ws.right_logp[chain] = logp0;
    ws.h0[chain] = h0;
    ws.prop_energy[chain] = h0;
    ws.prop_weight[chain] = 0.f;
    ws.prop_slpa[chain] = -__int_as_float(0x7f800000);
    ws.n_states[chain] = 0;
    ws.step[chain] = 0;
    ws.is_div[chain] = 0;
    ws.is_turn[chain] = 0;
    ws.key_int[2 * chain] = key_integrator.a;
    ws.key_int[2 * chain + 1] = key_integrator.b;
  }
  __syncwarp();
  nuts_begin(ws, chain, lane, max_doublings, ws.counters + 0);
}

// One leaf of every active sub-tree (trajectory.py:318-355): leapfrog from the endpoint being
// extended, energy/weight, progressive uniform sampling, momentum sum, checkpoint store (even leaf)
// and iterative U-turn scan (termination.py:56-104).  leaf index i and its checkpoint range are the
// same for all chains, so they are kernel arguments computed on the host (termination.py:75-84).
template <class R, int TK, bool DM>
__global__ void __launch_bounds__(kThreads) k_nuts_leaf(Params P, NutsWs ws, int i, int idx_min, int idx_max,
                                                        int* active_counter) {
  BJX_WARP_PROLOGUE();
  if (!ws.active[chain]) return;
  Ctx<R, TK, DM> c;
  const int dir = ws.dir[chain];
  float* eq = dir > 0 ? ws.right_q : ws.left_q;
  float* ep = dir > 0 ? ws.right_p : ws.left_p;
  float* eg = dir > 0 ? ws.right_g : ws.left_g;
  float q[R::NS], p[R::NS], g[R::NS];
  R::load(q, eq + roff, P.D, lane);
  R::load(p, ep + roff, P.D, lane);
  R::load(g, eg + roff, P.D, lane);
  c.init(P, chain, lane, sm);
  const float eps = (float)dir * (P.eps_dev ? P.eps_dev[chain] : P.eps);  // direction * step_size  :323
  float logp;
  c.leapfrog(P, q, p, g, logp, eps);
  R::store(q, eq + roff, P.D, lane);
  R::store(p, ep + roff, P.D, lane);
  R::store(g, eg + roff, P.D, lane);
  const float e_new = -logp + c.kinetic(P, p);
  const float w_new = safe_energy_diff(ws.h0[chain], e_new);  // proposal.py:94-98
  const float slpa_new = fminf(w_new, 0.f);
  const bool is_div = (-w_new) > P.div_thr;                   // :325
  float ps[R::NS];
  bool take;
  float w_tot, slpa_tot;
  if (i == 0) {  // :329-334 first leaf is taken unconditionally
#pragma unroll
    for (int s = 0; s < R::NS; ++s) ps[s] = p[s];
    take = true;
    w_tot = w_new;
    slpa_tot = slpa_new;
  } else {  // :335-338 append + progressive uniform sampling (proposal.py:118-143)
    R::load(ps, ws.sub_psum + roff, P.D, lane);
#pragma unroll
    for (int s = 0; s < R::NS; ++s) ps[s] = ps[s] + p[s];
    const Key tk{ws.traj_key[2 * chain], ws.traj_key[2 * chain + 1]};
    const float w_old = ws.sub_weight[chain];
    const float p_accept = expit_f(w_new - w_old);
    take = uniform01(fold_in(tk, (uint32_t)i)) < p_accept;
    w_tot = logaddexp_f(w_old, w_new);
    slpa_tot = logaddexp_f(ws.sub_slpa[chain], slpa_new);
  }
  R::store(ps, ws.sub_psum + roff, P.D, lane);
  if (take) {
    R::store(q, ws.sub_prop_q + roff, P.D, lane);
    R::store(g, ws.sub_prop_g + roff, P.D, lane);
  }
  const size_t coff = (size_t)chain * ws.max_depth * P.D;
  if ((i & 1) == 0) {  // termination.py:66-72
    R::store(p, ws.ckpt_p + coff + (size_t)idx_max * P.D, P.D, lane);
    R::store(ps, ws.ckpt_sum + coff + (size_t)idx_max * P.D, P.D, lane);
  }
  bool turning = false;
  for (int k = idx_max; k >= idx_min && !turning; --k) {  // termination.py:96-103
    float cp[R::NS], cs[R::NS];
    R::load(cp, ws.ckpt_p + coff + (size_t)k * P.D, P.D, lane);
    R::load(cs, ws.ckpt_sum + coff + (size_t)k * P.D, P.D, lane);
#pragma unroll
    for (int s = 0; s < R::NS; ++s) cs[s] = ps[s] - cs[s] + cp[s];
    turning = c.is_turning(P, cp, p, cs);
  }
  if (lane == 0) {
    if (dir > 0) ws.right_logp[chain] = logp; else ws.left_logp[chain] = logp;
    if (take) {
      ws.sub_logp[chain] = logp;
      ws.sub_energy[chain] = e_new;
    }
    ws.sub_weight[chain] = w_tot;
    ws.sub_slpa[chain] = slpa_tot;
    ws.sub_n[chain] = i + 1;
    ws.sub_div[chain] = is_div;
    ws.sub_term[chain] = turning;
    const bool still = !(is_div || turning);
    ws.active[chain] = still;
    if (active_counter && still) atomicAdd(active_counter, 1);
  }
}

// End of a doubling (trajectory.py:672-717): proposal update (biased progressive sampling or only
// sum_log_p_accept), merge, full-trajectory U-turn, then begin the next doubling.
template <class R, int TK, bool DM>
__global__ void __launch_bounds__(kThreads) k_nuts_end(Params P, NutsWs ws, float* q_out, float* logp_out,
                                                       float* g_out, int max_doublings, int* next_counter) {
  BJX_WARP_PROLOGUE();
  if (!ws.run[chain]) return;
  Ctx<R, TK_FUNNEL, DM> c;
  c.init(P, chain, lane, sm);
  const bool sub_div = ws.sub_div[chain], sub_term = ws.sub_term[chain];
  const bool bad = sub_div || sub_term;
  const float pw = ws.prop_weight[chain], sw = ws.sub_weight[chain];
  const float new_slpa = logaddexp_f(ws.prop_slpa[chain], ws.sub_slpa[chain]);
  bool take = false;
  if (!bad) {  // proposal.py:146-176
    const float p_accept = clip_max1(expf(sw - pw));
    const Key pk{ws.prop_key[2 * chain], ws.prop_key[2 * chain + 1]};
    take = uniform01(pk) < p_accept;
  }
  if (take) {
    float t[R::NS];
    R::load(t, ws.sub_prop_q + roff, P.D, lane);
    R::store(t, q_out + roff, P.D, lane);
    R::load(t, ws.sub_prop_g + roff, P.D, lane);
    R::store(t, g_out + roff, P.D, lane);
  }
  float pl[R::NS], pr[R::NS], ps[R::NS];
  {
    float sp[R::NS];
    R::load(ps, ws.psum + roff, P.D, lane);
    R::load(sp, ws.sub_psum + roff, P.D, lane);
#pragma unroll
    for (int s = 0; s < R::NS; ++s) ps[s] = ps[s] + sp[s];  // merge_trajectories  trajectory.py:102-125
    R::store(ps, ws.psum + roff, P.D, lane);
  }
  R::load(pl, ws.left_p + roff, P.D, lane);
  R::load(pr, ws.right_p + roff, P.D, lane);
  const bool turning = c.is_turning(P, pl, pr, ps);  // :706-710
  if (lane == 0) {
    if (take) {
      logp_out[chain] = ws.sub_logp[chain];
      ws.prop_energy[chain] = ws.sub_energy[chain];
    }
    if (!bad) ws.prop_weight[chain] = logaddexp_f(pw, sw);
    ws.prop_slpa[chain] = new_slpa;
    ws.n_states[chain] += ws.sub_n[chain];
    ws.step[chain] += 1;
    ws.is_div[chain] = sub_div;
    ws.is_turn[chain] = sub_term || turning;  // :715
  }
  __syncwarp();
  nuts_begin(ws, chain, lane, max_doublings, next_counter);
}

// nuts.py:303-319: acceptance_rate = exp(sum_log_p_accept) / num_states and the NUTSInfo scalars
static __global__ void k_nuts_finish(int C, NutsWs ws, InfoPtrs info) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= C) return;
  const int n = ws.n_states[chain];
  if (info.acceptance_rate) info.acceptance_rate[chain] = expf(ws.prop_slpa[chain]) / (float)n;
  if (info.is_divergent) info.is_divergent[chain] = ws.is_div[chain];
  if (info.is_turning) info.is_turning[chain] = ws.is_turn[chain];
  if (info.energy) info.energy[chain] = ws.prop_energy[chain];
  if (info.num_integration_steps) info.num_integration_steps[chain] = n;
  if (info.num_trajectory_expansions) info.num_trajectory_expansions[chain] = ws.step[chain];
}

}  // namespace bjx
