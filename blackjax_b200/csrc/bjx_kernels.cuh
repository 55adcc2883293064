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
  const Key k = chain_key(P, keys, chain);
  c.sample_momentum(P, chain, k, p);
  R::store(p, p_out + roff, P.D, lane);
}

// ---- static_integration of velocity_verlet (trajectory.py:136-167, integrators.py:62-152) ----------
// n_steps == 1 is the HBM-roofline kernel: reads q,p,g and writes q,p,g = 24*D bytes per chain.
template <class R, int TK, bool DM, bool GEN>
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
  for (int i = 0; i + 1 < n_steps; ++i) c.template step<GEN, false>(P, q, p, g, logp, eps);
  if (n_steps > 0) c.template step<GEN, true>(P, q, p, g, logp, eps);
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
template <class R, int TK, bool DM, bool GEN>
__global__ void __launch_bounds__(kThreads) k_hmc_transition(Params P, const uint32_t* __restrict__ keys,
                                                             const float* q_in, const float* logp_in, const float* g_in,
                                                             float* q_out, float* logp_out, float* g_out, int L_all,
                                                             InfoPtrs info) {
  BJX_WARP_PROLOGUE();
  Ctx<R, TK, DM> c;
  float q[R::NS], p[R::NS], g[R::NS];
  R::load(q, q_in + roff, P.D, lane);
  R::load(g, g_in + roff, P.D, lane);
  c.init(P, chain, lane, sm);
  const int L = P.steps_dev ? P.steps_dev[chain] : L_all;  // dynamic HMC: this chain's own trajectory length
  const Key rng = chain_key(P, keys, chain);
  const Key key_momentum = fold_in(rng, 0u);    // jax.random.split(rng_key, 2)  hmc.py:299
  const Key key_integrator = fold_in(rng, 1u);
  c.sample_momentum(P, chain, key_momentum, p);  // hmc.py:302
  if (info.momentum) R::store(p, info.momentum + roff, P.D, lane);
  const float logp0 = logp_in[chain];
  const float e0 = -logp0 + c.kinetic(P, p);    // hmc.py:159
  const float eps = P.eps_dev ? P.eps_dev[chain] : P.eps;
  float logp = logp0;
  for (int i = 0; i + 1 < L; ++i) c.template step<GEN, false>(P, q, p, g, logp, eps);  // trajectory.py:165
  if (L > 0) c.template step<GEN, true>(P, q, p, g, logp, eps);
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

// ---- generalized HMC (ghmc.py:118-189): persistent momentum, ONE leapfrog, non-reversible slice acceptance ------------
// In place on (q, p, logp, g, slice).  alpha / delta: per chain when the pointers are set.  Chains in [skip_begin,
// skip_end) draw and integrate like every other chain but keep their state (MEADS' frozen fold,
// meads_adaptation.py:639-650).  noise_dev: the values noise_fn(key_noise) of ghmc.py:172 per chain, evaluated by the caller
// (key_noise = split(rng_key)[1]); nullptr = the reference default (identically 0).
struct GhmcArgs {
  float* p_io;
  float* slice_io;
  float alpha, delta;
  const float* alpha_dev;
  const float* delta_dev;
  int param_group;  // step_size_dev / alpha_dev / delta_dev are indexed by chain / param_group (MEADS: one entry per fold)
  int skip_begin, skip_end;
  const float* noise_dev;
};

template <class R, int TK, bool DM, bool GEN>
__global__ void __launch_bounds__(kThreads) k_ghmc_transition(Params P, const uint32_t* __restrict__ keys, float* q_io,
                                                              float* logp_io, float* g_io, GhmcArgs A, InfoPtrs info) {
  BJX_WARP_PROLOGUE();
  Ctx<R, TK, DM> c;
  float q[R::NS], p[R::NS], g[R::NS], z[R::NS];
  R::load(q, q_io + roff, P.D, lane);
  R::load(g, g_io + roff, P.D, lane);
  R::load(p, A.p_io + roff, P.D, lane);
  c.init(P, chain, lane, sm);
  const Key rng = chain_key(P, keys, chain);
  const Key key_momentum = fold_in(rng, 0u);  // ghmc.py:169 (key_noise = fold_in(rng, 1): see noise_dev)
  const int gi = chain / A.param_group;
  const float alpha = A.alpha_dev ? A.alpha_dev[gi] : A.alpha;
  const float delta_s = A.delta_dev ? A.delta_dev[gi] : A.delta;
  c.sample_momentum(P, chain, key_momentum, z);
  const float keep = sqrtf(1.0f - alpha), mix = sqrtf(alpha);
#pragma unroll
  for (int s = 0; s < R::NS; ++s) p[s] = __fadd_rn(__fmul_rn(p[s], keep), __fmul_rn(mix, z[s]));  // ghmc.py:205-211
  if (info.momentum) R::store(p, info.momentum + roff, P.D, lane);
  float sl = A.slice_io[chain];
  {  // ((slice + 1 + delta + noise) % 2) - 1   ghmc.py:172; jnp.remainder = fmod + sign fix-up
    const float x = ((sl + 1.0f) + delta_s) + (A.noise_dev ? A.noise_dev[chain] : 0.0f);
    float r = fmodf(x, 2.0f);
    if (r != 0.0f && r < 0.0f) r += 2.0f;
    sl = r - 1.0f;
  }
  const float logp0 = logp_io[chain];
  const float e0 = -logp0 + c.kinetic(P, p);
  const float eps = P.eps_dev ? P.eps_dev[gi] : P.eps;
  float logp = logp0;
  float p0[R::NS];
#pragma unroll
  for (int s = 0; s < R::NS; ++s) p0[s] = p[s];
  c.template step<GEN, true>(P, q, p, g, logp, eps);
#pragma unroll
  for (int s = 0; s < R::NS; ++s) p[s] = -1.0f * p[s];                      // flip_momentum hmc.py:158
  const float e1 = -logp + c.kinetic(P, p);
  const float delta = safe_energy_diff(e0, e1);
  const bool is_div = (-delta) > P.div_thr;
  const float p_acc = clip_max1(expf(delta));                               // proposal.py:253
  const bool acc = logf(fabsf(sl)) <= delta;                                // proposal.py:254
  const float af = acc ? 1.0f : 0.0f;
  const float sl_next = sl * (expf(-delta) * af + (1.0f - af));             // proposal.py:255 (inf * 0 = NaN kept)
  if (info.proposal_position) R::store(q, info.proposal_position + roff, P.D, lane);
  if (info.proposal_momentum) R::store(p, info.proposal_momentum + roff, P.D, lane);
  if (lane == 0) {
    if (info.acceptance_rate) info.acceptance_rate[chain] = p_acc;
    if (info.is_accepted) info.is_accepted[chain] = acc;
    if (info.is_divergent) info.is_divergent[chain] = is_div;
    if (info.energy) info.energy[chain] = e1;
    if (info.num_integration_steps) info.num_integration_steps[chain] = 1;
  }
  if (chain >= A.skip_begin && chain < A.skip_end) return;
  if (acc) {  // the sampled state is flipped once more (ghmc.py:178): +p of the integrator
#pragma unroll
    for (int s = 0; s < R::NS; ++s) p[s] = -1.0f * p[s];
    R::store(q, q_io + roff, P.D, lane);
    R::store(g, g_io + roff, P.D, lane);
    R::store(p, A.p_io + roff, P.D, lane);
    if (lane == 0) logp_io[chain] = logp;
  } else {
#pragma unroll
    for (int s = 0; s < R::NS; ++s) p0[s] = -1.0f * p0[s];
    R::store(p0, A.p_io + roff, P.D, lane);
  }
  if (lane == 0) A.slice_io[chain] = sl_next;
}

// ---- multinomial HMC (hmc.py:181-248 + trajectory.py:170-232 static_progressive_integration) ---------------
// Same trajectory as k_hmc_transition, but every leaf competes through progressive uniform sampling
// (proposal.py:118-143) with step key fold_in(key_integrator, i); there is no Metropolis rejection.  The
// running proposal lives in the output buffers (it is rewritten ~log L times), the moving state in registers.
template <class R, int TK, bool DM, bool GEN>
__global__ void __launch_bounds__(kThreads) k_mhmc_transition(Params P, const uint32_t* __restrict__ keys,
                                                              const float* q_in, const float* logp_in, const float* g_in,
                                                              float* q_out, float* logp_out, float* g_out, int L_all,
                                                              InfoPtrs info) {
  BJX_WARP_PROLOGUE();
  Ctx<R, TK, DM> c;
  float q[R::NS], p[R::NS], g[R::NS];
  R::load(q, q_in + roff, P.D, lane);
  R::load(g, g_in + roff, P.D, lane);
  c.init(P, chain, lane, sm);
  const int L = P.steps_dev ? P.steps_dev[chain] : L_all;
  const Key rng = chain_key(P, keys, chain);
  const Key key_integrator = fold_in(rng, 1u);   // hmc.py:299
  c.sample_momentum(P, chain, fold_in(rng, 0u), p);
  if (info.momentum) R::store(p, info.momentum + roff, P.D, lane);
  const float logp0 = logp_in[chain];
  const float h0 = -logp0 + c.kinetic(P, p);     // trajectory.py:211
  if (q_out != q_in) {                           // init_proposal = the initial state (:212)
    R::store(q, q_out + roff, P.D, lane);
    R::store(g, g_out + roff, P.D, lane);
  }
  if (info.proposal_momentum) R::store(p, info.proposal_momentum + roff, P.D, lane);
  const float eps = P.eps_dev ? P.eps_dev[chain] : P.eps;
  float logp = logp0, prop_logp = logp0, prop_energy = h0;
  float weight = 0.f, slpa = -__int_as_float(0x7f800000);
  bool any_div = false;
  float u_batch = 0.f;
  for (int i = 0; i < L; ++i) {
    // step key fold_in(rng_key, i) (trajectory.py:216): lane l draws the uniform of step i+l, one pass per 32 steps
    if ((i & 31) == 0) u_batch = uniform01(fold_in(key_integrator, (uint32_t)(i + lane)));
    const float u_step = __shfl_sync(0xffffffffu, u_batch, i & 31);
    c.template step<GEN, true>(P, q, p, g, logp, eps);
    const float e_new = -logp + c.kinetic(P, p);
    const float w_new = safe_energy_diff(h0, e_new);          // proposal.py:94-98
    any_div = any_div || ((-w_new) > P.div_thr);              // trajectory.py:220-221
    const float p_accept = expit_f(w_new - weight);           // proposal.py:122
    const bool take = u_step < p_accept;
    if (take) {
      R::store(q, q_out + roff, P.D, lane);
      R::store(g, g_out + roff, P.D, lane);
      if (info.proposal_momentum) R::store(p, info.proposal_momentum + roff, P.D, lane);
      prop_logp = logp;
      prop_energy = e_new;
    }
    const float la = logaddexp_f(lane == 0 ? weight : slpa, lane == 0 ? w_new : fminf(w_new, 0.f));
    weight = __shfl_sync(0xffffffffu, la, 0);
    slpa = __shfl_sync(0xffffffffu, la, 1);
  }
  if (info.proposal_position && info.proposal_position != q_out) {  // HMCInfo.proposal = the selected state
    R::load(q, q_out + roff, P.D, lane);
    R::store(q, info.proposal_position + roff, P.D, lane);
  }
  if (lane == 0) {
    logp_out[chain] = prop_logp;
    if (info.acceptance_rate) info.acceptance_rate[chain] = expf(slpa) / (float)L;  // hmc.py:232
    if (info.is_accepted) info.is_accepted[chain] = 1;
    if (info.is_divergent) info.is_divergent[chain] = any_div;
    if (info.energy) info.energy[chain] = prop_energy;
    if (info.num_integration_steps) info.num_integration_steps[chain] = L;
  }
}

// =====================================================================================================
// NUTS (nuts.py:223-321, trajectory.py:273-393,616-725, termination.py:31-106)
//
// Chains never interact, so nothing forces them through the tree in lock step.  Host C++ drives the
// doubling loop (one launch per doubling over the chains that are still expanding, compacted into an
// index list); inside a launch each warp integrates its chain's WHOLE sub-tree -- up to 2^d leapfrog
// leaves -- with the moving endpoint (q, p, grad) and the sub-tree momentum sum resident in
// registers, stopping early on divergence / U-turn.  A sub-tree is integrated IN PLACE on the tree
// endpoint it extends (the last leaf always becomes the new endpoint after the merge,
// trajectory.py:376-385,697-704).  Per leaf the only global traffic is the U-turn checkpoint row
// (store on even leaves, loads on odd leaves) and the proposal row when the multinomial draw accepts.
// =====================================================================================================
struct NutsWs {
  float *left_q, *left_p, *left_g, *right_q, *right_p, *right_g;  // [C,D] trajectory endpoints
  float* psum;                                                    // [C,D] trajectory momentum sum
  float *sub_prop_q, *sub_prop_g;                                 // [C,D] sub-tree proposal
  float *ckpt_p, *ckpt_sum;                                       // [C,depth,D] U-turn checkpoints
  float *left_logp, *right_logp, *h0;
  float *prop_energy, *prop_weight, *prop_slpa;
  int *n_states, *step;
  uint8_t *is_div, *is_turn;
  uint32_t* key_int;                                              // [C,2]
  int *list_a, *list_b;                                           // [C] compacted indices of expanding chains
  int* counters;                                                  // [depth+2]
  int max_depth;                                                  // checkpoint capacity
};

// nuts.py:133-136,278-294: key split, momentum draw, initial proposal / trajectory / termination state
template <class R, int TK, bool DM>
__device__ __forceinline__ void nuts_init_row(const Params& P, const NutsWs& ws, const uint32_t* keys, const float* q_in,
                                              const float* logp_in, const float* g_in, float* q_out, float* logp_out,
                                              float* g_out, const float* mom_override, const uint32_t* keyint_override,
                                              float* mom_out, int chain, int lane, float* sm) {
  const size_t roff = (size_t)chain * P.D;
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
    const Key rng = chain_key(P, keys, chain);
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
}

template <class R, int TK, bool DM>
__global__ void __launch_bounds__(kThreads) k_nuts_init(Params P, NutsWs ws, const uint32_t* __restrict__ keys,
                                                        const float* q_in, const float* logp_in, const float* g_in,
                                                        float* q_out, float* logp_out, float* g_out,
                                                        const float* mom_override, const uint32_t* keyint_override,
                                                        float* mom_out) {
  BJX_WARP_PROLOGUE();
  (void)roff;
  nuts_init_row<R, TK, DM>(P, ws, keys, q_in, logp_in, g_in, q_out, logp_out, g_out, mom_override, keyint_override, mom_out,
                           chain, lane, sm);
}

// is_turning(ckpt_p, p, p_sum - ckpt_sum + ckpt_p) against one checkpoint row (termination.py:96-103,
// metrics.py:272-304), streaming the checkpoint through registers a slot at a time.
template <class R, int TK, bool DM>
__device__ __forceinline__ bool turning_vs_checkpoint(Ctx<R, TK, DM>& c, const Params& P, const float* cp_row,
                                                      const float* cs_row, const float (&p)[R::NS],
                                                      const float (&ps)[R::NS], int lane) {
  // cp_row / cs_row are GENERIC addresses (shared memory or the global workspace): plain loads only
  if constexpr (DM) {
    float cp[R::NS], cs[R::NS];
    R::load_generic(cp, cp_row, P.D, lane);
    R::load_generic(cs, cs_row, P.D, lane);
#pragma unroll
    for (int s = 0; s < R::NS; ++s) cs[s] = ps[s] - cs[s] + cp[s];
    return c.is_turning(P, cp, p, cs);
  } else {
    float al = 0.f, ar = 0.f;
    if constexpr (R::VEC) {
#pragma unroll
      for (int j = 0; j < R::NS / 4; ++j) {
        const int e = (j * 32 + lane) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (e < P.D) {
          a = *reinterpret_cast<const float4*>(cp_row + e);
          b = *reinterpret_cast<const float4*>(cs_row + e);
        }
        const float cpv[4] = {a.x, a.y, a.z, a.w}, csv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int s = 4 * j + v;
          const float sub = ps[s] - csv[v] + cpv[v];
          const float rho = sub - (p[s] + cpv[v]) / 2.0f;
          al = fmaf(c.mw[s] * cpv[v], rho, al);
          ar = fmaf(c.mw[s] * p[s], rho, ar);
        }
      }
    } else {
#pragma unroll
      for (int s = 0; s < R::NS; ++s) {
        const int e = s * 32 + lane;
        const float cpv = (e < P.D) ? cp_row[e] : 0.f;
        const float csv = (e < P.D) ? cs_row[e] : 0.f;
        const float sub = ps[s] - csv + cpv;
        const float rho = sub - (p[s] + cpv) / 2.0f;
        al = fmaf(c.mw[s] * cpv, rho, al);
        ar = fmaf(c.mw[s] * p[s], rho, ar);
      }
    }
    al = warp_sum(al);
    ar = warp_sum(ar);
    return (al <= 0.f) || (ar <= 0.f);
  }
}

// Tree doublings [d_begin, d_end) for every chain still expanding (trajectory.py:642-717).  Per doubling: draw
// the direction, integrate the sub-tree of up to 2^d leaves (trajectory.py:318-372) with progressive uniform
// sampling (proposal.py:118-143) and the iterative U-turn checkpoints (termination.py:56-104), then update the
// proposal (biased progressive sampling, proposal.py:146-176), merge the trajectories and test the
// full-trajectory U-turn.  The host fuses the first few doublings (every chain runs them and their cost is the
// fixed per-doubling row traffic) into one launch and gives each deep doubling its own launch over the compacted
// list of chains that still expand; chains that want doubling d_end are appended to list_out.
// Checkpoints live in shared memory when depth x D is small enough (ckpt_smem), else in the global workspace.
// STRIDE: the row count is read on the device (n_in_dev) and a fixed grid strides over the list; the loop costs ~30
// registers, so the launch over all chains (row count known on the host) is instantiated without it.
// One chain (one warp) through doublings [d_begin, d_end); returns whether the chain wants doubling d_end.
template <class R, int TK, bool DM, bool GEN>
__device__ __forceinline__ bool nuts_doubling_row(const Params& P, const NutsWs& ws, int chain, int d_begin, int d_end,
                                                  int max_doublings, float* q_out, float* logp_out, float* g_out,
                                                  int ckpt_smem, int lane, int wib, float* sm, float* bjx_smem) {
  const size_t roff = (size_t)chain * P.D;
  Ctx<R, TK, DM> c;
  c.init(P, chain, lane, sm);
  const Key ki{ws.key_int[2 * chain], ws.key_int[2 * chain + 1]};
  const float eps_c = P.eps_dev ? P.eps_dev[chain] : P.eps;
  const float h0 = ws.h0[chain];
  const float ninf = -__int_as_float(0x7f800000);
  // checkpoint rows: [depth][D] momentum and [depth][D] momentum sums
  float* ck_p;
  float* ck_s;
  if (ckpt_smem) {
    // a sub-tree of 2^d leaves touches checkpoint rows 0..d-1, so this launch needs d_end rows per warp
    float* base = bjx_smem + (size_t)(needs_row_smem<TK, DM>() ? kWarpsPerBlock * P.D : 0) +
                  (size_t)wib * 2 * d_end * P.D;
    ck_p = base;
    ck_s = base + (size_t)d_end * P.D;
  } else {
    ck_p = ws.ckpt_p + (size_t)chain * ws.max_depth * P.D;
    ck_s = ws.ckpt_sum + (size_t)chain * ws.max_depth * P.D;
  }
  float prop_weight = ws.prop_weight[chain], prop_slpa = ws.prop_slpa[chain];
  int n_states = ws.n_states[chain];
  bool run_next = true;
  // Rows of up to 8 slots per lane: BOTH trajectory endpoints and the trajectory momentum sum stay in registers
  // across the doublings of this launch (the "moving" endpoint is the one the current doubling extends, the "fixed"
  // one only contributes its momentum to the U-turn test); larger rows re-load the moving endpoint per doubling.
  constexpr bool RES = (R::NS <= 8);
  constexpr int NF = RES ? R::NS : 1;
  float q[R::NS], p[R::NS], g[R::NS], ps[R::NS];
  float fq[NF], fp[NF], fg[NF], tsum[NF];
  bool mov_right = true;
  if constexpr (RES) {
    R::load(q, ws.right_q + roff, P.D, lane);
    R::load(p, ws.right_p + roff, P.D, lane);
    R::load(g, ws.right_g + roff, P.D, lane);
    R::load(fq, ws.left_q + roff, P.D, lane);
    R::load(fp, ws.left_p + roff, P.D, lane);
    R::load(fg, ws.left_g + roff, P.D, lane);
    R::load(tsum, ws.psum + roff, P.D, lane);
  }
  float logp_mov = 0.f, logp_fix = 0.f;
  if (RES) {
    logp_mov = ws.right_logp[chain];
    logp_fix = ws.left_logp[chain];
  }
  // ---- key schedule of the doublings of this launch (trajectory.py:645-650), lane-parallel ------------------
  // Doubling d needs subkey = fold_in(rng_key, d), then (direction_key, trajectory_key, proposal_key) =
  // split(subkey, 3) and two scalar uniforms.  All of it is warp-uniform integer work, so lane j evaluates the
  // subkey of doubling d_begin + j, lane 3j + r the r-th split child of doubling j, and the same lanes their
  // uniform: three threefry passes per LAUNCH instead of six blocks per doubling (the host keeps
  // d_end - d_begin <= 10, i.e. at most 30 busy lanes).
  const int nd = d_end - d_begin;
  const Key sub_l = fold_in(ki, (uint32_t)(d_begin + (lane < nd ? lane : 0)));
  const int kj = (lane < 3 * nd) ? lane / 3 : 0, kr = lane % 3;
  const Key sub_j{__shfl_sync(0xffffffffu, sub_l.a, kj), __shfl_sync(0xffffffffu, sub_l.b, kj)};
  const Key k3 = fold_in(sub_j, (uint32_t)kr);   // r = 0 direction key, 1 trajectory key, 2 proposal key
  const float u3 = uniform01(k3);                 // meaningful in the direction / proposal lanes
  int d = d_begin;
  for (; d < d_end && run_next; ++d) {
    // ---- begin: direction and keys of this doubling (trajectory.py:645-655) ------------------------------
    const int kl = 3 * (d - d_begin);
    const Key tk{__shfl_sync(0xffffffffu, k3.a, kl + 1), __shfl_sync(0xffffffffu, k3.b, kl + 1)};
    const float u_prop = __shfl_sync(0xffffffffu, u3, kl + 2);
    const int dir = (__shfl_sync(0xffffffffu, u3, kl) < 0.5f) ? 1 : -1;
    float* eq = dir > 0 ? ws.right_q : ws.left_q;
    float* ep = dir > 0 ? ws.right_p : ws.left_p;
    float* eg = dir > 0 ? ws.right_g : ws.left_g;
    if constexpr (RES) {
      if ((dir > 0) != mov_right) {  // the other endpoint becomes the moving one: swap roles
#pragma unroll
        for (int s = 0; s < R::NS; ++s) {
          float t0 = q[s]; q[s] = fq[s]; fq[s] = t0;
          float t1 = p[s]; p[s] = fp[s]; fp[s] = t1;
          float t2 = g[s]; g[s] = fg[s]; fg[s] = t2;
        }
        const float tl = logp_mov; logp_mov = logp_fix; logp_fix = tl;
        mov_right = !mov_right;
      }
    } else {
      R::load(q, eq + roff, P.D, lane);
      R::load(p, ep + roff, P.D, lane);
      R::load(g, eg + roff, P.D, lane);
    }
    const float eps = (float)dir * eps_c;  // direction * step_size  :323

    // ---- the sub-tree (trajectory.py:318-372) ----------------------------------------------------------------
    float sub_weight = ninf, sub_slpa = ninf, sub_logp = 0.f, sub_energy = 0.f, logp = 0.f;
    bool sub_div = false, sub_term = false;
    int n = 0;
    const int n_leaves = 1 << d;
    float u_batch = 0.f;
    for (int i = 0; i < n_leaves; ++i) {
      // The multinomial draw of leaf i is uniform(fold_in(rng_key, i)) (trajectory.py:321): two threefry blocks of
      // warp-uniform integer work.  Lane l evaluates the draw of leaf i+l instead, so one pass serves 32 leaves.
      if ((i & 31) == 0) u_batch = uniform01(fold_in(tk, (uint32_t)(i + lane)));
      const float u_leaf = __shfl_sync(0xffffffffu, u_batch, i & 31);
      c.template step<GEN, true>(P, q, p, g, logp, eps);
      const float e_new = -logp + c.kinetic(P, p);
      const float w_new = safe_energy_diff(h0, e_new);  // proposal.py:94-98
      const float slpa_new = fminf(w_new, 0.f);
      const bool is_div = (-w_new) > P.div_thr;         // :325
      bool take;
      if (i == 0) {  // :329-334 the first leaf is taken unconditionally
#pragma unroll
        for (int s = 0; s < R::NS; ++s) ps[s] = p[s];
        take = true;
        sub_weight = w_new;
        sub_slpa = slpa_new;
      } else {  // :335-338 append + progressive uniform sampling
#pragma unroll
        for (int s = 0; s < R::NS; ++s) ps[s] = ps[s] + p[s];
        const float p_accept = expit_f(w_new - sub_weight);
        take = u_leaf < p_accept;
        // the two logaddexp updates (proposal.py:124-127) in one SIMD evaluation: lane 0 weight, other lanes slpa
        const float la = logaddexp_f(lane == 0 ? sub_weight : sub_slpa, lane == 0 ? w_new : slpa_new);
        sub_weight = __shfl_sync(0xffffffffu, la, 0);
        sub_slpa = __shfl_sync(0xffffffffu, la, 1);
      }
      if (take) {
        R::store(q, ws.sub_prop_q + roff, P.D, lane);
        R::store(g, ws.sub_prop_g + roff, P.D, lane);
        sub_logp = logp;
        sub_energy = e_new;
      }
      n = i + 1;
      // termination.py:75-84 checkpoint index range of leaf i
      const int idx_max = __popc((unsigned)i >> 1);
      const int idx_min = idx_max - __popc((~(unsigned)i & ((unsigned)i + 1u)) - 1u) + 1;
      if ((i & 1) == 0) {  // termination.py:66-72
        R::store_generic(p, ck_p + (size_t)idx_max * P.D, P.D, lane);
        R::store_generic(ps, ck_s + (size_t)idx_max * P.D, P.D, lane);
      }
      bool turning = false;
      for (int k = idx_max; k >= idx_min && !turning; --k)  // termination.py:96-103
        turning = turning_vs_checkpoint<R, TK, DM>(c, P, ck_p + (size_t)k * P.D, ck_s + (size_t)k * P.D, p, ps, lane);
      sub_div = is_div;
      sub_term = turning;
      if (is_div || turning) break;
    }
    // the last leaf is the new endpoint of the merged trajectory (trajectory.py:376-385,697-704)
    if constexpr (RES) {
      logp_mov = logp;
    } else {
      R::store(q, eq + roff, P.D, lane);
      R::store(p, ep + roff, P.D, lane);
      R::store(g, eg + roff, P.D, lane);
    }

    // ---- end of the doubling (trajectory.py:672-717) -----------------------------------------------------------
    const bool bad = sub_div || sub_term;
    bool take2 = false;
    if (!bad) take2 = u_prop < clip_max1(expf(sub_weight - prop_weight));  // proposal.py:155-156
    if (take2) {
      float t[R::NS];
      R::load(t, ws.sub_prop_q + roff, P.D, lane);
      R::store(t, q_out + roff, P.D, lane);
      R::load(t, ws.sub_prop_g + roff, P.D, lane);
      R::store(t, g_out + roff, P.D, lane);
    }
    bool turning;
    if constexpr (RES) {
#pragma unroll
      for (int s = 0; s < R::NS; ++s) tsum[s] = tsum[s] + ps[s];  // merge_trajectories  trajectory.py:102-125
      // :706-710 is_turning(p_left, p_right, p_sum)
      turning = (dir > 0) ? c.is_turning(P, fp, p, tsum) : c.is_turning(P, p, fp, tsum);
    } else {
      float t[R::NS];
      R::load(t, ws.psum + roff, P.D, lane);
#pragma unroll
      for (int s = 0; s < R::NS; ++s) ps[s] = t[s] + ps[s];
      R::store(ps, ws.psum + roff, P.D, lane);
      R::load(t, (dir > 0 ? ws.left_p : ws.right_p) + roff, P.D, lane);  // the endpoint that did not move
      turning = (dir > 0) ? c.is_turning(P, t, p, ps) : c.is_turning(P, p, t, ps);
    }
    const bool is_turn = sub_term || turning;  // :715
    run_next = (d + 1 < max_doublings) && !sub_div && !is_turn;
    if (!bad) prop_weight = logaddexp_f(prop_weight, sub_weight);
    prop_slpa = logaddexp_f(prop_slpa, sub_slpa);
    n_states += n;
    if (lane == 0) {
      if (!RES) {
        if (dir > 0) ws.right_logp[chain] = logp; else ws.left_logp[chain] = logp;
      }
      if (take2) {
        logp_out[chain] = sub_logp;
        ws.prop_energy[chain] = sub_energy;
      }
      ws.is_div[chain] = sub_div;
      ws.is_turn[chain] = is_turn;
    }
  }
  if constexpr (RES) {  // write the trajectory back once
    R::store(q, (mov_right ? ws.right_q : ws.left_q) + roff, P.D, lane);
    R::store(p, (mov_right ? ws.right_p : ws.left_p) + roff, P.D, lane);
    R::store(g, (mov_right ? ws.right_g : ws.left_g) + roff, P.D, lane);
    R::store(fq, (mov_right ? ws.left_q : ws.right_q) + roff, P.D, lane);
    R::store(fp, (mov_right ? ws.left_p : ws.right_p) + roff, P.D, lane);
    R::store(fg, (mov_right ? ws.left_g : ws.right_g) + roff, P.D, lane);
    R::store(tsum, ws.psum + roff, P.D, lane);
  }
  if (lane == 0) {
    if (RES) {
      ws.right_logp[chain] = mov_right ? logp_mov : logp_fix;
      ws.left_logp[chain] = mov_right ? logp_fix : logp_mov;
    }
    ws.prop_weight[chain] = prop_weight;
    ws.prop_slpa[chain] = prop_slpa;
    ws.n_states[chain] = n_states;
    ws.step[chain] = d;
  }
  return run_next && d < max_doublings;
}

template <class R, int TK, bool DM, bool GEN, bool STRIDE>
__global__ void __launch_bounds__(kThreads) k_nuts_doubling(Params P, NutsWs ws, int d_begin, int d_end, int max_doublings,
                                                            const int* __restrict__ list_in, int n_in,
                                                            const int* __restrict__ n_in_dev,
                                                            int* __restrict__ list_out, int* counter_out,
                                                            float* q_out, float* logp_out, float* g_out, int ckpt_smem) {
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  extern __shared__ __align__(16) float bjx_smem[];
  float* sm = bjx_smem + (size_t)wib * P.D;  // small dense matvec slice (first kWarpsPerBlock*D floats when used)
  // n_in_dev: the row count was produced on the device by the previous launch (no host round trip); the grid is then
  // a fixed number of CTAs whose warps stride over the compacted list (surplus warps leave at once)
  const int n_rows = (STRIDE && n_in_dev) ? *n_in_dev : n_in;
  for (int w = blockIdx.x * kWarpsPerBlock + wib; w < n_rows; w += STRIDE ? gridDim.x * kWarpsPerBlock : n_rows) {
    const int chain = list_in ? list_in[w] : w;
    const bool more = nuts_doubling_row<R, TK, DM, GEN>(P, ws, chain, d_begin, d_end, max_doublings, q_out, logp_out, g_out,
                                                        ckpt_smem, lane, wib, sm, bjx_smem);
    if (lane == 0 && more) list_out[atomicAdd(counter_out, 1)] = chain;
    if constexpr (!STRIDE) return;  // one row per warp: no loop for the compiler to carry state across
    __syncwarp();
  }  // next row of the list
}

// ---- run_inference_algorithm for NUTS with the chains decoupled (bjx_nuts_sample) --------------------------------
// Chains never interact and transition t of chain c only needs step key t, so nothing forces the chains through the
// transitions in lock step: a persistent grid whose warps pull chains from a queue and run ALL num_steps transitions of a
// chain back to back (init row -> doublings 0..max -> info), in place.  The deep trees that leave a per-transition
// launch with 10 % of its warp slots busy (profiles/r02_ncu_nuts.md) are then hidden behind the other chains' work.
// Results are bit-identical to num_steps calls of bjx_nuts_step (same device functions, same keys).
struct NutsSampleArgs {
  const uint32_t* step_keys;  // [num_steps, 2]
  int num_steps;
  int max_doublings;
  int ckpt_smem;
  float* history;             // [num_steps / thin, C, D] or nullptr
  int thin;
  float* acceptance_history;  // [num_steps, C] or nullptr
  int* nint_history;          // [num_steps, C] or nullptr
  unsigned long long* leapfrogs;  // += sum of num_integration_steps, or nullptr
  int* queue;                 // zeroed by the host
};

template <class R, int TK, bool DM, bool GEN>
__global__ void __launch_bounds__(kThreads, (R::NS <= 4 && !DM) ? 5 : 1) k_nuts_chains(Params P, NutsWs ws, float* q_io, float* logp_io, float* g_io,
                                                          NutsSampleArgs A) {
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  extern __shared__ __align__(16) float bjx_smem[];
  float* sm = bjx_smem + (size_t)wib * P.D;
  for (;;) {
    int chain = 0;
    if (lane == 0) chain = atomicAdd(A.queue, 1);
    chain = __shfl_sync(0xffffffffu, chain, 0);
    if (chain >= P.C) return;
    const size_t roff = (size_t)chain * P.D;
    unsigned long long n_leapfrogs = 0;
    int dmax = 0;
    for (int t = 0; t < A.num_steps; ++t) {
      nuts_init_row<R, TK, DM>(P, ws, A.step_keys + 2 * t, q_io, logp_io, g_io, q_io, logp_io, g_io, nullptr, nullptr, nullptr,
                               chain, lane, sm);
      __syncwarp();
      for (int d0 = 0; d0 < A.max_doublings; d0 += 10) {  // the lane-parallel key schedule covers 10 doublings per call
        const int d1 = min(A.max_doublings, d0 + 10);
        const bool more = nuts_doubling_row<R, TK, DM, GEN>(P, ws, chain, d0, d1, A.max_doublings, q_io, logp_io, g_io,
                                                            A.ckpt_smem, lane, wib, sm, bjx_smem);
        __syncwarp();
        if (!more) break;
      }
      const int n = ws.n_states[chain];
      n_leapfrogs += (unsigned long long)n;
      dmax = max(dmax, ws.step[chain]);
      if (lane == 0) {
        if (A.acceptance_history) A.acceptance_history[(size_t)t * P.C + chain] = expf(ws.prop_slpa[chain]) / (float)n;
        if (A.nint_history) A.nint_history[(size_t)t * P.C + chain] = n;
      }
      if (A.history && ((t + 1) % A.thin) == 0) {
        float q[R::NS];
        R::load(q, q_io + roff, P.D, lane);
        R::store(q, A.history + ((size_t)((t + 1) / A.thin - 1) * P.C + chain) * P.D, P.D, lane);
      }
      __syncwarp();
    }
    if (lane == 0) {
      if (A.leapfrogs) atomicAdd(A.leapfrogs, n_leapfrogs);
      atomicMax(ws.counters + 63, dmax);
    }
  }
}

// nuts.py:303-319: acceptance_rate = exp(sum_log_p_accept) / num_states and the NUTSInfo scalars
static __global__ void k_nuts_finish(int C, NutsWs ws, InfoPtrs info) {
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  int dmax = 0;
  if (chain < C) {
    const int n = ws.n_states[chain];
    if (info.acceptance_rate) info.acceptance_rate[chain] = expf(ws.prop_slpa[chain]) / (float)n;
    if (info.is_divergent) info.is_divergent[chain] = ws.is_div[chain];
    if (info.is_turning) info.is_turning[chain] = ws.is_turn[chain];
    if (info.energy) info.energy[chain] = ws.prop_energy[chain];
    if (info.num_integration_steps) info.num_integration_steps[chain] = n;
    if (info.num_trajectory_expansions) info.num_trajectory_expansions[chain] = ws.step[chain];
    dmax = ws.step[chain];
  }
  // deepest tree of this transition (bjx_nuts_last_stats): one atomic per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dmax = max(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  if ((threadIdx.x & 31) == 0) atomicMax(ws.counters + 63, dmax);
}

}  // namespace bjx
