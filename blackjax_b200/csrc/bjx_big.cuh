// Kernels for rows beyond a warp (1024 < D <= 18432, D % 4 == 0): one CTA owns one chain (see bjx_big.cu for the design
// notes).  Header so that the plug-in of a user-defined target (bjx_plugin.cu with -DBJX_PLUGIN_BIG=1) instantiates the same
// kernels around the user's CTA-level value_and_grad (bjx_user::BigModel, include/bjx_user_target.h) as libbjx does around
// the built-in targets.
#pragma once
#include "../../include/bjx.h"
#include "bjx_kernels.cuh"

namespace bjx_user {
struct BigModel;  // defined by a big-row plug-in's source
}

namespace bjx {

constexpr int kBigThreads = 768;  // one CTA per SM (shared-memory bound): 24 warps hide the MUFU / L2 latency
constexpr int kBigWarps = kBigThreads / 32;

struct BigParams {
  int C, D;
  int kind;
  const float* inv_var;
  const float* mean;
  float logp_offset;
  const float* data_x;
  const uint8_t* data_y;
  int G;
  const float* user;   // BJX_TARGET_USER: the model's parameter block
  int n_user;
  const float* imm;
  long long imm_stride;
  const float* msqrt;
  float eps;
  const float* eps_dev;
  float div_thr;
  int key_shared;
  uint32_t chain_offset;
};

// block-wide sum of up to NV values per thread (result valid in all threads)
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red /*[NV*kBigWarps]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  __syncthreads();  // protect red from the previous use
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) red[k * kBigWarps + wid] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kBigWarps; ++w) s += red[k * kBigWarps + w];
    v[k] = s;
  }
}

// value_and_grad of the target at the row q (shared memory) -> g (shared memory), returns logp (all threads).
// WANT_LOGP = false: only the gradient is consumed (interior steps of a fixed-length trajectory, where the value is
// dead -- as in the warp kernels); the hierarchical logistic model then skips the softplus (one MUFU of three and a
// quarter of the instructions per observation).
// CTA-level context of a user-defined big-row model (include/bjx_user_target.h)
struct BigUserCtx {
  const float* theta;  // parameter block [n_theta] (device memory, read-only)
  int n_theta;
  int D;
  int tid;             // threadIdx.x in [0, kBigThreads)
};
template <int TK>
struct BigUserOf { using type = void; };
template <>
struct BigUserOf<BJX_TARGET_USER> { using type = bjx_user::BigModel; };

template <int TK, bool WANT_LOGP = true>
__device__ __forceinline__ float big_value_and_grad(const BigParams& P, const float* q, float* g, float* red) {
  const int D = P.D, tid = threadIdx.x;
  if constexpr (TK == BJX_TARGET_USER) {
    const BigUserCtx u{P.user, P.n_user, D, tid};
    using M = typename BigUserOf<TK>::type;
    return M::template value_and_grad<WANT_LOGP>(u, q, g, red) + P.logp_offset;
  } else if constexpr (TK == BJX_TARGET_DIAG_GAUSSIAN) {
    float acc[1] = {0.f};
    for (int i = tid; i < D; i += kBigThreads) {
      const float d = P.mean ? q[i] - __ldg(P.mean + i) : q[i];
      const float t = d * -__ldg(P.inv_var + i);
      acc[0] = fmaf(d, t, acc[0]);
      g[i] = t;
    }
    if constexpr (!WANT_LOGP) return 0.f;
    block_sum<1>(acc, red);
    return 0.5f * acc[0] + P.logp_offset;
  } else if constexpr (TK == BJX_TARGET_FUNNEL) {
    const float y = q[0];
    float acc[1] = {0.f};
    for (int i = tid; i < D; i += kBigThreads) acc[0] = (i == 0) ? acc[0] : fmaf(q[i], q[i], acc[0]);
    block_sum<1>(acc, red);
    const float ss = acc[0], ey = expf(-y), n = (float)(D - 1), t = y / 3.0f;
    for (int i = tid; i < D; i += kBigThreads) g[i] = (i == 0) ? (-y / 9.0f + 0.5f * ey * ss - 0.5f * n) : -(ey * q[i]);
    return -0.5f * (t * t) + (-0.5f * ey * ss - 0.5f * n * y) + P.logp_offset;
  } else {  // BJX_TARGET_HIER_LOGIT
    const float mu = q[0], lt = q[1], b0 = q[2], b1 = q[3];
    const float e2 = expf(-2.0f * lt);
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // ll, sum d, sum d^2, grad b0, grad b1
    // Software pipeline over this thread's groups, two register sets (A, B): the 64 bytes of covariates + the outcome byte
    // of the NEXT group are requested before the 8 observations of the current one are evaluated (the data is
    // L2-resident, ~1 us away; ncu before: half of the stall samples on the first FMA that consumes a covariate).
    auto load_group = [&](int gidx, float4 (&xv)[4], unsigned& bv, float& av) {
      if (gidx < P.G) {
        const float4* xr = reinterpret_cast<const float4*>(P.data_x + (size_t)gidx * 16);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) xv[k2] = __ldg(xr + k2);
        bv = __ldg(P.data_y + gidx);
        av = q[4 + gidx];
      }
    };
    auto eval_group = [&](int gidx, const float4 (&xv4)[4], unsigned bits, float alpha) {
      const float d = alpha - mu;
      float ga = 0.f;
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        const float4 xv = xv4[k2];  // (x_{2k2,0}, x_{2k2,1}, x_{2k2+1,0}, x_{2k2+1,1})
        const float xs[2][2] = {{xv.x, xv.y}, {xv.z, xv.w}};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const bool yb = (bits >> (2 * k2 + u)) & 1u;
          const float eta = alpha + b0 * xs[u][0] + b1 * xs[u][1];
          // One exponential feeds both the sigmoid and the softplus, and the softplus reuses the sigmoid's reciprocal:
          //   ex = exp(-|eta|), r = 1 / (1 + ex), sigmoid = eta >= 0 ? r : ex r, softplus = max(eta, 0) - log(r).
          // ex2.approx / rcp.approx / lg2.approx: 3 MUFU + ~20 FP32/ALU instructions per observation (expf + an IEEE
          // division + log1pf + int-to-float conversions cost 70; __expf / __frcp_rn / (float)bit still 41, of which a
          // dozen guard subnormal ranges these arguments never reach).  Error bounds, checked against float64 in
          // tests/test_gpu_round2.py: |sigmoid error| <= 4e-7 (ex2 2 ulp, rcp 1 ulp), |softplus error| <= 3e-7 absolute.
          float ex, rc;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fabsf(eta) * -1.4426950408889634f));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(1.0f + ex));   // 1 + ex in [1, 2]
          const float sig = (eta >= 0.f) ? rc : ex * rc;
          const float r = (yb ? 1.0f : 0.0f) - sig;
          if constexpr (WANT_LOGP) {
            float l2;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(rc));        // rc in [1/2, 1]
            const float softplus = fmaf(-0.69314718f, l2, fmaxf(eta, 0.f));
            acc[0] += (yb ? eta : 0.0f) - softplus;
          }
          ga += r;
          acc[3] = fmaf(r, xs[u][0], acc[3]);
          acc[4] = fmaf(r, xs[u][1], acc[4]);
        }
      }
      acc[1] += d;
      acc[2] = fmaf(d, d, acc[2]);
      g[4 + gidx] = -d * e2 + ga;
    };
    float4 xa[4], xb[4];
    unsigned ba = 0, bb = 0;
    float aa = 0.f, ab = 0.f;
    load_group(tid, xa, ba, aa);
    for (int gi = tid; gi < P.G; gi += 2 * kBigThreads) {
      load_group(gi + kBigThreads, xb, bb, ab);
      eval_group(gi, xa, ba, aa);
      load_group(gi + 2 * kBigThreads, xa, ba, aa);
      if (gi + kBigThreads < P.G) eval_group(gi + kBigThreads, xb, bb, ab);
    }
    block_sum<5>(acc, red);
    if (tid == 0) {
      g[0] = -0.01f * mu + e2 * acc[1];
      g[1] = -lt + e2 * acc[2] - (float)P.G;
      g[2] = -0.16f * b0 + acc[3];
      g[3] = -0.16f * b1 + acc[4];
    }
    return -0.005f * mu * mu - 0.5f * lt * lt - 0.08f * (b0 * b0 + b1 * b1) + (-0.5f * e2 * acc[2] - (float)P.G * lt) +
           acc[0] + P.logp_offset;
  }
}

__device__ __forceinline__ float big_kinetic(const BigParams& P, const float* imm, const float* p, float* red) {
  float acc[1] = {0.f};
  for (int i = threadIdx.x; i < P.D; i += kBigThreads) acc[0] = fmaf(__ldg(imm + i) * p[i], p[i], acc[0]);
  block_sum<1>(acc, red);
  return 0.5f * acc[0];
}

// n velocity-Verlet steps (integrators.py:104-150) on the row in shared memory; returns the log-density after the last
// one.  Between two steps the closing half kick of one and the opening half kick of the next run in the same pass over
// the row (two separately rounded FMAs, bit-identical to the step-by-step form), so an interior step is one pass +
// one gradient evaluation (gradient only: its log-density is dead) instead of two passes.
template <int TK>
__device__ __forceinline__ float big_trajectory(const BigParams& P, const float* imm, float* q, float* p, float* g,
                                                float eps, int n, float* red) {
  const float eh = eps * 0.5f, e1 = eps * 1.0f;
  if (n <= 0) return 0.f;
  for (int i = threadIdx.x; i < P.D; i += kBigThreads) {
    const float pn = fmaf(eh, g[i], p[i]);
    p[i] = pn;
    q[i] = fmaf(e1, __ldg(imm + i) * pn, q[i]);
  }
  __syncthreads();
  for (int s = 0; s + 1 < n; ++s) {
    big_value_and_grad<TK, false>(P, q, g, red);
    __syncthreads();
    for (int i = threadIdx.x; i < P.D; i += kBigThreads) {
      float pn = fmaf(eh, g[i], p[i]);   // second half kick of step s
      pn = fmaf(eh, g[i], pn);           // first half kick of step s + 1
      p[i] = pn;
      q[i] = fmaf(e1, __ldg(imm + i) * pn, q[i]);
    }
    __syncthreads();
  }
  const float logp = big_value_and_grad<TK, true>(P, q, g, red);
  __syncthreads();
  for (int i = threadIdx.x; i < P.D; i += kBigThreads) p[i] = fmaf(eh, g[i], p[i]);
  __syncthreads();
  return logp;
}

__device__ __forceinline__ void big_load(float* dst, const float* __restrict__ src, int D) {
  for (int i = threadIdx.x; i < D / 4; i += kBigThreads)
    reinterpret_cast<float4*>(dst)[i] = __ldcs(reinterpret_cast<const float4*>(src) + i);
}
__device__ __forceinline__ void big_store(float* __restrict__ dst, const float* src, int D) {
  for (int i = threadIdx.x; i < D / 4; i += kBigThreads)
    __stcs(reinterpret_cast<float4*>(dst) + i, reinterpret_cast<const float4*>(src)[i]);
}

template <int TK>
__global__ void __launch_bounds__(kBigThreads) k_big_init(BigParams P, const float* __restrict__ q_in,
                                                          float* __restrict__ logp_out, float* __restrict__ g_out) {
  extern __shared__ __align__(16) float sm[];
  float *q = sm, *g = sm + P.D, *red = sm + 2 * (size_t)P.D;
  const size_t ro = (size_t)blockIdx.x * P.D;
  big_load(q, q_in + ro, P.D);
  __syncthreads();
  const float logp = big_value_and_grad<TK>(P, q, g, red);
  __syncthreads();
  big_store(g_out + ro, g, P.D);
  if (threadIdx.x == 0) logp_out[blockIdx.x] = logp;
}

// n leapfrog steps per launch (bjx_leapfrog; n = 1 is the one-step kernel the HBM roofline is quoted on).  Thread t owns
// elements t + 768 k of the row, so the momentum is thread-private and stays in REGISTERS (NK per thread); only q and grad,
// which the target's value_and_grad addresses freely, live in shared memory: 2 x 4 D bytes instead of 3 x 4 D, and for
// rows up to 10752 dims two CTAs fit an SM, so one CTA's loads and stores overlap the other's arithmetic (with three rows
// and one CTA per SM the phases load -> compute -> store ran back to back: 0.45 of the HBM peak at 32768 x 10000).
template <int TK, int NK>
__global__ void __launch_bounds__(kBigThreads, (NK <= 14 && (TK == BJX_TARGET_DIAG_GAUSSIAN || TK == BJX_TARGET_FUNNEL)) ? 2 : 1)
    k_big_leapfrog(BigParams P, float* q_io, float* p_io, float* logp_io, float* g_io, int n_steps) {
  extern __shared__ __align__(16) float sm[];
  float *q = sm, *g = sm + P.D, *red = sm + 2 * (size_t)P.D;
  const int c = blockIdx.x, D = P.D, tid = threadIdx.x;
  const size_t ro = (size_t)c * D;
  const float* imm = P.imm + (size_t)c * P.imm_stride;
  const float eps = P.eps_dev ? P.eps_dev[c] : P.eps;
  const float eh = eps * 0.5f, e1 = eps * 1.0f;
  float p[NK];
  if (n_steps <= 0) return;
  // load + first half kick + position update in one pass (integrators.py:199-203,235-245)
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int i = tid + kBigThreads * k;
    p[k] = 0.f;
    if (i < D) {
      const float pn = fmaf(eh, __ldcs(g_io + ro + i), __ldcs(p_io + ro + i));
      p[k] = pn;
      q[i] = fmaf(e1, __ldg(imm + i) * pn, __ldcs(q_io + ro + i));
    }
  }
  __syncthreads();
  for (int s = 0; s + 1 < n_steps; ++s) {
    big_value_and_grad<TK, false>(P, q, g, red);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int i = tid + kBigThreads * k;
      if (i < D) {
        const float gv = g[i];
        float pn = fmaf(eh, gv, p[k]);   // second half kick of step s
        pn = fmaf(eh, gv, pn);           // first half kick of step s + 1
        p[k] = pn;
        q[i] = fmaf(e1, __ldg(imm + i) * pn, q[i]);
      }
    }
    __syncthreads();
  }
  const float logp = big_value_and_grad<TK, true>(P, q, g, red);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int i = tid + kBigThreads * k;
    if (i < D) {
      const float gv = g[i];
      __stcs(p_io + ro + i, fmaf(eh, gv, p[k]));
      __stcs(g_io + ro + i, gv);
      __stcs(q_io + ro + i, q[i]);
    }
  }
  if (tid == 0) logp_io[c] = logp;
}

// whole HMC transition (hmc.py:279-312) with the row resident in shared memory
template <int TK>
__global__ void __launch_bounds__(kBigThreads) k_big_hmc(BigParams P, const uint32_t* __restrict__ keys, const float* q_in,
                                                         const float* logp_in, const float* g_in, float* q_out,
                                                         float* logp_out, float* g_out, int L, InfoPtrs info) {
  extern __shared__ __align__(16) float sm[];
  float *q = sm, *p = sm + P.D, *g = sm + 2 * (size_t)P.D, *red = sm + 3 * (size_t)P.D;
  const int c = blockIdx.x;
  const size_t ro = (size_t)c * P.D;
  big_load(q, q_in + ro, P.D);
  big_load(g, g_in + ro, P.D);
  const Key rng = P.key_shared ? fold_in(Key{keys[0], keys[1]}, P.chain_offset + (uint32_t)c) : Key{keys[2 * c], keys[2 * c + 1]};
  const Key km = fold_in(rng, 0u), ki = fold_in(rng, 1u);  // hmc.py:299
  const float* imm = P.imm + (size_t)c * P.imm_stride;
  const float* ms = P.msqrt + (size_t)c * P.imm_stride;
  for (int i = threadIdx.x; i < P.D; i += kBigThreads) {
    const float pv = __ldg(ms + i) * normal_at(km, (uint32_t)i);  // hmc.py:302
    p[i] = pv;
    if (info.momentum) info.momentum[ro + i] = pv;
  }
  __syncthreads();
  const float logp0 = logp_in[c];
  const float e0 = -logp0 + big_kinetic(P, imm, p, red);  // hmc.py:159
  const float eps = P.eps_dev ? P.eps_dev[c] : P.eps;
  const float logp = L > 0 ? big_trajectory<TK>(P, imm, q, p, g, eps, L, red) : logp0;  // trajectory.py:165
  const float e1 = -logp + big_kinetic(P, imm, p, red);  // hmc.py:160 (kinetic energy is even in p: the flip is implicit)
  float delta = e0 - e1;
  if (isnan(delta)) delta = -__int_as_float(0x7f800000);  // proposal.py:45-48
  const bool is_div = (-delta) > P.div_thr;
  float pa = expf(delta);
  pa = pa > 1.0f ? 1.0f : pa;                             // proposal.py:225
  const bool acc = uniform01(ki) < pa;                    // proposal.py:226
  if (info.proposal_position) big_store(info.proposal_position + ro, q, P.D);
  if (info.proposal_momentum)
    for (int i = threadIdx.x; i < P.D; i += kBigThreads) info.proposal_momentum[ro + i] = -1.0f * p[i];  // hmc.py:158
  if (acc) {
    big_store(q_out + ro, q, P.D);
    big_store(g_out + ro, g, P.D);
  } else if (q_out != q_in) {
    for (int i = threadIdx.x; i < P.D / 4; i += kBigThreads) {
      __stcs(reinterpret_cast<float4*>(q_out + ro) + i, __ldcs(reinterpret_cast<const float4*>(q_in + ro) + i));
      __stcs(reinterpret_cast<float4*>(g_out + ro) + i, __ldcs(reinterpret_cast<const float4*>(g_in + ro) + i));
    }
  }
  if (threadIdx.x == 0) {
    if (acc || q_out != q_in) logp_out[c] = acc ? logp : logp0;
    if (info.acceptance_rate) info.acceptance_rate[c] = pa;
    if (info.is_accepted) info.is_accepted[c] = acc;
    if (info.is_divergent) info.is_divergent[c] = is_div;
    if (info.energy) info.energy[c] = e1;
    if (info.num_integration_steps) info.num_integration_steps[c] = L;
  }
}

// ---- launch table shared by libbjx (built-in targets) and big-row plug-ins (TK = BJX_TARGET_USER) -------------------------
enum BigKernelId { BK_INIT = 0, BK_LEAPFROG = 1, BK_HMC = 2 };
struct BigLaunchArgs {
  BigParams P;
  const uint32_t* keys;
  const float *q_in, *logp_in, *g_in;
  float *q_out, *logp_out, *g_out;   // BK_INIT: logp_out, g_out; BK_LEAPFROG: q_out, logp_out, g_out in place (+ p_io)
  float* p_io;
  int n;                             // leapfrog steps | L
  InfoPtrs info;
  cudaStream_t stream;
};

// returns 0 or the cudaError_t of the attribute call / launch
template <int TK>
static int big_launch(int kernel_id, const BigLaunchArgs& a) {
  const int D = a.P.D, C = a.P.C;
  cudaStream_t st = a.stream;
  auto go = [&](auto kern, int rows, auto... args) -> int {
    const size_t smem = ((size_t)rows * D + 5 * kBigWarps) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<C, kBigThreads, smem, st>>>(args...);
    return (int)cudaGetLastError();
  };
  switch (kernel_id) {
    case BK_INIT:
      return go(k_big_init<TK>, 2, a.P, a.q_in, a.logp_out, a.g_out);
    case BK_LEAPFROG:
      if (D <= 14 * kBigThreads) return go(k_big_leapfrog<TK, 14>, 2, a.P, a.q_out, a.p_io, a.logp_out, a.g_out, a.n);
      return go(k_big_leapfrog<TK, 24>, 2, a.P, a.q_out, a.p_io, a.logp_out, a.g_out, a.n);
    case BK_HMC:
      return go(k_big_hmc<TK>, 3, a.P, a.keys, a.q_in, a.logp_in, a.g_in, a.q_out, a.logp_out, a.g_out, a.n, a.info);
    default:
      return (int)cudaErrorInvalidValue;
  }
}
}  // namespace bjx
