/*
 * bjx_user_target.h -- contract of a user-defined target (BJX_TARGET_USER in bjx.h).
 *
 * BlackJAX accepts any JAX callable as `logdensity_fn` and obtains its gradient with
 * `jax.value_and_grad(logdensity_fn)` (blackjax/mcmc/hmc.py:91, nuts.py:133, integrators.py:189,204).  There is no
 * tracing compiler on this side of the boundary, so the plug-in point is the fused value_and_grad itself: ONE small CUDA
 * device struct, written against the row layout of the kernels, compiled by nvcc together with
 * blackjax_b200/csrc/bjx_plugin.cu into a small shared library that holds every transition kernel of the path
 * (HMC, multinomial / generalized HMC, NUTS, the leapfrog and init kernels) instantiated around it.  The warm-up schemes,
 * samplers, metrics (diagonal, per-chain diagonal, dense and low-rank up to their row limits) and integrators then work with
 * the new model exactly as with the built-in ones.
 *
 * What the user's source must define (C++17, device code; this header is documentation, the declarations live in
 * blackjax_b200/csrc/bjx_row.cuh which the plug-in translation unit includes first):
 *
 *   namespace bjx_user {
 *   template <class R>
 *   struct Model {
 *     // optional per-kernel state, held in registers for the whole launch (e.g. float w[R::NS];)
 *     __device__ __forceinline__ void init(const bjx::UserCtx& u);          // once per kernel launch and chain row
 *     template <bool WANT_LOGP>
 *     __device__ __forceinline__ void value_and_grad(const bjx::UserCtx& u, const float (&q)[R::NS],
 *                                                    float (&g)[R::NS], float& logp) const;
 *   };
 *   }
 *
 * Layout.  One warp owns one chain row of length u.D.  Lane `u.lane` holds R::NS slots of it in registers; slot s is
 * element R::idx(s, u.lane) (vector layout for D % 4 == 0: slot 4j+v = element (32 j + lane) 4 + v; scalar layout
 * otherwise: slot s = element 32 s + lane).  Slots whose element index is >= u.D hold 0 in q and MUST be 0 in g.
 *
 *   u.theta, u.n_theta   the target's parameter block (device memory, read-only: data, hyper-parameters), as passed in
 *                        bjx_target_desc.user_params; read it with __ldg.
 *   u.row_smem           D floats of shared memory private to the warp.  bjx::row_stage<R>(u, q) copies the row there
 *                        (element e at u.row_smem[e], fenced with __syncwarp) when lanes need elements they do not own.
 *   bjx::row_at<R, E>(q) element E (compile-time index) broadcast from its owner's register.
 *   bjx::warp_sum(x)     all-lanes sum (xor-shuffle tree: the same value, bit for bit, on every lane).
 *   bjx::Vec<R::NS>      packed-FP32 helpers (axpy, mul, scale, add, dot_partial).
 *
 * init.  Called once per kernel launch for the chain row the warp owns, before any value_and_grad: load what every
 * evaluation needs (scales, a few hyper-parameters) into the struct's members so that it stays in registers across the
 * leapfrog steps of the launch.
 *
 * Results.  g = d logp / d q for the slots of this lane.  When WANT_LOGP is true, `logp` must come back with the
 * log-density, identical on all 32 lanes (finish reductions with bjx::warp_sum); when it is false the value is dead
 * (interior leapfrog steps of a fixed-length trajectory) and the reduction may be skipped.  The constant
 * bjx_target_desc.logp_offset is added by the caller.  The function must be deterministic and must not write global
 * memory.  All 32 lanes call it together (no divergent early return around warp-level primitives).
 *
 * Rows beyond a warp (1024 < dim <= 18432, dim % 4 == 0).  One CTA of bjx::kBigThreads (768) threads owns the chain row,
 * held in shared memory; the plug-in is built with -DBJX_PLUGIN_BIG=1 and the source defines instead
 *
 *   namespace bjx_user {
 *   struct BigModel {
 *     template <bool WANT_LOGP>
 *     __device__ static __forceinline__ float value_and_grad(const bjx::BigUserCtx& u, const float* q, float* g,
 *                                                            float* red);
 *   };
 *   }
 *
 * q[0..D) is complete on entry (shared memory, read-only); the function writes every g[i], i < D, and returns the
 * log-density on ALL threads (bjx::block_sum<NV>(acc, red), NV <= 5, sums per-thread partials over the CTA; `red` is its
 * scratch).  All kBigThreads threads call it together; the caller places the barriers before and after.  u.tid is
 * threadIdx.x; by convention thread t walks elements t, t + kBigThreads, ...  HMC, the leapfrog / init / momentum / energy
 * building blocks and diagonal metrics exist in this size class (as for the built-in targets); NUTS does not.
 * Examples: blackjax_b200/user_targets/diag_gaussian_big.cuh, hier_logit_big.cuh (BASELINE config 5's model).
 *
 * Example (the linear regression posterior of the reference's own sampling tests, tests/mcmc/test_sampling.py:103-111):
 * blackjax_b200/user_targets/linear_regression.cuh.
 */
#ifndef BJX_USER_TARGET_H_
#define BJX_USER_TARGET_H_
#include "bjx.h"
#endif
