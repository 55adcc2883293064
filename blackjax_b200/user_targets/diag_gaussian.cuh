// The diagonal Gaussian of tests/fixtures.py:60-78 (`std_normal_logdensity(x, scale)`) written as a user-defined target:
// the smallest complete example of the contract in include/bjx_user_target.h, and the plumbing check of the plug-in
// path (its draws are bit-identical to the built-in BJX_TARGET_DIAG_GAUSSIAN and it runs at the same speed;
// tests/test_gpu_user_target.py, scripts/bench_user_target.py).
//
//   logp(x) = -1/2 sum_i x_i^2 / s_i^2          theta = [1/s_0^2 ... 1/s_{D-1}^2]
#pragma once
namespace bjx_user {
template <class R>
struct Model {
  float w[R::NS];  // -1/s^2 for this lane's slots, loaded once per kernel (slots past D come back 0, so g stays 0 there)

  __device__ __forceinline__ void init(const bjx::UserCtx& u) {
    R::load_const(w, u.theta, u.D, u.lane);
#pragma unroll
    for (int s = 0; s < R::NS; ++s) w[s] = -w[s];
  }

  template <bool WANT_LOGP>
  __device__ __forceinline__ void value_and_grad(const bjx::UserCtx& u, const float (&q)[R::NS], float (&g)[R::NS],
                                                 float& logp) const {
    bjx::Vec<R::NS>::mul(g, q, w);  // g_i = -x_i / s_i^2
    if (WANT_LOGP) logp = 0.5f * bjx::warp_sum(bjx::Vec<R::NS>::dot_partial(q, g));
  }
};
}  // namespace bjx_user
