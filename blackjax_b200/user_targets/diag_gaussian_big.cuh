// The diagonal Gaussian as a user-defined target for rows BEYOND A WARP (1024 < dim <= 18432): the CTA-level contract of
// include/bjx_user_target.h (bjx_user::BigModel).  One CTA of bjx::kBigThreads threads owns the chain row; q (complete on
// entry) and g live in shared memory; thread t conventionally walks elements t, t + kBigThreads, ...
// Same arithmetic as the built-in BJX_TARGET_DIAG_GAUSSIAN big-row branch: the draws are bit-identical.
//
//   logp(x) = -1/2 sum_i x_i^2 / s_i^2          theta = [1/s_0^2 ... 1/s_{D-1}^2]
#pragma once
namespace bjx_user {
struct BigModel {
  template <bool WANT_LOGP>
  __device__ static __forceinline__ float value_and_grad(const bjx::BigUserCtx& u, const float* q, float* g, float* red) {
    float acc[1] = {0.f};
    for (int i = u.tid; i < u.D; i += bjx::kBigThreads) {
      const float d = q[i];
      const float t = d * -__ldg(u.theta + i);
      acc[0] = fmaf(d, t, acc[0]);
      g[i] = t;
    }
    if constexpr (!WANT_LOGP) return 0.f;   // the value is dead in interior leapfrog steps
    bjx::block_sum<1>(acc, red);            // block-wide sum, identical on every thread
    return 0.5f * acc[0];
  }
};
}  // namespace bjx_user
