// Bayesian linear regression: the posterior the reference's sampling tests run every sampler on
// (tests/mcmc/test_sampling.py:103-111 `regression_logprob`, used by test_window_adaptation :322-379 and the MALA / GHMC /
// MEADS / ChEES cases), as a user-defined target (include/bjx_user_target.h), generalised from one coefficient to K <= 16.
//
//   position  x = [log_scale, coefs_0 .. coefs_{K-1}]                      D = 1 + K
//   scale = exp(log_scale)
//   logp  = expon.logpdf(scale, 0, 1) + log_scale                          (-scale + log_scale)
//         + sum_k norm.logpdf(coefs_k, 0, 5)
//         + sum_n norm.logpdf(y_n, X_n . coefs, scale)
//   d/d log_scale = -scale + 1 - N + exp(-2 log_scale) sum_n r_n^2         r_n = y_n - X_n . coefs
//   d/d coefs_k   = -coefs_k / 25 + exp(-2 log_scale) sum_n r_n X_nk
//
//   theta = [N, K, X (N x K row-major), y (N)]   (N, K stored as floats)
//
// Each lane walks the observations n = lane, lane + 32, ...; the K + 1 sums finish with warp shuffles.
#pragma once
namespace bjx_user {
constexpr int kMaxCoefs = 16;

template <class R>
struct Model {
  __device__ __forceinline__ void init(const bjx::UserCtx&) {}  // nothing to keep between calls: the data is read per call

  template <bool WANT_LOGP>
  __device__ __forceinline__ void value_and_grad(const bjx::UserCtx& u, const float (&q)[R::NS], float (&g)[R::NS],
                                                 float& logp) const {
  const int N = (int)__ldg(u.theta), K = u.D - 1;
  const float* __restrict__ X = u.theta + 2;
  const float* __restrict__ y = X + (size_t)N * K;
  bjx::row_stage<R>(u, q);  // every lane needs every coefficient
  const float ls = u.row_smem[0];
  float c[kMaxCoefs], acc[kMaxCoefs];
#pragma unroll
  for (int k = 0; k < kMaxCoefs; ++k) {
    c[k] = (k < K) ? u.row_smem[1 + k] : 0.f;
    acc[k] = 0.f;
  }
  float ss = 0.f;
  for (int n = u.lane; n < N; n += 32) {
    const float* xr = X + (size_t)n * K;
    float xv[kMaxCoefs];
    float pred = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxCoefs; ++k) {
      xv[k] = (k < K) ? __ldg(xr + k) : 0.f;
      pred = fmaf(xv[k], c[k], pred);
    }
    const float r = __ldg(y + n) - pred;
    ss = fmaf(r, r, ss);
#pragma unroll
    for (int k = 0; k < kMaxCoefs; ++k) acc[k] = fmaf(r, xv[k], acc[k]);
  }
  ss = bjx::warp_sum(ss);
#pragma unroll
  for (int k = 0; k < kMaxCoefs; ++k)
    if (k < K) acc[k] = bjx::warp_sum(acc[k]);
  const float scale = expf(ls);
  const float w = expf(-2.0f * ls);  // 1 / scale^2
  float cc = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxCoefs; ++k) cc = fmaf(c[k], c[k], cc);
#pragma unroll
  for (int s = 0; s < R::NS; ++s) {
    const int e = R::idx(s, u.lane);
    float v = 0.f;
    if (e == 0) v = (w * ss - scale) + (1.0f - (float)N);
#pragma unroll
    for (int k = 0; k < kMaxCoefs; ++k)
      if (e == k + 1 && k < K) v = fmaf(w, acc[k], -c[k] / 25.0f);
    g[s] = v;
  }
  if (WANT_LOGP) {
    const float kLog5 = 1.6094379124341003f, kHalfLog2Pi = 0.9189385332046727f;
    logp = (ls - scale) - (cc / 50.0f + (float)K * (kLog5 + kHalfLog2Pi)) -
           (0.5f * w * ss + (float)N * (ls + kHalfLog2Pi));
  }
  }
};
}  // namespace bjx_user
