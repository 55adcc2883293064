// The hierarchical logistic regression of BASELINE config 5 (blackjax_b200/targets.py HierLogit) as a user-defined big-row
// target (bjx_user::BigModel, include/bjx_user_target.h): what a user with a 10^4-parameter hierarchical model writes.
//
//   x = [mu, log_tau, beta0, beta1, alpha_0 .. alpha_{G-1}],  D = 4 + G;  8 Bernoulli-logit observations per group
//   theta = [G, 0, 0, 0,  covariates (G x 8 x 2),  outcome bits (G values in 0..255)]      (16-byte aligned covariates)
//
// Same arithmetic and summation order as the built-in one-chain-per-CTA branch (csrc/bjx_big.cuh), so bjx_init_state and
// bjx_leapfrog agree with the built-in target bit for bit; the built-in HMC transition runs the two-chains-per-CTA kernel
// (another summation order), a plug-in runs k_big_hmc.
#pragma once
namespace bjx_user {
struct BigModel {
  template <bool WANT_LOGP>
  __device__ static __forceinline__ float value_and_grad(const bjx::BigUserCtx& u, const float* q, float* g, float* red) {
    const int tid = u.tid, G = (int)__ldg(u.theta);
    const float* data_x = u.theta + 4;
    const float* data_y = data_x + (size_t)G * 16;
    const float mu = q[0], lt = q[1], b0 = q[2], b1 = q[3];
    const float e2 = expf(-2.0f * lt);
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // ll, sum d, sum d^2, grad b0, grad b1
    for (int gidx = tid; gidx < G; gidx += bjx::kBigThreads) {
      const float4* xr = reinterpret_cast<const float4*>(data_x + (size_t)gidx * 16);
      const unsigned bits = (unsigned)__ldg(data_y + gidx);
      const float alpha = q[4 + gidx];
      const float d = alpha - mu;
      float ga = 0.f;
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) {
        const float4 xv = __ldg(xr + k2);
        const float xs[2][2] = {{xv.x, xv.y}, {xv.z, xv.w}};
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const bool yb = (bits >> (2 * k2 + v)) & 1u;
          const float eta = alpha + b0 * xs[v][0] + b1 * xs[v][1];
          float ex, rc;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fabsf(eta) * -1.4426950408889634f));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(1.0f + ex));
          const float sig = (eta >= 0.f) ? rc : ex * rc;
          const float r = (yb ? 1.0f : 0.0f) - sig;
          if constexpr (WANT_LOGP) {
            float l2;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(rc));
            const float softplus = fmaf(-0.69314718f, l2, fmaxf(eta, 0.f));
            acc[0] += (yb ? eta : 0.0f) - softplus;
          }
          ga += r;
          acc[3] = fmaf(r, xs[v][0], acc[3]);
          acc[4] = fmaf(r, xs[v][1], acc[4]);
        }
      }
      acc[1] += d;
      acc[2] = fmaf(d, d, acc[2]);
      g[4 + gidx] = -d * e2 + ga;
    }
    bjx::block_sum<5>(acc, red);
    if (tid == 0) {
      g[0] = -0.01f * mu + e2 * acc[1];
      g[1] = -lt + e2 * acc[2] - (float)G;
      g[2] = -0.16f * b0 + acc[3];
      g[3] = -0.16f * b1 + acc[4];
    }
    return -0.005f * mu * mu - 0.5f * lt * lt - 0.08f * (b0 * b0 + b1 * b1) + (-0.5f * e2 * acc[2] - (float)G * lt) + acc[0];
  }
};
}  // namespace bjx_user
