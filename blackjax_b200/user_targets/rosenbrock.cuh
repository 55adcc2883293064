// A neighbour-coupled model as a user-defined target (include/bjx_user_target.h): the D-dimensional Rosenbrock density
//
//   logp(x) = -beta * sum_{i=0}^{D-2} [ a (x_{i+1} - x_i^2)^2 + (1 - x_i)^2 ]        theta = [a, beta]
//   d/dx_i  = -beta * [ -4 a x_i (x_{i+1} - x_i^2) - 2 (1 - x_i)   (i <= D-2)
//                       + 2 a (x_i - x_{i-1}^2)                    (i >= 1) ]
//
// Every element needs its two neighbours, which other lanes (or other slots) hold: the row is staged in the warp's
// shared-memory scratch (bjx::row_stage) and read back by element index -- the pattern for any model whose terms couple
// elements (chains, lattices, small dense blocks), at every row size class of the kernels.
#pragma once
namespace bjx_user {
template <class R>
struct Model {
  float a, beta;
  __device__ __forceinline__ void init(const bjx::UserCtx& u) {
    a = __ldg(u.theta);
    beta = __ldg(u.theta + 1);
  }

  template <bool WANT_LOGP>
  __device__ __forceinline__ void value_and_grad(const bjx::UserCtx& u, const float (&q)[R::NS], float (&g)[R::NS],
                                                 float& logp) const {
    bjx::row_stage<R>(u, q);
    const float* x = u.row_smem;
    float acc = 0.f;
#pragma unroll
    for (int s = 0; s < R::NS; ++s) {
      const int e = R::idx(s, u.lane);
      float gv = 0.f;
      if (e < u.D) {
        const float xi = q[s];
        float d = 0.f;
        if (e + 1 < u.D) {
          const float t = x[e + 1] - xi * xi;
          const float o = 1.0f - xi;
          d = -4.0f * a * xi * t - 2.0f * o;
          if (WANT_LOGP) acc += a * t * t + o * o;
        }
        if (e >= 1) {
          const float xm = x[e - 1];
          d += 2.0f * a * (xi - xm * xm);
        }
        gv = -beta * d;
      }
      g[s] = gv;
    }
    if (WANT_LOGP) logp = -beta * bjx::warp_sum(acc);
  }
};
}  // namespace bjx_user
