// Large-D dense path (D > 128): HMC with a dense inverse mass matrix and/or a dense Gaussian target
// (BASELINE config 2: 1024-D correlated Gaussian, dense mass matrix).  The two linear maps of every
// leapfrog -- v = M^-1 p (blackjax/mcmc/integrators.py:242 -> metrics.py:263-270 -> util.py:57-61) and
// grad = -P q (the target's autodiff) -- are [C,D] x [D,D] GEMMs on the tensor cores (bjx_gemm.cu); the
// elementwise glue (half kicks, energies, accept/select) are the streaming row kernels below.
#include <vector>

#include "bjx_handle.h"
#include "bjx_internal.h"
#include <cuda_fp16.h>

#include "bjx_prng.cuh"

using namespace bjx;

namespace bjx {
size_t gemm_workspace_bytes(int M, int N, int K3);
int gemm_split(const void* Xs, const void* As, float* Y, const float* Cin, const float* row_alpha, float beta, int M, int N,
               int K3, void* workspace, cudaStream_t stream, bool double_kick);

constexpr int kRowWarps = 8;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- float32 -> 2 x binary16 operand split (see bjx_gemm.cu) ------------------------------------------------
// The power of two 2^s that lifts amax into [2^13, 2^14): high enough in binary16's range that the second term
// x2 = x - fp16(x) (~2^-11 |x|) of every element that matters stays a normal number, with two binades of headroom
// below 65504.  Zero / non-finite maxima leave the row unscaled (their NaN/inf propagate like the reference's).
__device__ __forceinline__ float pow2_lift(float amax) {
  if (!(amax > 0.f) || amax > 3.0e38f) return 1.f;
  int s = 13 - ((int)((__float_as_uint(amax) >> 23) & 0xffu) - 127);
  s = s < -126 ? -126 : (s > 126 ? 126 : s);
  return __uint_as_float((uint32_t)(s + 127) << 23);
}
__device__ __forceinline__ void split2(float x, uint16_t& h1, uint16_t& h2) {
  const __half a = __float2half_rn(x);
  const __half b = __float2half_rn(x - __half2float(a));
  h1 = __half_as_ushort(a);
  h2 = __half_as_ushort(b);
}
__device__ __forceinline__ uint2 pack4(const uint16_t (&v)[4]) {
  return make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));
}
// Plane stride: K rounded up to 8 halves so that every row of planes is 16-byte aligned for TMA; the pad columns
// are zeroed once at allocation and never written (zero x zero adds nothing to the product).
__host__ __device__ __forceinline__ int plane_stride(int K) { return (K + 7) & ~7; }
// the (x1, x2, x1) planes of four consecutive elements of an activation row, scaled by sc
__device__ __forceinline__ void store_planes(uint16_t* row, int KP, int k, float4 v, float sc) {
  uint16_t p1[4], p2[4];
  split2(v.x * sc, p1[0], p2[0]);
  split2(v.y * sc, p1[1], p2[1]);
  split2(v.z * sc, p1[2], p2[2]);
  split2(v.w * sc, p1[3], p2[3]);
  const uint2 u1 = pack4(p1), u2 = pack4(p2);
  *reinterpret_cast<uint2*>(row + k) = u1;
  *reinterpret_cast<uint2*>(row + (size_t)KP + k) = u2;
  *reinterpret_cast<uint2*>(row + (size_t)2 * KP + k) = u1;
}

// Activation rows, one warp per row: X [R, K] float32 -> X' [R, 3K] fp16 planes (x1, x2, x1) of 2^s_r x, and the
// per-row epilogue factor of the product that consumes them: row_alpha[r] = alpha_r * 2^-s_r * 2^-s_A
// (alpha_r = alpha_dev[r] * alpha, or alpha), with 2^-s_A the constant matrix's unscale factor.
__global__ void k_rows_split2(int R, int K, const float* __restrict__ x, uint16_t* __restrict__ xs, float alpha,
                              const float* __restrict__ alpha_dev, const float* __restrict__ mat_unscale,
                              float* __restrict__ row_alpha) {
  const int lane = threadIdx.x & 31, r = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (r >= R) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)r * K);
  float amax = 0.f;
  for (int i = lane; i < K / 4; i += 32) {
    const float4 v = xr[i];
    amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  const float sc = pow2_lift(wmax(amax));
  const int KP = plane_stride(K);
  uint16_t* row = xs + (size_t)r * 3 * KP;
  for (int i = lane; i < K / 4; i += 32) store_planes(row, KP, 4 * i, xr[i], sc);  // second read: L1/L2
  if (lane == 0) row_alpha[r] = (alpha_dev ? alpha_dev[r] * alpha : alpha) * (1.0f / sc) * mat_unscale[0];
}

// max |x| of a constant matrix into *out (float bits; non-negative floats order like ints); *out zeroed beforehand
__global__ void k_absmax(long long n4, const float* __restrict__ x, float* out) {
  float m = 0.f;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[t];
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  m = wmax(m);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}
// Constant matrices: A [R, K] float32 -> A' [R, 3K] fp16 planes (a1, a1, a2) of 2^s_A a; *unscale = 2^-s_A
__global__ void k_matrix_split2(long long R, int K, const float* __restrict__ x, uint16_t* __restrict__ xs,
                                const float* __restrict__ amax, float* __restrict__ unscale) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const float sc = pow2_lift(amax[0]);
  if (t == 0) unscale[0] = 1.0f / sc;
  if (t >= R * K / 4) return;
  const long long e = t * 4;
  const long long r = e / K;
  const int k = (int)(e % K);
  const float4 v = reinterpret_cast<const float4*>(x)[t];
  uint16_t p1[4], p2[4];
  split2(v.x * sc, p1[0], p2[0]);
  split2(v.y * sc, p1[1], p2[1]);
  split2(v.z * sc, p1[2], p2[2]);
  split2(v.w * sc, p1[3], p2[3]);
  const uint2 u1 = pack4(p1), u2 = pack4(p2);
  const int KP = plane_stride(K);
  uint16_t* row = xs + r * 3 * KP + k;
  *reinterpret_cast<uint2*>(row) = u1;
  *reinterpret_cast<uint2*>(row + (size_t)KP) = u1;
  *reinterpret_cast<uint2*>(row + (size_t)2 * KP) = u2;
}

// z[c, i] = normal(key_c, (D,))[i] with key_c = split(rng_key_c, 2)[0] when split_first (hmc.py:299,302 -> util.py:89-91)
__global__ void k_dense_normal(int C, int D, const uint32_t* __restrict__ keys, float* __restrict__ z, bool split_first,
                               int key_shared, uint32_t chain_offset) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n4 = (long long)C * D / 4;
  if (t >= n4) return;
  const long long e = t * 4;
  const int c = (int)(e / D);
  const uint32_t i = (uint32_t)(e % D);
  Key km = key_shared ? fold_in(Key{keys[0], keys[1]}, chain_offset + (uint32_t)c) : Key{keys[2 * c], keys[2 * c + 1]};
  if (split_first) km = fold_in(km, 0u);
  float4 o;
  o.x = normal_at(km, i);
  o.y = normal_at(km, i + 1);
  o.z = normal_at(km, i + 2);
  o.w = normal_at(km, i + 3);
  reinterpret_cast<float4*>(z)[t] = o;
}

// y[c,:] = a[:] * x[c,:]  (diagonal metric with a dense target)
__global__ void k_rows_scale(int C, int D, const float* __restrict__ a, long long a_stride, const float* __restrict__ x,
                             float* __restrict__ y) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n4 = (long long)C * D / 4;
  if (t >= n4) return;
  const long long e = t * 4;
  const int c = (int)(e / D);
  const int i = (int)(e % D);
  const float4 av = *reinterpret_cast<const float4*>(a + (size_t)c * a_stride + i);
  const float4 xv = __ldcs(reinterpret_cast<const float4*>(x) + t);
  __stcs(reinterpret_cast<float4*>(y) + t, make_float4(av.x * xv.x, av.y * xv.y, av.z * xv.z, av.w * xv.w));
}

// x[c,:] += (eps_c * coef) * y[c,:]     (integrators.py:199-203,235-239)
__global__ void k_rows_axpy(int C, int D, float* __restrict__ x, const float* __restrict__ y, float eps,
                            const float* __restrict__ eps_dev, float coef) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n4 = (long long)C * D / 4;
  if (t >= n4) return;
  const int c = (int)(t * 4 / D);
  const float a = (eps_dev ? eps_dev[c] : eps) * coef;
  float4 xv = __ldcs(reinterpret_cast<const float4*>(x) + t);
  const float4 yv = __ldcs(reinterpret_cast<const float4*>(y) + t);
  xv.x = fmaf(a, yv.x, xv.x); xv.y = fmaf(a, yv.y, xv.y); xv.z = fmaf(a, yv.z, xv.z); xv.w = fmaf(a, yv.w, xv.w);
  __stcs(reinterpret_cast<float4*>(x) + t, xv);
}

// e[c] = -logp[c] + 0.5 * sum_i v[c,i] p[c,i]      (trajectory.py:745-748, metrics.py:263-270)
__global__ void k_rows_energy(int C, int D, const float* __restrict__ v, const float* __restrict__ p,
                              const float* __restrict__ logp, float* __restrict__ e, float p_sign) {
  const int lane = threadIdx.x & 31, c = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (c >= C) return;
  const float4* vr = reinterpret_cast<const float4*>(v + (size_t)c * D);
  const float4* pr = reinterpret_cast<const float4*>(p + (size_t)c * D);
  float acc = 0.f;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 a = __ldcs(vr + i), b = __ldcs(pr + i);
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
  }
  acc = wsum(acc);
  if (lane == 0) e[c] = -logp[c] + 0.5f * acc * p_sign;
}

// Gradient + second half kick for one leapfrog, streaming the row:
//   TARGET 2 (dense Gaussian): g = -P q was written by the GEMM epilogue (alpha = -1); logp = 1/2 q.g + offset
//   TARGET 0 (diag Gaussian) : g = -(q - mean) / s^2;      logp = -1/2 sum (q-mean)^2/s^2 + offset
//   then (if p != null) p += (eps_c * 0.5) * g                                  (integrators.py:134-141)
//   and, when another leapfrog follows (kicks == 2), that step's first half kick p += (eps_c * 0.5) * g
//   (integrators.py:235-239) -- same gradient, two separately rounded FMAs, one pass over the row.
template <int TARGET>
__global__ void k_rows_grad_kick(int C, int D, const float* __restrict__ q, const float* __restrict__ aux,
                                 const float* __restrict__ inv_var, const float* __restrict__ mean, float offset,
                                 float* __restrict__ p, float eps, const float* __restrict__ eps_dev,
                                 float* __restrict__ g_out, float* __restrict__ logp_out, int kicks,
                                 uint16_t* __restrict__ p_split /* [C,3D] fp16 planes of the new p, or null */,
                                 float* __restrict__ row_alpha, const float* __restrict__ mat_unscale) {
  const int lane = threadIdx.x & 31, c = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (c >= C) return;
  const size_t ro = (size_t)c * D;
  const float eps_c = eps_dev ? eps_dev[c] : eps;
  const float eh = eps_c * 0.5f;
  float acc = 0.f, amax = 0.f;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 qv = __ldcs(reinterpret_cast<const float4*>(q + ro) + i);
    float4 gv;
    if (TARGET == 2) {
      gv = __ldcs(reinterpret_cast<const float4*>(aux + ro) + i);  // aux == g_out: already the gradient
      acc = fmaf(qv.x, gv.x, acc); acc = fmaf(qv.y, gv.y, acc); acc = fmaf(qv.z, gv.z, acc); acc = fmaf(qv.w, gv.w, acc);
    } else {
      const float4 w = __ldg(reinterpret_cast<const float4*>(inv_var) + i);
      float4 d = qv;
      if (mean) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + i);
        d = make_float4(qv.x - m.x, qv.y - m.y, qv.z - m.z, qv.w - m.w);
      }
      gv = make_float4(d.x * -w.x, d.y * -w.y, d.z * -w.z, d.w * -w.w);
      acc = fmaf(d.x, gv.x, acc); acc = fmaf(d.y, gv.y, acc); acc = fmaf(d.z, gv.z, acc); acc = fmaf(d.w, gv.w, acc);
    }
    if (TARGET != 2) __stcs(reinterpret_cast<float4*>(g_out + ro) + i, gv);
    if (p) {
      float4 pv = __ldcs(reinterpret_cast<const float4*>(p + ro) + i);
      pv.x = fmaf(eh, gv.x, pv.x); pv.y = fmaf(eh, gv.y, pv.y); pv.z = fmaf(eh, gv.z, pv.z); pv.w = fmaf(eh, gv.w, pv.w);
      if (kicks == 2) {
        pv.x = fmaf(eh, gv.x, pv.x); pv.y = fmaf(eh, gv.y, pv.y); pv.z = fmaf(eh, gv.z, pv.z); pv.w = fmaf(eh, gv.w, pv.w);
      }
      if (p_split) {
        reinterpret_cast<float4*>(p + ro)[i] = pv;  // re-read below: keep it cached
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(pv.x), fabsf(pv.y))), fmaxf(fabsf(pv.z), fabsf(pv.w)));
      } else {
        __stcs(reinterpret_cast<float4*>(p + ro) + i, pv);
      }
    }
  }
  if (p_split) {
    // the operand planes of the next step's M^-1 p product (and its per-row epilogue factor eps_c * 1.0 * 2^-s),
    // written while the row is hot: each lane re-reads exactly the elements it stored above
    const float sc = pow2_lift(wmax(amax));
    const int KP = plane_stride(D);
    uint16_t* row = p_split + (size_t)c * 3 * KP;
    for (int i = lane; i < D / 4; i += 32) store_planes(row, KP, 4 * i, reinterpret_cast<const float4*>(p + ro)[i], sc);
    if (lane == 0) row_alpha[c] = (eps_c * 1.0f) * (1.0f / sc) * mat_unscale[0];
  }
  acc = wsum(acc);
  if (lane == 0) logp_out[c] = 0.5f * acc + offset;
}

// Metropolis accept + select (hmc.py:158-163, proposal.py:214-235) for the dense path
__global__ void k_rows_accept(int C, int D, const uint32_t* __restrict__ keys, const float* __restrict__ e0,
                              const float* __restrict__ e1, float div_thr, const float* __restrict__ qw,
                              const float* __restrict__ gw, const float* __restrict__ lw, const float* q_in,
                              const float* g_in, const float* l_in, float* q_out, float* g_out, float* l_out, int L,
                              InfoPtrs info, int key_shared, uint32_t chain_offset) {
  const int lane = threadIdx.x & 31, c = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (c >= C) return;
  const Key rk = key_shared ? fold_in(Key{keys[0], keys[1]}, chain_offset + (uint32_t)c) : Key{keys[2 * c], keys[2 * c + 1]};
  const Key ki = fold_in(rk, 1u);
  float delta = e0[c] - e1[c];
  if (isnan(delta)) delta = -__int_as_float(0x7f800000);
  const bool is_div = (-delta) > div_thr;
  float pa = expf(delta);
  pa = pa > 1.0f ? 1.0f : pa;
  const bool acc = uniform01(ki) < pa;
  const size_t ro = (size_t)c * D;
  if (acc || q_out != q_in) {
    const float4* sq = reinterpret_cast<const float4*>((acc ? qw : q_in) + ro);
    const float4* sg = reinterpret_cast<const float4*>((acc ? gw : g_in) + ro);
    for (int i = lane; i < D / 4; i += 32) {
      __stcs(reinterpret_cast<float4*>(q_out + ro) + i, __ldcs(sq + i));
      __stcs(reinterpret_cast<float4*>(g_out + ro) + i, __ldcs(sg + i));
    }
    if (lane == 0) l_out[c] = acc ? lw[c] : l_in[c];
  }
  if (lane == 0) {
    if (info.acceptance_rate) info.acceptance_rate[c] = pa;
    if (info.is_accepted) info.is_accepted[c] = acc;
    if (info.is_divergent) info.is_divergent[c] = is_div;
    if (info.energy) info.energy[c] = e1[c];
    if (info.num_integration_steps) info.num_integration_steps[c] = L;
  }
}
}  // namespace bjx

#define DN_CUDA(call)                                                \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) return bjx_cuda_fail(h, e_, #call);       \
  } while (0)
#define DN_LAUNCH(where)                                             \
  do {                                                               \
    cudaError_t e_ = cudaGetLastError();                             \
    if (e_ != cudaSuccess) return bjx_cuda_fail(h, e_, where);       \
  } while (0)

static inline dim3 g4(long long n4) { return dim3((unsigned)((n4 + 255) / 256)); }
static inline dim3 grow(int C) { return dim3((C + kRowWarps - 1) / kRowWarps); }

struct DenseWs {
  float *p, *v, *q, *g, *lw, *e0, *e1;
  float* alpha;      // [C] per-row epilogue factors of the product whose operand planes sit in xs
  uint16_t* xs;      // [C, 3D] fp16 split activations
  float* mat_max;    // [3] max |a| of the constant matrices
  float* mat_unscale;  // [3] 2^-s_A
};
enum { MAT_IMM = 0, MAT_MSQRT = 1, MAT_PREC = 2 };

// The chain batch is processed as kParts independent slices on separate streams: the GEMMs are tensor-bound and
// the split / kick / energy row kernels are HBM-bound, so while one slice's GEMM occupies the MMA pipes the other
// slice's row kernels stream through HBM (chains never interact, so the slices share nothing but the constant
// matrices).  Slices fork from and join back into the handle's stream with events.
constexpr int kParts = 2;
struct Part {
  int c0, n;          // chains [c0, c0 + n)
  cudaStream_t st;
  void* gws;          // CUTLASS workspace of this slice
  int index;
};

static int dense_ws(bjx_handle_t h, DenseWs& w) {
  const size_t C = h->cfg.n_chains, D = h->cfg.dim;
  const size_t row = ((C * D * sizeof(float)) + 255) & ~(size_t)255, vec = ((C * sizeof(float)) + 255) & ~(size_t)255;
  const size_t KP = plane_stride((int)D);
  const size_t xsb = ((C * 3 * KP * sizeof(uint16_t)) + 255) & ~(size_t)255;
  const size_t mat = ((D * 3 * KP * sizeof(uint16_t)) + 255) & ~(size_t)255;
  const size_t need = 4 * row + 4 * vec + xsb + 3 * mat + 256;
  if (h->dense_bytes < need) {
    if (h->dense_block) DN_CUDA(cudaFree(h->dense_block));
    h->dense_block = nullptr;
    DN_CUDA(cudaMalloc((void**)&h->dense_block, need));
    h->dense_bytes = need;
  }
  const size_t gw = ((gemm_workspace_bytes((int)C, (int)D, (int)(3 * KP)) + 255) & ~(size_t)255) + 256;
  if (kParts * gw > h->gemm_ws_bytes) {
    if (h->gemm_ws) DN_CUDA(cudaFree(h->gemm_ws));
    h->gemm_ws = nullptr;
    DN_CUDA(cudaMalloc(&h->gemm_ws, kParts * gw));
    h->gemm_ws_bytes = kParts * gw;
  }
  if (!h->dense_streams_ready) {
    for (int k = 0; k < kParts; ++k) {
      DN_CUDA(cudaStreamCreateWithFlags(&h->dense_stream[k], cudaStreamNonBlocking));
      DN_CUDA(cudaEventCreateWithFlags(&h->dense_join[k], cudaEventDisableTiming));
    }
    DN_CUDA(cudaEventCreateWithFlags(&h->dense_fork, cudaEventDisableTiming));
    DN_CUDA(cudaEventCreateWithFlags(&h->dense_stagger, cudaEventDisableTiming));
    h->dense_streams_ready = true;
  }
  char* b = (char*)h->dense_block;
  w.p = (float*)b; w.v = (float*)(b + row); w.q = (float*)(b + 2 * row); w.g = (float*)(b + 3 * row);
  w.lw = (float*)(b + 4 * row); w.e0 = (float*)(b + 4 * row + vec); w.e1 = (float*)(b + 4 * row + 2 * vec);
  w.alpha = (float*)(b + 4 * row + 3 * vec);
  w.xs = (uint16_t*)(b + 4 * row + 4 * vec);
  for (int m = 0; m < 3; ++m) h->dense_mat_s[m] = (uint16_t*)(b + 4 * row + 4 * vec + xsb + m * mat);
  w.mat_max = (float*)(b + 4 * row + 4 * vec + xsb + 3 * mat);
  w.mat_unscale = w.mat_max + 4;
  if (h->dense_bytes_built != need) {  // fresh block: zero the plane pads, every split matrix must be rebuilt
    DN_CUDA(cudaMemsetAsync(w.xs, 0, xsb + 3 * mat, h->stream));
    for (int m = 0; m < 3; ++m) h->dense_mat_src[m] = nullptr;
    h->dense_bytes_built = need;
  }
  // (re)build the split copies of the constant matrices whose source changed (set_metric / set_target)
  const float* src[3] = {h->metric_kind == BJX_METRIC_DENSE ? h->imm : nullptr,
                         h->metric_kind == BJX_METRIC_DENSE ? h->msqrt : nullptr,
                         h->cfg.target.kind == BJX_TARGET_DENSE_GAUSSIAN ? h->cfg.target.precision : nullptr};
  for (int m = 0; m < 3; ++m) {
    if (src[m] && (h->dense_mat_src[m] != src[m] || h->dense_mat_ver[m] != h->dense_version)) {
      DN_CUDA(cudaMemsetAsync(w.mat_max + m, 0, sizeof(float), h->stream));
      k_absmax<<<148, 256, 0, h->stream>>>((long long)D * D / 4, src[m], w.mat_max + m);
      DN_LAUNCH("k_absmax");
      k_matrix_split2<<<g4((long long)D * D / 4), 256, 0, h->stream>>>((long long)D, (int)D, src[m], h->dense_mat_s[m],
                                                                      w.mat_max + m, w.mat_unscale + m);
      DN_LAUNCH("k_matrix_split2");
      h->dense_mat_src[m] = src[m];
      h->dense_mat_ver[m] = h->dense_version;
    }
  }
  return 0;
}

// Run fn(part) for every slice: fork from the handle's stream, one stream per slice, join back.
template <class F>
static int for_parts(bjx_handle_t h, F fn) {
  const int C = h->cfg.n_chains;
  const int parts = (C >= 8192) ? kParts : 1;
  const size_t gw = h->gemm_ws_bytes / kParts;
  if (parts == 1) {
    Part pt{0, C, h->stream, h->gemm_ws, -1};
    return fn(pt);
  }
  DN_CUDA(cudaEventRecord(h->dense_fork, h->stream));
  int rc = 0;
  h->dense_stagger_armed = true;  // slice 0 records dense_stagger once its first operand split is enqueued
  for (int k = 0; k < parts; ++k) {
    const int c0 = (int)((long long)C * k / parts) & ~7, c1 = (k + 1 == parts) ? C : ((int)((long long)C * (k + 1) / parts) & ~7);
    Part pt{c0, c1 - c0, h->dense_stream[k], (char*)h->gemm_ws + k * gw, k};
    DN_CUDA(cudaStreamWaitEvent(pt.st, h->dense_fork, 0));
    // phase offset: slice 1 starts when slice 0 reaches its first GEMM, so that from then on one slice's tensor-bound
    // GEMM runs beside the other slice's HBM-bound row kernels instead of both slices doing the same thing at once
    if (k > 0 && !h->dense_stagger_armed) DN_CUDA(cudaStreamWaitEvent(pt.st, h->dense_stagger, 0));
    if (rc == 0) rc = fn(pt);
    DN_CUDA(cudaEventRecord(h->dense_join[k], pt.st));
    DN_CUDA(cudaStreamWaitEvent(h->stream, h->dense_join[k], 0));
  }
  return rc;
}

// Y[c,:] = alpha_c * (X . A^T)[c,:] + beta * Cin[c,:] for the slice, alpha_c = alpha * alpha_dev[c] (or alpha), A one of the
// handle's constant matrices (float32-accurate, bjx_gemm.cu).  X, Y, Cin, alpha_dev are FULL [C,D] / [C] arrays; the
// slice's rows are addressed here.  presplit: the producer of X already wrote its operand planes and w.alpha.
// double_kick: Y = alpha_c acc + (alpha_c acc + Cin) instead (the two half kicks between consecutive leapfrog steps).
static int gemm(bjx_handle_t h, DenseWs& w, const Part& pt, const float* X, int mat, float* Y, const float* Cin, float alpha,
                const float* alpha_dev, float beta, bool presplit = false, bool double_kick = false) {
  const int D = h->cfg.dim;
  const size_t ro = (size_t)pt.c0 * D;
  const int KP = plane_stride(D);
  uint16_t* xs = w.xs + (size_t)pt.c0 * 3 * KP;
  if (!presplit) {
    k_rows_split2<<<grow(pt.n), kRowWarps * 32, 0, pt.st>>>(pt.n, D, X + ro, xs, alpha, alpha_dev ? alpha_dev + pt.c0 : nullptr,
                                                          w.mat_unscale + mat, w.alpha + pt.c0);
    DN_LAUNCH("k_rows_split2");
  }
  if (pt.index == 0 && h->dense_stagger_armed) {
    DN_CUDA(cudaEventRecord(h->dense_stagger, pt.st));
    h->dense_stagger_armed = false;
  }
  const int rc = gemm_split(xs, h->dense_mat_s[mat], Y + ro, Cin ? Cin + ro : nullptr, w.alpha + pt.c0, beta, pt.n, D, 3 * KP,
                            pt.gws, pt.st, double_kick);
  if (rc) return bjx_fail(h, BJX_E_UNSUPPORTED, "tensor-core GEMM failed (cutlass status " + std::to_string(rc) + ")");
  DN_LAUNCH("gemm");
  return 0;
}

// v = M^-1 p
static int dense_velocity(bjx_handle_t h, DenseWs& w, const Part& pt, const float* p, float* v) {
  const int D = h->cfg.dim;
  if (h->metric_kind == BJX_METRIC_DENSE) return gemm(h, w, pt, p, MAT_IMM, v, nullptr, 1.f, nullptr, 0.f);
  const bool per_chain = (h->metric_kind == BJX_METRIC_DIAG_PER_CHAIN);
  const size_t ro = (size_t)pt.c0 * D;
  k_rows_scale<<<g4((long long)pt.n * D / 4), 256, 0, pt.st>>>(pt.n, D, h->imm + (per_chain ? ro : 0), per_chain ? D : 0,
                                                              p + ro, v + ro);
  DN_LAUNCH("k_rows_scale");
  return 0;
}

// g, logp = value_and_grad(q); optionally p += eh * g (once or twice).  aux: [C,D] scratch for P q.
static int dense_grad(bjx_handle_t h, DenseWs& w, const Part& pt, const float* q, float* aux, float* p, float eps,
                      const float* eps_dev, float* g, float* logp, int kicks = 1, bool split_p = false) {
  const int D = h->cfg.dim;
  const bjx_target_desc& t = h->cfg.target;
  const size_t ro = (size_t)pt.c0 * D;
  const float* ed = eps_dev ? eps_dev + pt.c0 : nullptr;
  float* pp = p ? p + ro : nullptr;
  uint16_t* ps = (split_p && pp) ? w.xs + (size_t)pt.c0 * 3 * plane_stride(D) : nullptr;
  (void)aux;
  if (t.kind == BJX_TARGET_DENSE_GAUSSIAN) {
    int rc = gemm(h, w, pt, q, MAT_PREC, g, nullptr, -1.f, nullptr, 0.f);  // g = -(q P): the sign rides on the GEMM epilogue
    if (rc) return rc;
    k_rows_grad_kick<2><<<grow(pt.n), kRowWarps * 32, 0, pt.st>>>(pt.n, D, q + ro, g + ro, nullptr, nullptr, t.logp_offset,
                                                                pp, eps, ed, g + ro, logp + pt.c0, kicks, ps, w.alpha + pt.c0,
                                                                w.mat_unscale + MAT_IMM);
  } else if (t.kind == BJX_TARGET_DIAG_GAUSSIAN) {
    k_rows_grad_kick<0><<<grow(pt.n), kRowWarps * 32, 0, pt.st>>>(pt.n, D, q + ro, nullptr, t.inv_var, t.mean, t.logp_offset,
                                                                pp, eps, ed, g + ro, logp + pt.c0, kicks, ps, w.alpha + pt.c0,
                                                                w.mat_unscale + MAT_IMM);
  } else {
    return bjx_fail(h, BJX_E_UNSUPPORTED, "large-D dense path supports DENSE_GAUSSIAN and DIAG_GAUSSIAN targets");
  }
  DN_LAUNCH("k_rows_grad_kick");
  return 0;
}

static int dense_momentum(bjx_handle_t h, DenseWs& w, const Part& pt, const uint32_t* keys, float* p_out, bool split_first) {
  const int D = h->cfg.dim;
  const size_t ro = (size_t)pt.c0 * D;
  const bool dense_m = (h->metric_kind == BJX_METRIC_DENSE);
  float* z = dense_m ? w.v : p_out;
  const uint32_t* kp = h->key_shared ? keys : keys + 2 * (size_t)pt.c0;
  k_dense_normal<<<g4((long long)pt.n * D / 4), 256, 0, pt.st>>>(pt.n, D, kp, z + ro, split_first, h->key_shared,
                                                                h->chain_offset + (uint32_t)pt.c0);
  DN_LAUNCH("k_dense_normal");
  if (dense_m) return gemm(h, w, pt, z, MAT_MSQRT, p_out, nullptr, 1.f, nullptr, 0.f);  // p = L^-T z
  const bool per_chain = (h->metric_kind == BJX_METRIC_DIAG_PER_CHAIN);
  k_rows_scale<<<g4((long long)pt.n * D / 4), 256, 0, pt.st>>>(pt.n, D, h->msqrt + (per_chain ? ro : 0), per_chain ? D : 0,
                                                              z + ro, p_out + ro);
  DN_LAUNCH("k_rows_scale");
  return 0;
}

static int dense_energy(bjx_handle_t h, DenseWs& w, const Part& pt, const float* p, const float* logp, float* e_out) {
  const int D = h->cfg.dim;
  int rc = dense_velocity(h, w, pt, p, w.v);
  if (rc) return rc;
  const size_t ro = (size_t)pt.c0 * D;
  k_rows_energy<<<grow(pt.n), kRowWarps * 32, 0, pt.st>>>(pt.n, D, w.v + ro, p + ro, logp + pt.c0, e_out + pt.c0, 1.f);
  DN_LAUNCH("k_rows_energy");
  return 0;
}

// n velocity-Verlet steps in place (integrators.py:104-150)
static int dense_leapfrog_core(bjx_handle_t h, DenseWs& w, const Part& pt, float* q, float* p, float* logp, float* g,
                               float eps, const float* eps_dev, int n_steps) {
  const int D = h->cfg.dim;
  const long long n4 = (long long)pt.n * D / 4;
  const size_t ro = (size_t)pt.c0 * D;
  const float* ed = eps_dev ? eps_dev + pt.c0 : nullptr;
  if (n_steps > 0) {
    k_rows_axpy<<<g4(n4), 256, 0, pt.st>>>(pt.n, D, p + ro, g + ro, eps, ed, 0.5f);  // first half kick p += (eps/2) g
    DN_LAUNCH("k_rows_axpy");
  }
  const bool dense_m = (h->metric_kind == BJX_METRIC_DENSE);
  bool p_presplit = false;
  for (int s = 0; s < n_steps; ++s) {
    int rc;
    const bool presplit = p_presplit;  // the previous step's kick kernel left split(p) in w.xs
    if (dense_m) {
      // q = q + (eps_c * 1.0) * (p M^-1): the axpy rides on the GEMM epilogue's per-row factor
      rc = gemm(h, w, pt, p, MAT_IMM, q, q, eps_dev ? 1.0f : eps * 1.0f, eps_dev, 1.f, presplit);
      if (rc) return rc;
    } else {
      rc = dense_velocity(h, w, pt, p, w.v);
      if (rc) return rc;
      k_rows_axpy<<<g4(n4), 256, 0, pt.st>>>(pt.n, D, q + ro, w.v + ro, eps, ed, 1.0f);
      DN_LAUNCH("k_rows_axpy");
    }
    const bool more = (s + 1 < n_steps);
    if (more && dense_m && h->cfg.target.kind == BJX_TARGET_DENSE_GAUSSIAN) {
      // Between two steps neither g nor logp is observable, only p += (eps_c/2) g twice (this step's second half kick
      // and the next step's first, integrators.py:134-141,235-239) with g = -(q P): both FMAs ride on the epilogue of
      // the gradient product, per-row factor -(eps_c * 0.5) (bit-identical to kicking with the stored gradient, since
      // the row scales are powers of two).  The next iteration splits the new p itself.
      rc = gemm(h, w, pt, q, MAT_PREC, p, p, eps_dev ? -0.5f : -(eps * 0.5f), eps_dev, 1.f, false, true);
      if (rc) return rc;
      p_presplit = false;
      continue;
    }
    // g, logp at the new q; p += (eps/2) g; plus the next step's first half kick -- and, with a dense metric, the
    // operand planes of that step's M^-1 p product -- when one follows
    rc = dense_grad(h, w, pt, q, w.v, p, eps, eps_dev, g, logp, more ? 2 : 1, more && dense_m);
    if (rc) return rc;
    p_presplit = more && dense_m;
  }
  return 0;
}

int bjx_dense_velocity(bjx_handle_t h, const float* p, float* v) {
  DenseWs w;
  int rc = dense_ws(h, w);
  if (rc) return rc;
  return for_parts(h, [&](const Part& pt) { return dense_velocity(h, w, pt, p, v); });
}

int bjx_dense_init_state(bjx_handle_t h, const float* q, float* logp_out, float* grad_out) {
  DenseWs w;
  int rc = dense_ws(h, w);
  if (rc) return rc;
  return for_parts(h, [&](const Part& pt) { return dense_grad(h, w, pt, q, w.v, nullptr, 0.f, nullptr, grad_out, logp_out); });
}

int bjx_dense_sample_momentum(bjx_handle_t h, const uint32_t* keys, float* p_out, bool split_first) {
  DenseWs w;
  int rc = dense_ws(h, w);
  if (rc) return rc;
  return for_parts(h, [&](const Part& pt) { return dense_momentum(h, w, pt, keys, p_out, split_first); });
}

int bjx_dense_energy(bjx_handle_t h, const float* p, const float* logp, float* e_out) {
  DenseWs w;
  int rc = dense_ws(h, w);
  if (rc) return rc;
  return for_parts(h, [&](const Part& pt) { return dense_energy(h, w, pt, p, logp, e_out); });
}

int bjx_dense_leapfrog(bjx_handle_t h, float* q, float* p, float* logp, float* g, float eps, const float* eps_dev,
                       int n_steps) {
  DenseWs w;
  int rc = dense_ws(h, w);
  if (rc) return rc;
  return for_parts(h, [&](const Part& pt) { return dense_leapfrog_core(h, w, pt, q, p, logp, g, eps, eps_dev, n_steps); });
}

static int dense_hmc_part(bjx_handle_t h, DenseWs& w, const Part& pt, const uint32_t* keys, const float* q_in,
                          const float* logp_in, const float* g_in, float* q_out, float* logp_out, float* g_out, float eps,
                          const float* eps_dev, int L, const InfoPtrs& info) {
  const int D = h->cfg.dim;
  const size_t ro = (size_t)pt.c0 * D;
  const size_t bytes = (size_t)pt.n * D * sizeof(float);
  int rc = dense_momentum(h, w, pt, keys, w.p, true);  // hmc.py:299-302
  if (rc) return rc;
  if (info.momentum) DN_CUDA(cudaMemcpyAsync(info.momentum + ro, w.p + ro, bytes, cudaMemcpyDeviceToDevice, pt.st));
  rc = dense_energy(h, w, pt, w.p, logp_in, w.e0);  // hmc.py:159
  if (rc) return rc;
  DN_CUDA(cudaMemcpyAsync(w.q + ro, q_in + ro, bytes, cudaMemcpyDeviceToDevice, pt.st));
  DN_CUDA(cudaMemcpyAsync(w.g + ro, g_in + ro, bytes, cudaMemcpyDeviceToDevice, pt.st));
  DN_CUDA(cudaMemcpyAsync(w.lw + pt.c0, logp_in + pt.c0, (size_t)pt.n * sizeof(float), cudaMemcpyDeviceToDevice, pt.st));
  rc = dense_leapfrog_core(h, w, pt, w.q, w.p, w.lw, w.g, eps, eps_dev, L);  // trajectory.py:165
  if (rc) return rc;
  // kinetic energy of the (flipped) end momentum: (-p)^T M^-1 (-p) = p^T M^-1 p   (hmc.py:158-160)
  rc = dense_energy(h, w, pt, w.p, w.lw, w.e1);
  if (rc) return rc;
  if (info.proposal_position) DN_CUDA(cudaMemcpyAsync(info.proposal_position + ro, w.q + ro, bytes, cudaMemcpyDeviceToDevice, pt.st));
  if (info.proposal_momentum) {  // flipped momentum (hmc.py:158)
    DN_CUDA(cudaMemsetAsync(info.proposal_momentum + ro, 0, bytes, pt.st));
    k_rows_axpy<<<g4((long long)pt.n * D / 4), 256, 0, pt.st>>>(pt.n, D, info.proposal_momentum + ro, w.p + ro, -1.0f, nullptr, 1.0f);
    DN_LAUNCH("k_rows_axpy");
  }
  InfoPtrs ip = info;  // per-chain outputs of this slice
  if (ip.acceptance_rate) ip.acceptance_rate += pt.c0;
  if (ip.is_accepted) ip.is_accepted += pt.c0;
  if (ip.is_divergent) ip.is_divergent += pt.c0;
  if (ip.energy) ip.energy += pt.c0;
  if (ip.num_integration_steps) ip.num_integration_steps += pt.c0;
  const uint32_t* kp = h->key_shared ? keys : keys + 2 * (size_t)pt.c0;
  k_rows_accept<<<grow(pt.n), kRowWarps * 32, 0, pt.st>>>(pt.n, D, kp, w.e0 + pt.c0, w.e1 + pt.c0, h->cfg.divergence_threshold,
                                                         w.q + ro, w.g + ro, w.lw + pt.c0, q_in + ro, g_in + ro, logp_in + pt.c0,
                                                         q_out + ro, g_out + ro, logp_out + pt.c0, L, ip, h->key_shared,
                                                         h->chain_offset + (uint32_t)pt.c0);
  DN_LAUNCH("k_rows_accept");
  return 0;
}

int bjx_dense_hmc_step(bjx_handle_t h, const uint32_t* keys, const float* q_in, const float* logp_in, const float* g_in,
                       float* q_out, float* logp_out, float* g_out, float eps, const float* eps_dev, int L,
                       const InfoPtrs& info) {
  DenseWs w;
  int rc = dense_ws(h, w);
  if (rc) return rc;
  return for_parts(h, [&](const Part& pt) {
    return dense_hmc_part(h, w, pt, keys, q_in, logp_in, g_in, q_out, logp_out, g_out, eps, eps_dev, L, info);
  });
}
