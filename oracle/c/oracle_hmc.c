/*
 * C/pthreads twin of the numpy oracle's HMC transition (oracle/hmc.py), used ONLY as the CPU
 * timing stand-in (bench.py cpu_baseline / --impl reference) and cross-checked against the numpy
 * oracle in tests/test_oracle_c.py.  TEST INFRASTRUCTURE -- never linked into the product.
 *
 * "restated oracle, not JAX": the reference (blackjax-devs/blackjax) needs jax 0.10.0, which is not
 * installable here, so this restates blackjax/mcmc/hmc.py:279-312 (kernel), integrators.py:104-150
 * (velocity Verlet), metrics.py:260-270 (momentum draw, kinetic energy), proposal.py:214-235
 * (Metropolis accept) and the jax.random threefry2x32 PRNG, float32, chains split across pthreads --
 * the per-chain program that jax.vmap batches on CPU.
 *
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=fast -pthread -shared -fPIC oracle_hmc.c -o liboracle_hmc.so -lm
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

static inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static void threefry2x32(uint32_t k0, uint32_t k1, uint32_t x0, uint32_t x1, uint32_t* o0, uint32_t* o1) {
  static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  x0 += ks[0];
  x1 += ks[1];
  for (int i = 0; i < 5; ++i) {
    for (int j = 0; j < 4; ++j) {
      x0 += x1;
      x1 = rotl(x1, R[i & 1][j]);
      x1 ^= x0;
    }
    x0 += ks[(i + 1) % 3];
    x1 += ks[(i + 2) % 3] + (uint32_t)(i + 1);
  }
  *o0 = x0;
  *o1 = x1;
}

static inline float bits_to_unit(uint32_t b) {
  union { uint32_t u; float f; } c;
  c.u = (b >> 9) | 0x3F800000u;
  return c.f - 1.0f;
}

static float erfinv_f32(float x) {
  static const float A[9] = {2.81022636e-08f, 3.43273939e-07f, -3.5233877e-06f, -4.39150654e-06f, 0.00021858087f,
                             -0.00125372503f, -0.00417768164f, 0.246640727f, 1.50140941f};
  static const float B[9] = {-0.000200214257f, 0.000100950558f, 0.00134934322f, -0.00367342844f, 0.00573950773f,
                             -0.0076224613f, 0.00943887047f, 1.00167406f, 2.83297682f};
  float w = -log1pf(-(x * x));
  const float* c = (w < 5.0f) ? A : B;
  float ww = (w < 5.0f) ? (w - 2.5f) : (sqrtf(w) - 3.0f);
  float p = c[0];
  for (int i = 1; i < 9; ++i) p = c[i] + p * ww;
  if (fabsf(x) == 1.0f) return x * INFINITY;
  return p * x;
}

static inline float normal_at(uint32_t k0, uint32_t k1, uint32_t i) {
  uint32_t a, b;
  threefry2x32(k0, k1, 0u, i, &a, &b);
  const float lo = -0.99999994f;
  float u = bits_to_unit(a ^ b) * (1.0f - lo) + lo;
  if (u < lo) u = lo;
  return 1.41421356237309515f * erfinv_f32(u);
}

/* target kinds: 0 diagonal Gaussian (inv_var[D]), 1 Neal's funnel, 2 hierarchical logistic regression (BASELINE config 5:
 * x = [mu, log_tau, beta0, beta1, alpha_0 .. alpha_{G-1}], data_x [G,8,2], data_y [G] outcome bits; oracle/targets.py
 * HierLogit, with the library's exact expf / log1pf as a CPU implementation would use them) */
static float value_and_grad(int kind, int D, const float* inv_var, const float* data_x, const unsigned char* data_y,
                            const float* q, float* g) {
  if (kind == 0) {
    float acc = 0.f;
#pragma omp simd reduction(+ : acc)
    for (int i = 0; i < D; ++i) {
      float t = q[i] * inv_var[i];
      acc += q[i] * t;
      g[i] = -t;
    }
    return -0.5f * acc;
  }
  if (kind == 2) {
    const int G = D - 4;
    const float mu = q[0], lt = q[1], b0 = q[2], b1 = q[3];
    const float e2 = expf(-2.0f * lt);
    float ll = 0.f, sd = 0.f, sd2 = 0.f, gb0 = 0.f, gb1 = 0.f;
    for (int gi = 0; gi < G; ++gi) {
      const float alpha = q[4 + gi], d = alpha - mu;
      const float* xr = data_x + (size_t)gi * 16;
      const unsigned bits = data_y[gi];
      float ga = 0.f;
      for (int k = 0; k < 8; ++k) {
        const float x0 = xr[2 * k], x1 = xr[2 * k + 1];
        const float y = (float)((bits >> k) & 1u);
        const float eta = alpha + b0 * x0 + b1 * x1;
        const float sig = 1.0f / (1.0f + expf(-eta));
        const float softplus = fmaxf(eta, 0.f) + log1pf(expf(-fabsf(eta)));
        const float r = y - sig;
        ll += y * eta - softplus;
        ga += r;
        gb0 += r * x0;
        gb1 += r * x1;
      }
      sd += d;
      sd2 += d * d;
      g[4 + gi] = -d * e2 + ga;
    }
    g[0] = -0.01f * mu + e2 * sd;
    g[1] = -lt + e2 * sd2 - (float)G;
    g[2] = -0.16f * b0 + gb0;
    g[3] = -0.16f * b1 + gb1;
    return -0.005f * mu * mu - 0.5f * lt * lt - 0.08f * (b0 * b0 + b1 * b1) + (-0.5f * e2 * sd2 - (float)G * lt) + ll;
  }
  float y = q[0], ss = 0.f;
#pragma omp simd reduction(+ : ss)
  for (int i = 1; i < D; ++i) ss += q[i] * q[i];
  float ey = expf(-y), n = (float)(D - 1), t = y / 3.0f;
  for (int i = 1; i < D; ++i) g[i] = -(ey * q[i]);
  g[0] = -y / 9.0f + 0.5f * ey * ss - 0.5f * n;
  return -0.5f * (t * t) + (-0.5f * ey * ss - 0.5f * n * y);
}

typedef struct {
  int c0, c1, D, kind, L;
  const float *inv_var, *imm, *msqrt;
  const uint32_t* keys;
  float *q, *logp, *g, eps, *acc_rate;
  unsigned char* accepted;
  const float* data_x;
  const unsigned char* data_y;
} job_t;

static void* worker(void* arg) {
  job_t* J = (job_t*)arg;
  const int D = J->D, L = J->L;
  const float *imm = J->imm, *msqrt = J->msqrt, eps = J->eps;
  float* p = (float*)malloc(sizeof(float) * D * 3);
  float* q1 = p + D;
  float* g1 = q1 + D;
  for (int c = J->c0; c < J->c1; ++c) {
    uint32_t km0, km1, ki0, ki1;
    threefry2x32(J->keys[2 * c], J->keys[2 * c + 1], 0u, 0u, &km0, &km1); /* split(rng_key, 2) hmc.py:299 */
    threefry2x32(J->keys[2 * c], J->keys[2 * c + 1], 0u, 1u, &ki0, &ki1);
    float* qc = J->q + (size_t)c * D;
    float* gc = J->g + (size_t)c * D;
    float k0 = 0.f;
    for (int i = 0; i < D; ++i) {
      p[i] = msqrt[i] * normal_at(km0, km1, (uint32_t)i);
      k0 += (imm[i] * p[i]) * p[i];
    }
    const float e0 = -J->logp[c] + 0.5f * k0;
    memcpy(q1, qc, sizeof(float) * D);
    memcpy(g1, gc, sizeof(float) * D);
    float lp = J->logp[c];
    const float eh = eps * 0.5f, e1s = eps * 1.0f;
    for (int s = 0; s < L; ++s) { /* integrators.py:104-150 */
      for (int i = 0; i < D; ++i) {
        p[i] = p[i] + eh * g1[i];
        q1[i] = q1[i] + e1s * (imm[i] * p[i]);
      }
      lp = value_and_grad(J->kind, D, J->inv_var, J->data_x, J->data_y, q1, g1);
      for (int i = 0; i < D; ++i) p[i] = p[i] + eh * g1[i];
    }
    float k1 = 0.f;
#pragma omp simd reduction(+ : k1)
    for (int i = 0; i < D; ++i) {
      float pf = -1.0f * p[i];
      k1 += (imm[i] * pf) * pf;
    }
    const float e1 = -lp + 0.5f * k1;
    float delta = e0 - e1;
    if (isnan(delta)) delta = -INFINITY;
    float pa = expf(delta);
    if (pa > 1.0f) pa = 1.0f;
    uint32_t a, b;
    threefry2x32(ki0, ki1, 0u, 0u, &a, &b);
    const float u = bits_to_unit(a ^ b);
    const int acc = u < pa;
    if (acc) {
      memcpy(qc, q1, sizeof(float) * D);
      memcpy(gc, g1, sizeof(float) * D);
      J->logp[c] = lp;
    }
    if (J->acc_rate) J->acc_rate[c] = pa;
    if (J->accepted) J->accepted[c] = (unsigned char)acc;
  }
  free(p);
  return NULL;
}

/* CPUs this process may actually use: the scheduler affinity mask, capped by the cgroup CPU quota (a container that
 * reports 128 online CPUs but is throttled to a few cores' worth of time runs faster with that many threads). */
static int affinity_cpus(int* ids, int cap) {
  cpu_set_t set;
  CPU_ZERO(&set);
  int n = 0;
  if (sched_getaffinity(0, sizeof(set), &set) == 0)
    for (int c = 0; c < CPU_SETSIZE && n < cap; ++c)
      if (CPU_ISSET(c, &set)) ids[n++] = c;
  return n;
}
static double cgroup_cpu_quota(void) {
  FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r");
  double q = 0.0;
  if (f) {
    char a[64];
    long long period = 0;
    if (fscanf(f, "%63s %lld", a, &period) == 2 && strcmp(a, "max") != 0 && period > 0) q = atof(a) / (double)period;
    fclose(f);
  } else { /* cgroup v1 */
    FILE* fq = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r");
    FILE* fp = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r");
    long long quota = -1, period = 0;
    if (fq && fp && fscanf(fq, "%lld", &quota) == 1 && fscanf(fp, "%lld", &period) == 1 && quota > 0 && period > 0)
      q = (double)quota / (double)period;
    if (fq) fclose(fq);
    if (fp) fclose(fp);
  }
  return q; /* 0: unlimited */
}
int oracle_num_threads(void) {
  int ids[CPU_SETSIZE];
  int n = affinity_cpus(ids, CPU_SETSIZE);
  if (n <= 0) {
    long m = sysconf(_SC_NPROCESSORS_ONLN);
    n = m > 0 ? (int)m : 1;
  }
  const double q = cgroup_cpu_quota();
  if (q >= 1.0 && q < (double)n) n = (int)q;
  return n > 0 ? n : 1;
}
/* pin worker t to the t-th CPU of the affinity mask (no migration between the timed repeats) */
static void pin_thread(pthread_t th, int t) {
  int ids[CPU_SETSIZE];
  const int n = affinity_cpus(ids, CPU_SETSIZE);
  if (n <= 0) return;
  cpu_set_t one;
  CPU_ZERO(&one);
  CPU_SET(ids[t % n], &one);
  pthread_setaffinity_np(th, sizeof(one), &one);
}

static long long hmc_step_impl(int C, int D, int kind, const float* inv_var, const float* data_x, const unsigned char* data_y,
                               const float* imm, const uint32_t* keys, float* q, float* logp, float* g, float eps, int L,
                               float* acc_rate, unsigned char* accepted, int n_threads) {
  float* msqrt = (float*)malloc(sizeof(float) * D);
  for (int i = 0; i < D; ++i) msqrt[i] = 1.0f / sqrtf(imm[i]);
  int T = n_threads > 0 ? n_threads : oracle_num_threads();
  if (T > C) T = C;
  if (T < 1) T = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * T);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * T);
  for (int t = 0; t < T; ++t) {
    job_t j = {(int)((long long)C * t / T), (int)((long long)C * (t + 1) / T), D, kind, L, inv_var, imm, msqrt, keys,
               q, logp, g, eps, acc_rate, accepted, data_x, data_y};
    jobs[t] = j;
    pthread_create(&th[t], NULL, worker, &jobs[t]);
    pin_thread(th[t], t);
  }
  for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
  free(msqrt);
  return (long long)C * L;
}

/* One HMC transition for chains [0, C): in-place on q, logp, g.  Returns the number of leapfrogs done. */
long long oracle_hmc_step(int C, int D, int kind, const float* inv_var, const float* imm /*[D]*/,
                          const uint32_t* keys /*[C,2]*/, float* q, float* logp, float* g, float eps, int L,
                          float* acc_rate /*[C] or NULL*/, unsigned char* accepted /*[C] or NULL*/, int n_threads) {
  return hmc_step_impl(C, D, kind, inv_var, NULL, NULL, imm, keys, q, logp, g, eps, L, acc_rate, accepted, n_threads);
}

/* The same transition on the hierarchical logistic regression of BASELINE config 5 (D = 4 + G). */
long long oracle_hmc_step_hier(int C, int D, const float* data_x /*[G,8,2]*/, const unsigned char* data_y /*[G]*/,
                               const float* imm, const uint32_t* keys, float* q, float* logp, float* g, float eps, int L,
                               float* acc_rate, unsigned char* accepted, int n_threads) {
  return hmc_step_impl(C, D, 2, NULL, data_x, data_y, imm, keys, q, logp, g, eps, L, acc_rate, accepted, n_threads);
}

/* ------------------------------------------------------------------------------------------------------
 * Dense variant (BASELINE config 2): dense Gaussian target logp = -1/2 x^T P x, dense inverse mass matrix.
 * Chains are processed in blocks of BLK through a register-blocked GEMM micro-kernel (the matrices are symmetric;
 * L^-T is transposed once per call).
 * ------------------------------------------------------------------------------------------------------ */
#define BLK 48
typedef struct {
  int c0, c1, D, L;
  const float *prec, *imm, *msqrt_t;
  const uint32_t* keys;
  float *q, *logp, *g, eps, *acc_rate;
  unsigned char* accepted;
} djob_t;

/* Y[b][i] = sum_j X[b][j] At[j][i]   (At row-major [D,D] = A^T; the symmetric matrices are passed as they are).
 * A register-blocked GEMM micro-kernel: 6 chains x 16 outputs of accumulators, broadcast X, stream rows of At --
 * what a compiled "vmap of linear_map" amounts to on a CPU (XLA:CPU hands the batched product to an Eigen GEMM). */
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
static void matmul_blk(int nb, int D, const float* At, const float* X, float* Y) {
  const int D16 = D & ~15;
  for (int i0 = 0; i0 < D16; i0 += 16) {
    int b0 = 0;
    for (; b0 + 6 <= nb; b0 += 6) {
      __m256 c00 = _mm256_setzero_ps(), c01 = c00, c10 = c00, c11 = c00, c20 = c00, c21 = c00, c30 = c00, c31 = c00,
             c40 = c00, c41 = c00, c50 = c00, c51 = c00;
      const float* x0 = X + (size_t)b0 * D;
      for (int j = 0; j < D; ++j) {
        const __m256 a0 = _mm256_loadu_ps(At + (size_t)j * D + i0), a1 = _mm256_loadu_ps(At + (size_t)j * D + i0 + 8);
        __m256 x;
        x = _mm256_broadcast_ss(x0 + j);                 c00 = _mm256_fmadd_ps(x, a0, c00); c01 = _mm256_fmadd_ps(x, a1, c01);
        x = _mm256_broadcast_ss(x0 + (size_t)D + j);     c10 = _mm256_fmadd_ps(x, a0, c10); c11 = _mm256_fmadd_ps(x, a1, c11);
        x = _mm256_broadcast_ss(x0 + (size_t)2 * D + j); c20 = _mm256_fmadd_ps(x, a0, c20); c21 = _mm256_fmadd_ps(x, a1, c21);
        x = _mm256_broadcast_ss(x0 + (size_t)3 * D + j); c30 = _mm256_fmadd_ps(x, a0, c30); c31 = _mm256_fmadd_ps(x, a1, c31);
        x = _mm256_broadcast_ss(x0 + (size_t)4 * D + j); c40 = _mm256_fmadd_ps(x, a0, c40); c41 = _mm256_fmadd_ps(x, a1, c41);
        x = _mm256_broadcast_ss(x0 + (size_t)5 * D + j); c50 = _mm256_fmadd_ps(x, a0, c50); c51 = _mm256_fmadd_ps(x, a1, c51);
      }
      float* y0 = Y + (size_t)b0 * D + i0;
      _mm256_storeu_ps(y0, c00);                 _mm256_storeu_ps(y0 + 8, c01);
      _mm256_storeu_ps(y0 + (size_t)D, c10);     _mm256_storeu_ps(y0 + (size_t)D + 8, c11);
      _mm256_storeu_ps(y0 + (size_t)2 * D, c20); _mm256_storeu_ps(y0 + (size_t)2 * D + 8, c21);
      _mm256_storeu_ps(y0 + (size_t)3 * D, c30); _mm256_storeu_ps(y0 + (size_t)3 * D + 8, c31);
      _mm256_storeu_ps(y0 + (size_t)4 * D, c40); _mm256_storeu_ps(y0 + (size_t)4 * D + 8, c41);
      _mm256_storeu_ps(y0 + (size_t)5 * D, c50); _mm256_storeu_ps(y0 + (size_t)5 * D + 8, c51);
    }
    for (; b0 < nb; ++b0) { /* leftover chains of the block */
      __m256 c0 = _mm256_setzero_ps(), c1 = c0;
      const float* x0 = X + (size_t)b0 * D;
      for (int j = 0; j < D; ++j) {
        const __m256 x = _mm256_broadcast_ss(x0 + j);
        c0 = _mm256_fmadd_ps(x, _mm256_loadu_ps(At + (size_t)j * D + i0), c0);
        c1 = _mm256_fmadd_ps(x, _mm256_loadu_ps(At + (size_t)j * D + i0 + 8), c1);
      }
      _mm256_storeu_ps(Y + (size_t)b0 * D + i0, c0);
      _mm256_storeu_ps(Y + (size_t)b0 * D + i0 + 8, c1);
    }
  }
  for (int i = D16; i < D; ++i) /* leftover outputs */
    for (int b = 0; b < nb; ++b) {
      float acc = 0.f;
      for (int j = 0; j < D; ++j) acc += X[(size_t)b * D + j] * At[(size_t)j * D + i];
      Y[(size_t)b * D + i] = acc;
    }
}
#else
static void matmul_blk(int nb, int D, const float* At, const float* X, float* Y) {
  for (int b = 0; b < nb; ++b) {
    float* y = Y + (size_t)b * D;
    for (int i = 0; i < D; ++i) y[i] = 0.f;
    for (int j = 0; j < D; ++j) {
      const float x = X[(size_t)b * D + j];
      const float* row = At + (size_t)j * D;
#pragma omp simd
      for (int i = 0; i < D; ++i) y[i] += x * row[i];
    }
  }
}
#endif

static void* dworker(void* arg) {
  djob_t* J = (djob_t*)arg;
  const int D = J->D, L = J->L;
  const float eps = J->eps, eh = eps * 0.5f, e1s = eps * 1.0f;
  float* buf = (float*)malloc(sizeof(float) * (size_t)BLK * D * 5);
  float *p = buf, *v = p + (size_t)BLK * D, *q1 = v + (size_t)BLK * D, *g1 = q1 + (size_t)BLK * D, *z = g1 + (size_t)BLK * D;
  for (int c0 = J->c0; c0 < J->c1; c0 += BLK) {
    const int nb = (c0 + BLK <= J->c1) ? BLK : (J->c1 - c0);
    uint32_t ki0[BLK], ki1[BLK];
    float e0[BLK], lp[BLK];
    for (int b = 0; b < nb; ++b) {
      const int c = c0 + b;
      uint32_t km0, km1;
      threefry2x32(J->keys[2 * c], J->keys[2 * c + 1], 0u, 0u, &km0, &km1);
      threefry2x32(J->keys[2 * c], J->keys[2 * c + 1], 0u, 1u, &ki0[b], &ki1[b]);
      for (int i = 0; i < D; ++i) z[(size_t)b * D + i] = normal_at(km0, km1, (uint32_t)i);
      memcpy(q1 + (size_t)b * D, J->q + (size_t)c * D, sizeof(float) * D);
      memcpy(g1 + (size_t)b * D, J->g + (size_t)c * D, sizeof(float) * D);
      lp[b] = J->logp[c];
    }
    matmul_blk(nb, D, J->msqrt_t, z, p); /* p = L^-T z */
    matmul_blk(nb, D, J->imm, p, v);
    for (int b = 0; b < nb; ++b) {
      float k = 0.f;
      for (int i = 0; i < D; ++i) k += v[(size_t)b * D + i] * p[(size_t)b * D + i];
      e0[b] = -lp[b] + 0.5f * k;
    }
    for (int s = 0; s < L; ++s) {
      for (size_t t = 0; t < (size_t)nb * D; ++t) p[t] = p[t] + eh * g1[t];
      matmul_blk(nb, D, J->imm, p, v);
      for (size_t t = 0; t < (size_t)nb * D; ++t) q1[t] = q1[t] + e1s * v[t];
      matmul_blk(nb, D, J->prec, q1, v); /* P q */
      for (int b = 0; b < nb; ++b) {
        float acc = 0.f;
        for (int i = 0; i < D; ++i) {
          const size_t t = (size_t)b * D + i;
          g1[t] = -v[t];
          acc += q1[t] * g1[t];
          p[t] = p[t] + eh * g1[t];
        }
        lp[b] = 0.5f * acc;
      }
    }
    matmul_blk(nb, D, J->imm, p, v);
    for (int b = 0; b < nb; ++b) {
      const int c = c0 + b;
      float k = 0.f;
      for (int i = 0; i < D; ++i) k += v[(size_t)b * D + i] * p[(size_t)b * D + i];
      const float e1 = -lp[b] + 0.5f * k;
      float delta = e0[b] - e1;
      if (isnan(delta)) delta = -INFINITY;
      float pa = expf(delta);
      if (pa > 1.0f) pa = 1.0f;
      uint32_t a, bb;
      threefry2x32(ki0[b], ki1[b], 0u, 0u, &a, &bb);
      const int acc = bits_to_unit(a ^ bb) < pa;
      if (acc) {
        memcpy(J->q + (size_t)c * D, q1 + (size_t)b * D, sizeof(float) * D);
        memcpy(J->g + (size_t)c * D, g1 + (size_t)b * D, sizeof(float) * D);
        J->logp[c] = lp[b];
      }
      if (J->acc_rate) J->acc_rate[c] = pa;
      if (J->accepted) J->accepted[c] = (unsigned char)acc;
    }
  }
  free(buf);
  return NULL;
}

long long oracle_hmc_dense_step(int C, int D, const float* prec, const float* imm, const float* msqrt,
                                const uint32_t* keys, float* q, float* logp, float* g, float eps, int L, float* acc_rate,
                                unsigned char* accepted, int n_threads) {
  int T = n_threads > 0 ? n_threads : oracle_num_threads();
  int nblk = (C + BLK - 1) / BLK;
  if (T > nblk) T = nblk;
  if (T < 1) T = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * T);
  djob_t* jobs = (djob_t*)malloc(sizeof(djob_t) * T);
  float* msqrt_t = (float*)malloc(sizeof(float) * (size_t)D * D); /* (L^-T)^T, so that p_i = sum_j z_j msqrt_t[j][i] */
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) msqrt_t[(size_t)j * D + i] = msqrt[(size_t)i * D + j];
  for (int t = 0; t < T; ++t) {
    int b0 = (int)((long long)nblk * t / T), b1 = (int)((long long)nblk * (t + 1) / T);
    djob_t j = {b0 * BLK, b1 * BLK < C ? b1 * BLK : C, D, L, prec, imm, msqrt_t, keys, q, logp, g, eps, acc_rate, accepted};
    jobs[t] = j;
    pthread_create(&th[t], NULL, dworker, &jobs[t]);
    pin_thread(th[t], t);
  }
  for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
  free(msqrt_t);
  return (long long)C * L;
}
