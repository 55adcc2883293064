// Shared (cross-chain, cross-GPU) window adaptation, device resident, and the collective of the path.
//
// Reference semantics (blackjax/adaptation/staged_adaptation.py:153-171,906-966): ONE dual-averaging update per
// warm-up step fed with mean(acceptance_rate over ALL chains), and in slow windows the chain-pooled moment block
// (metric_buffers.py:396-420) merged into the window accumulator by Chan-Golub-LeVeque (metric_buffers.py:334-393);
// window end: regularised inverse mass matrix (mass_matrix.py:335-357), accumulator reset, dual averaging
// re-initialised at exp(log_step_avg) (staged_adaptation.py:233-249).  The reference's collective for the cross-device
// case is lax.psum over the "chains" axis (blackjax/eca.py:56-62); here it is ONE NCCL all-gather of the per-GPU
// summary blocks per warm-up step, followed by the same merge on every rank (no broadcast needed).
//
// Results do not depend on the GPU count: the chains of a rank are reduced in FIXED blocks of BJX_STAT_BLOCK_CHAINS
// chains (mean and centred second moment per block, deterministic tree), the blocks of all ranks are gathered in global
// chain order and merged sequentially, so 1 x 65536 chains and 2 x 32768 chains execute the same float operations.
// Everything (merge, dual averaging, mass-matrix finalisation, the per-chain step-size array the transition kernels read)
// stays on the device: no .item(), no host round trip per warm-up step.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>

#include "bjx_handle.h"
#include "bjx_internal.h"

namespace bjx {

constexpr int kStatBlock = BJX_STAT_BLOCK_CHAINS;
constexpr int kHdr = 16;  // scalar slots at the head of the adaptation state

// ---- per-block chain-pooled statistics ------------------------------------------------------------------------
// grid (ceil(D/32), n_blocks), block (32, 8): block (bx, by) reduces columns [32 bx, +32) over chains [by*B, (by+1)*B)
// to mean and sum of squared deviations (two passes, shared-memory tree in a fixed order); bx == 0 also sums the
// acceptance rates of those chains.  out[by] = (sum accept, n, mean[D], M2[D]).
__global__ void k_block_stats(int C, int D, const float* __restrict__ x, const float* __restrict__ acc, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int c0 = blockIdx.y * kStatBlock;
  const int c1 = min(C, c0 + kStatBlock);
  const int n = c1 - c0;
  const int col = blockIdx.x * 32 + threadIdx.x;
  float* o = out + (size_t)blockIdx.y * (2 + 2 * D);
  float a = 0.f;
  if (col < D)
    for (int c = c0 + threadIdx.y; c < c1; c += 8) a += x[(size_t)c * D + col];
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    red[0][threadIdx.x] = s / (float)n;
  }
  __syncthreads();
  const float mean = red[0][threadIdx.x];
  __syncthreads();
  a = 0.f;
  if (col < D)
    for (int c = c0 + threadIdx.y; c < c1; c += 8) {
      const float d = x[(size_t)c * D + col] - mean;
      a = fmaf(d, d, a);
    }
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && col < D) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    o[2 + col] = mean;
    o[2 + D + col] = s;
  }
  if (blockIdx.x == 0) {  // acceptance rates of the block's chains: 256 partial sums, fixed-order tree
    __syncthreads();
    const int t = threadIdx.y * 32 + threadIdx.x;
    float s = 0.f;
    for (int c = c0 + t; c < c1; c += 256) s += acc[c];
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (threadIdx.x == 0) red[threadIdx.y][0] = s;
    __syncthreads();
    if (t == 0) {
      float tot = 0.f;
      for (int k = 0; k < 8; ++k) tot += red[k][0];
      o[0] = tot;
      o[1] = (float)n;
    }
  }
}

// ---- merge + dual averaging + window bookkeeping: one CTA ---------------------------------------------------------
// state: [0] log_step [1] log_step_avg [2] step [3] avg_error [4] mu [5] eps [6] w_n [7] t; [16, 16+D) w_mean; [16+D, 16+2D) w_m2
struct MergeScalars {
  float n, acc;
};
__device__ __forceinline__ void cgl_factors(float na, float nb, float& f_mean, float& f_cross, float& n_ab) {
  // metric_buffers.py:334-393 with the reference's scalar arithmetic (exact integer counts, one rounding to float32)
  const double ab = (double)na + (double)nb;
  n_ab = (float)ab;
  f_mean = (float)((double)nb / ab);
  f_cross = (float)((double)na * (double)nb / ab);
}

__global__ void k_shared_adapt(int D, int nblk, const float* __restrict__ blocks, float* __restrict__ st, int stage, int window_end,
                               float target, float* __restrict__ imm, float* __restrict__ eps_hist) {
  const int S = 2 + 2 * D;
  // every thread replays the scalar recurrences (identical results), threads own dimensions
  float n = blocks[1], acc = blocks[0];
  const float w_n0 = st[6];
  float wn_new = w_n0, f_mean_w = 0.f, f_cross_w = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float nn = blocks[1];
    float mean = blocks[2 + d], m2 = blocks[2 + D + d];
    for (int b = 1; b < nblk; ++b) {
      const float* g = blocks + (size_t)b * S;
      float fm, fc, nab;
      cgl_factors(nn, g[1], fm, fc, nab);
      const float delta = g[2 + d] - mean;
      mean = mean + delta * fm;
      m2 = m2 + g[2 + D + d] + delta * delta * fc;
      nn = nab;
    }
    if (stage == 1) {  // merge this step's pooled block into the window accumulator (staged_adaptation.py:283-297)
      float fm, fc, nab;
      cgl_factors(w_n0, nn, fm, fc, nab);
      const float delta = mean - st[kHdr + d];
      const float wm = st[kHdr + d] + delta * fm;
      const float w2 = st[kHdr + D + d] + m2 + delta * delta * fc;
      if (window_end) {  // mass_matrix.py:335-357, then reset
        const float cov = w2 / (nab - 1.f);
        const float denom = nab + 5.f;
        imm[d] = nab / denom * cov + 5.f / denom * 1e-3f;
        st[kHdr + d] = 0.f;
        st[kHdr + D + d] = 0.f;
      } else {
        st[kHdr + d] = wm;
        st[kHdr + D + d] = w2;
      }
    }
  }
  for (int b = 1; b < nblk; ++b) {
    const float* g = blocks + (size_t)b * S;
    float fm, fc, nab;
    cgl_factors(n, g[1], fm, fc, nab);
    n = nab;
    acc += g[0];
  }
  if (stage == 1) {
    cgl_factors(w_n0, n, f_mean_w, f_cross_w, wn_new);
    if (window_end) wn_new = 0.f;
  }
  __syncthreads();  // every thread has read the scalar slots
  if (threadIdx.x == 0) {
    // dual averaging on the mean acceptance rate of all chains (dual_averaging.py:101-123; staged_adaptation.py:153-171)
    float log_step = st[0], avg_log_step = st[1], step = st[2], avg_error = st[3], mu = st[4];
    const float gradient = target - acc / n;
    const float reg_step = step + 10.f;
    const float eta_t = powf(step, -0.75f);
    avg_error = (1.f - (1.f / reg_step)) * avg_error + gradient / reg_step;
    const float log_x = mu - (sqrtf(step) / 0.05f) * avg_error;
    const float log_x_avg = eta_t * log_step + (1.f - eta_t) * avg_log_step;
    log_step = log_x; avg_log_step = log_x_avg; step = step + 1.f;
    if (window_end) {  // staged_adaptation.py:233-249
      const float x = expf(avg_log_step);
      log_step = logf(x); avg_log_step = 0.f; step = 1.f; avg_error = 0.f; mu = logf(10.f * x);
    }
    const float eps = expf(log_step);
    st[0] = log_step; st[1] = avg_log_step; st[2] = step; st[3] = avg_error; st[4] = mu; st[5] = eps;
    st[6] = wn_new;
    const int t = (int)st[7];
    if (eps_hist) eps_hist[t] = eps;
    st[7] = (float)(t + 1);
  }
}

__global__ void k_fill_from(int n, const float* __restrict__ src, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[0];
}
__global__ void k_adapt_init(int D, float* st, float eps0, float* imm) {
  for (int d = threadIdx.x; d < 2 * D; d += blockDim.x) st[kHdr + d] = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) imm[d] = 1.f;
  if (threadIdx.x == 0) {
    st[0] = logf(eps0); st[1] = 0.f; st[2] = 1.f; st[3] = 0.f; st[4] = logf(10.f * eps0); st[5] = expf(logf(eps0));
    st[6] = 0.f; st[7] = 0.f;
    for (int k = 8; k < kHdr; ++k) st[k] = 0.f;
  }
}

// ---- NCCL, bound at run time (libbjx.so itself has no NCCL dependency: single-GPU users never load it) -----------
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};
static NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // prefer a copy that is already in the process (PyTorch bundles one), else the system library
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) {
      api.err = std::string("cannot load libnccl.so.2: ") + dlerror();
      return;
    }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather) api.err = "libnccl.so.2 lacks the expected symbols";
  });
  return api;
}
}  // namespace bjx

using namespace bjx;

#define AD_CUDA(call)                                                \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) return bjx_cuda_fail(h, e_, #call);       \
  } while (0)
#define AD_LAUNCH(where)                                             \
  do {                                                               \
    cudaError_t e_ = cudaGetLastError();                             \
    if (e_ != cudaSuccess) return bjx_cuda_fail(h, e_, where);       \
  } while (0)

static int nccl_fail(bjx_handle_t h, ncclResult_t r, const char* where) {
  NcclApi& a = nccl();
  return bjx_fail(h, BJX_E_UNSUPPORTED, std::string(where) + ": " + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error"));
}

extern "C" int bjx_nccl_unique_id(void* id128_out) {
  NcclApi& a = nccl();
  if (!a.err.empty()) return bjx_fail(nullptr, BJX_E_UNSUPPORTED, a.err);
  if (!id128_out) return bjx_fail(nullptr, BJX_E_INVALID, "null argument");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclResult_t r = a.GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128_out));
  return r == ncclSuccess ? 0 : nccl_fail(nullptr, r, "ncclGetUniqueId");
}

extern "C" int bjx_nccl_comm_init_rank(const void* id128, int32_t n_ranks, int32_t rank, int32_t device, void** comm_out) {
  NcclApi& a = nccl();
  if (!a.err.empty()) return bjx_fail(nullptr, BJX_E_UNSUPPORTED, a.err);
  if (!id128 || !comm_out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return bjx_fail(nullptr, BJX_E_INVALID, "bad argument");
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return bjx_cuda_fail(nullptr, e, "cudaSetDevice");
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  ncclResult_t r = a.CommInitRank(&comm, n_ranks, id, rank);
  if (r != ncclSuccess) return nccl_fail(nullptr, r, "ncclCommInitRank");
  *comm_out = comm;
  return 0;
}

extern "C" int bjx_nccl_comm_destroy(void* comm) {
  NcclApi& a = nccl();
  if (!a.err.empty() || !comm) return 0;
  ncclResult_t r = a.CommDestroy(reinterpret_cast<ncclComm_t>(comm));
  return r == ncclSuccess ? 0 : nccl_fail(nullptr, r, "ncclCommDestroy");
}

extern "C" int bjx_allgather_stats(bjx_handle_t h, void* nccl_comm, const float* block, int64_t count, float* gathered) {
  if (!h || !block || !gathered || count <= 0) return bjx_fail(h, BJX_E_INVALID, "bad argument");
  AD_CUDA(cudaSetDevice(h->cfg.device));
  if (!nccl_comm) {  // one rank: the gathered array is the block
    if (gathered != block) AD_CUDA(cudaMemcpyAsync(gathered, block, (size_t)count * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    return 0;
  }
  NcclApi& a = nccl();
  if (!a.err.empty()) return bjx_fail(h, BJX_E_UNSUPPORTED, a.err);
  ncclResult_t r = a.AllGather(block, gathered, (size_t)count, ncclFloat, reinterpret_cast<ncclComm_t>(nccl_comm), h->stream);
  return r == ncclSuccess ? 0 : nccl_fail(h, r, "ncclAllGather");
}

static inline int n_stat_blocks(int C) { return (C + kStatBlock - 1) / kStatBlock; }

extern "C" int64_t bjx_adapt_shared_state_floats(int32_t n_chains_local, int32_t dim, int32_t n_ranks) {
  if (n_chains_local <= 0 || dim <= 0 || n_ranks <= 0) return 0;
  const int64_t blk = 2 + 2 * (int64_t)dim;
  return kHdr + 2 * (int64_t)dim + (int64_t)n_stat_blocks(n_chains_local) * blk * (1 + n_ranks);
}

extern "C" int bjx_adapt_shared_init(bjx_handle_t h, float* state, float initial_step_size, float* step_size_chain_out, float* imm_out) {
  if (!h || !state || !step_size_chain_out || !imm_out) return bjx_fail(h, BJX_E_INVALID, "null argument");
  if (!(initial_step_size > 0.f)) return bjx_fail(h, BJX_E_INVALID, "initial_step_size must be positive");
  AD_CUDA(cudaSetDevice(h->cfg.device));
  const int C = h->cfg.n_chains, D = h->cfg.dim;
  k_adapt_init<<<1, 256, 0, h->stream>>>(D, state, initial_step_size, imm_out);
  AD_LAUNCH("k_adapt_init");
  k_fill_from<<<(C + 255) / 256, 256, 0, h->stream>>>(C, state + 5, step_size_chain_out);
  AD_LAUNCH("k_fill_from");
  return bjx_set_metric(h, BJX_METRIC_DIAG, imm_out);
}

extern "C" int bjx_adapt_shared_update(bjx_handle_t h, void* nccl_comm, int32_t n_ranks, float* state, const float* q,
                                       const float* acceptance_rate, int32_t stage, int32_t window_end, float target_acceptance,
                                       float* step_size_chain, float* imm, float* eps_history) {
  if (!h || !state || !q || !acceptance_rate || !step_size_chain || !imm) return bjx_fail(h, BJX_E_INVALID, "null argument");
  if (n_ranks < 1 || (n_ranks > 1 && !nccl_comm)) return bjx_fail(h, BJX_E_INVALID, "n_ranks > 1 needs a communicator");
  if (window_end && stage != 1) return bjx_fail(h, BJX_E_INVALID, "a window ends on a slow-stage step");
  AD_CUDA(cudaSetDevice(h->cfg.device));
  const int C = h->cfg.n_chains, D = h->cfg.dim;
  const int nb = n_stat_blocks(C);
  const size_t blk = 2 + 2 * (size_t)D;
  float* local = state + kHdr + 2 * (size_t)D;
  float* gathered = local + (size_t)nb * blk;
  k_block_stats<<<dim3((D + 31) / 32, nb), dim3(32, 8), 0, h->stream>>>(C, D, q, acceptance_rate, local);
  AD_LAUNCH("k_block_stats");
  const float* merged_in = local;
  if (n_ranks > 1) {
    int rc = bjx_allgather_stats(h, nccl_comm, local, (int64_t)(nb * blk), gathered);
    if (rc) return rc;
    merged_in = gathered;
  }
  k_shared_adapt<<<1, 512, 0, h->stream>>>(D, nb * n_ranks, merged_in, state, stage, window_end, target_acceptance, imm, eps_history);
  AD_LAUNCH("k_shared_adapt");
  k_fill_from<<<(C + 255) / 256, 256, 0, h->stream>>>(C, state + 5, step_size_chain);
  AD_LAUNCH("k_fill_from");
  if (window_end) return bjx_set_metric(h, BJX_METRIC_DIAG, imm);  // mass_matrix_sqrt of the new metric, on the stream
  return 0;
}

extern "C" int bjx_adapt_shared_final(bjx_handle_t h, const float* state, float* step_size_out) {
  if (!h || !state || !step_size_out) return bjx_fail(h, BJX_E_INVALID, "null argument");
  AD_CUDA(cudaSetDevice(h->cfg.device));
  launch_da(3, 1, const_cast<float*>(state), nullptr, 0.f, step_size_out, h->stream);  // exp(log_step_avg)
  AD_LAUNCH("k_da_final");
  return 0;
}

// =====================================================================================================================
// ChEES-HMC warm-up (blackjax/adaptation/chees_adaptation.py, mass_matrix_estimation=None): dual averaging on the
// harmonic mean of the acceptance probabilities (:341-360) and Adam ascent on log(trajectory length) along the ChEES
// gradient jitter * T * (|dx'|^2 - |dx|^2) <dx', p'> (:362-480), both from cross-chain statistics.  Same structure as
// the shared window adaptation above: fixed blocks of BJX_STAT_BLOCK_CHAINS chains, merged in global chain order after
// an all-gather (two per step: the centring means must be global before the per-chain dot products can be formed).
// =====================================================================================================================
namespace bjx {
constexpr int kChHdr = 32;
// header slots
enum { CH_EPS = 0, CH_LOG_EPS_MA, CH_T, CH_LOG_T_MA, CH_DA_LOGX, CH_DA_LOGX_AVG, CH_DA_STEP, CH_DA_ERR, CH_DA_MU, CH_AD_COUNT,
       CH_AD_MU, CH_AD_NU, CH_RGA, CH_STEP, CH_MAX_BITS, CH_JITTER_AMOUNT, CH_T_IDX, CH_L, CH_HM_SUM, CH_HM_N, CH_GRAD };

__device__ __forceinline__ float halton_f(int i, int max_bits) {  // dynamic_hmc.py:205-215
  float s = 0.f;
  for (int k = 0; k < max_bits; ++k) s += (float)(((i + 1) >> k) & 1) * (0.5f / (float)(1u << k));
  return s;
}

// per chain: w' = is_divergent ? 0 : acceptance, zeroed when the proposal row holds a non-finite entry (:239-247)
__global__ void k_chees_w(int C, int D, const float* __restrict__ prop_q, const float* __restrict__ acc,
                          const uint8_t* __restrict__ is_div, float* __restrict__ w) {
  const int lane = threadIdx.x & 31, c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  bool fin = true;
  for (int i = lane; i < D; i += 32) fin = fin && isfinite(prop_q[(size_t)c * D + i]);
  fin = __all_sync(0xffffffffu, fin);
  if (lane == 0) w[c] = (!is_div[c] && fin) ? acc[c] : 0.f;
}

// block (bx, by): columns [32 bx, +32) over chains of statistic block by.
// out[by] = (sum_{nd} 1/acc, n_nd, sum w', sum_c w'_c x_c [D], sum_c finite q_c [D], count of non-NaN q [D])
__global__ void k_chees_pass1(int C, int D, const float* __restrict__ prop_q, const float* __restrict__ init_q,
                              const float* __restrict__ acc, const uint8_t* __restrict__ is_div, const float* __restrict__ w,
                              float* __restrict__ out) {
  __shared__ float red[3][8][33];
  const int c0 = blockIdx.y * kStatBlock, c1 = min(C, c0 + kStatBlock);
  const int col = blockIdx.x * 32 + threadIdx.x;
  float* o = out + (size_t)blockIdx.y * (3 + 3 * (size_t)D);
  float sw = 0.f, sq = 0.f, cq = 0.f;
  if (col < D)
    for (int c = c0 + threadIdx.y; c < c1; c += 8) {
      const float x = prop_q[(size_t)c * D + col], q = init_q[(size_t)c * D + col];
      sw = fmaf(w[c], isfinite(x) ? x : 0.f, sw);
      if (!isnan(q)) { sq += q; cq += 1.f; }
    }
  red[0][threadIdx.y][threadIdx.x] = sw;
  red[1][threadIdx.y][threadIdx.x] = sq;
  red[2][threadIdx.y][threadIdx.x] = cq;
  __syncthreads();
  if (threadIdx.y == 0 && col < D) {
    float a = 0.f, b = 0.f, n = 0.f;
    for (int k = 0; k < 8; ++k) { a += red[0][k][threadIdx.x]; b += red[1][k][threadIdx.x]; n += red[2][k][threadIdx.x]; }
    o[3 + col] = a;
    o[3 + D + col] = b;
    o[3 + 2 * D + col] = n;
  }
  if (blockIdx.x == 0) {
    __syncthreads();
    const int t = threadIdx.y * 32 + threadIdx.x;
    float h = 0.f, n = 0.f, ws = 0.f;
    for (int c = c0 + t; c < c1; c += 256) {
      if (!is_div[c]) { h += 1.0f / acc[c]; n += 1.f; }
      ws += w[c];
    }
    for (int off = 16; off > 0; off >>= 1) {
      h += __shfl_xor_sync(0xffffffffu, h, off);
      n += __shfl_xor_sync(0xffffffffu, n, off);
      ws += __shfl_xor_sync(0xffffffffu, ws, off);
    }
    if (threadIdx.x == 0) { red[0][threadIdx.y][0] = h; red[1][threadIdx.y][0] = n; red[2][threadIdx.y][0] = ws; }
    __syncthreads();
    if (t == 0) {
      float a = 0.f, b = 0.f, c_ = 0.f;
      for (int k = 0; k < 8; ++k) { a += red[0][k][0]; b += red[1][k][0]; c_ += red[2][k][0]; }
      o[0] = a; o[1] = b; o[2] = c_;
    }
  }
}

// merge the pass-1 blocks in global order: proposal mean (weighted), initial mean (nanmean), harmonic-mean sums
__global__ void k_chees_means(int D, int nblk, const float* __restrict__ blocks, float* __restrict__ st) {
  const size_t S = 3 + 3 * (size_t)D;
  float* pm = st + kChHdr;
  float* qm = pm + D;
  float sw = 0.f;
  for (int b = 0; b < nblk; ++b) sw += blocks[b * S + 2];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a = 0.f, q = 0.f, n = 0.f;
    for (int b = 0; b < nblk; ++b) {
      a += blocks[b * S + 3 + d];
      q += blocks[b * S + 3 + D + d];
      n += blocks[b * S + 3 + 2 * D + d];
    }
    pm[d] = a / (sw + 1e-20f);
    qm[d] = q / n;
  }
  if (threadIdx.x == 0) {
    float h = 0.f, n = 0.f;
    for (int b = 0; b < nblk; ++b) { h += blocks[b * S]; n += blocks[b * S + 1]; }
    st[CH_HM_SUM] = h;
    st[CH_HM_N] = n;
  }
}

// per chain (one warp): (|x' - E x'|^2 - |x - E x|^2) <x' - E x', p'>, then its acceptance-weighted terms
__global__ void k_chees_dots(int C, int D, const float* __restrict__ prop_q, const float* __restrict__ prop_p,
                             const float* __restrict__ init_q, const float* __restrict__ acc, const uint8_t* __restrict__ is_div,
                             const float* __restrict__ st, float* __restrict__ num, float* __restrict__ den) {
  const int lane = threadIdx.x & 31, c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const float* pm = st + kChHdr;
  const float* qm = pm + D;
  float a = 0.f, b = 0.f, d = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float pc = prop_q[(size_t)c * D + i] - pm[i];
    const float ic = init_q[(size_t)c * D + i] - qm[i];
    a = fmaf(pc, pc, a);
    b = fmaf(ic, ic, b);
    d = fmaf(pc, prop_p[(size_t)c * D + i], d);
  }
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
    d += __shfl_xor_sync(0xffffffffu, d, off);
  }
  if (lane == 0) {
    const bool nd = !is_div[c];
    num[c] = nd ? acc[c] * ((a - b) * d) : 0.f;
    den[c] = nd ? acc[c] + 1e-20f : 0.f;
  }
}
// fixed-order sums of the two per-chain arrays over each statistic block -> out[by] = (sum num, sum den)
__global__ void k_chees_pass2(int C, const float* __restrict__ num, const float* __restrict__ den, float* __restrict__ out) {
  __shared__ float red[2][8];
  const int c0 = blockIdx.x * kStatBlock, c1 = min(C, c0 + kStatBlock);
  float a = 0.f, b = 0.f;
  for (int c = c0 + threadIdx.x; c < c1; c += 256) { a += num[c]; b += den[c]; }
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = 0.f, y = 0.f;
    for (int k = 0; k < 8; ++k) { x += red[0][k]; y += red[1][k]; }
    out[2 * blockIdx.x] = x;
    out[2 * blockIdx.x + 1] = y;
  }
}

// the scalar update (chees_adaptation.py:341-360,470-511) and the next transition's step count
__global__ void k_chees_finish(int nblk, const float* __restrict__ blocks2, float* __restrict__ st, float lr, float b1, float b2,
                               float target, float decay, float max_leapfrog, float* __restrict__ hist) {
  if (threadIdx.x != 0) return;
  const int max_bits = (int)st[CH_MAX_BITS];
  const float jam = st[CH_JITTER_AMOUNT];
  const int rga = (int)st[CH_RGA];
  const float step = st[CH_STEP];
  float eps = st[CH_EPS], T = st[CH_T];
  // step size: dual averaging on target - harmonic mean of the non-divergent acceptance probabilities
  float hm = 1.0f / (st[CH_HM_SUM] / st[CH_HM_N]);
  if (!isfinite(hm)) hm = 0.f;
  float log_x = st[CH_DA_LOGX], log_x_avg = st[CH_DA_LOGX_AVG], da_step = st[CH_DA_STEP], err = st[CH_DA_ERR];
  const float mu = st[CH_DA_MU];
  {
    const float gradient = target - hm;
    const float reg_step = da_step + 10.f;
    const float eta_t = powf(da_step, -0.75f);
    const float nerr = (1.f - (1.f / reg_step)) * err + gradient / reg_step;
    const float nlog_x = mu - (sqrtf(da_step) / 0.05f) * nerr;
    const float navg = eta_t * log_x + (1.f - eta_t) * log_x_avg;
    const float neps = expf(nlog_x);
    if (isfinite(neps)) { eps = neps; log_x = nlog_x; log_x_avg = navg; da_step += 1.f; err = nerr; }
  }
  const float uw = powf(step, -decay);
  const float log_eps_ma = (1.f - uw) * st[CH_LOG_EPS_MA] + uw * log_x;
  // trajectory length: Adam on log T along the ChEES gradient
  float num = 0.f, den = 0.f;
  for (int b = 0; b < nblk; ++b) { num += blocks2[2 * b]; den += blocks2[2 * b + 1]; }
  const float jit = halton_f(rga, max_bits) * jam + (1.f - jam);
  const float grad = (jit * T) * num / den;
  const float log_T = logf(T);
  float ad_mu = b1 * st[CH_AD_MU] + (1.f - b1) * grad;
  float ad_nu = b2 * st[CH_AD_NU] + (1.f - b2) * grad * grad;
  const float count = st[CH_AD_COUNT] + 1.f;
  const float mu_hat = ad_mu / (1.f - powf(b1, count));
  const float nu_hat = ad_nu / (1.f - powf(b2, count));
  float upd = -lr * (mu_hat / (sqrtf(nu_hat) + 1e-8f));
  upd = fminf(fmaxf(upd, -0.35f), 0.35f);
  if (isnan(-lr * (mu_hat / (sqrtf(nu_hat) + 1e-8f)))) upd = __int_as_float(0x7fc00000);
  float new_log_T = log_T;
  if (isfinite(log_T + upd)) {
    new_log_T = log_T + upd;
    st[CH_AD_MU] = ad_mu; st[CH_AD_NU] = ad_nu; st[CH_AD_COUNT] = count;
  }
  const float log_T_ma = (1.f - uw) * st[CH_LOG_T_MA] + uw * new_log_T;
  float newT = expf(log_T_ma);
  newT = fminf(fmaxf(newT, eps), max_leapfrog * eps);
  st[CH_EPS] = eps; st[CH_LOG_EPS_MA] = log_eps_ma; st[CH_T] = newT; st[CH_LOG_T_MA] = log_T_ma;
  st[CH_DA_LOGX] = log_x; st[CH_DA_LOGX_AVG] = log_x_avg; st[CH_DA_STEP] = da_step; st[CH_DA_ERR] = err;
  st[CH_RGA] = (float)(rga + 1);
  st[CH_STEP] = step + 1.f;
  st[CH_GRAD] = grad;
  // next transition: ceil(jitter(i + 1) * T / eps) leapfrog steps (integration_steps_fn :775-779)
  const float jn = halton_f(rga + 1, max_bits) * jam + (1.f - jam);
  st[CH_L] = ceilf(jn * (newT / eps));
  const int t = (int)st[CH_T_IDX];
  if (hist) { hist[4 * t] = eps; hist[4 * t + 1] = newT; hist[4 * t + 2] = st[CH_L]; hist[4 * t + 3] = grad; }
  st[CH_T_IDX] = (float)(t + 1);
}
__global__ void k_chees_init(float* st, float eps0, int max_bits, float jitter_amount) {
  if (threadIdx.x != 0) return;
  for (int k = 0; k < kChHdr; ++k) st[k] = 0.f;
  st[CH_EPS] = eps0; st[CH_T] = eps0;
  st[CH_DA_LOGX] = logf(eps0); st[CH_DA_STEP] = 1.f; st[CH_DA_MU] = logf(10.f * eps0);
  st[CH_STEP] = 1.f; st[CH_MAX_BITS] = (float)max_bits; st[CH_JITTER_AMOUNT] = jitter_amount;
  const float j0 = halton_f(0, max_bits) * jitter_amount + (1.f - jitter_amount);
  st[CH_L] = ceilf(j0 * (eps0 / eps0));
}
__global__ void k_fill_steps(int n, const float* __restrict__ src, int32_t* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (int32_t)src[0];
}
}  // namespace bjx

extern "C" int64_t bjx_chees_state_floats(int32_t n_chains_local, int32_t dim, int32_t n_ranks) {
  if (n_chains_local <= 0 || dim <= 0 || n_ranks <= 0) return 0;
  const int64_t nb = n_stat_blocks(n_chains_local), D = dim;
  return kChHdr + 2 * D + 3 * (int64_t)n_chains_local + nb * (3 + 3 * D) * (1 + n_ranks) + nb * 2 * (1 + n_ranks);
}

extern "C" int bjx_chees_init(bjx_handle_t h, float* state, float step_size, int32_t max_bits, float jitter_amount,
                              float* step_size_chain_out, int32_t* steps_chain_out) {
  if (!h || !state || !step_size_chain_out || !steps_chain_out) return bjx_fail(h, BJX_E_INVALID, "null argument");
  if (!(step_size > 0.f) || max_bits < 1 || max_bits > 30) return bjx_fail(h, BJX_E_INVALID, "bad step_size / max_bits");
  AD_CUDA(cudaSetDevice(h->cfg.device));
  const int C = h->cfg.n_chains;
  k_chees_init<<<1, 32, 0, h->stream>>>(state, step_size, max_bits, jitter_amount);
  AD_LAUNCH("k_chees_init");
  k_fill_from<<<(C + 255) / 256, 256, 0, h->stream>>>(C, state + CH_EPS, step_size_chain_out);
  k_fill_steps<<<(C + 255) / 256, 256, 0, h->stream>>>(C, state + CH_L, steps_chain_out);
  AD_LAUNCH("k_fill");
  return 0;
}

extern "C" int bjx_chees_update(bjx_handle_t h, void* nccl_comm, int32_t n_ranks, float* state, const float* initial_position,
                                const float* proposal_position, const float* proposal_momentum, const float* acceptance_rate,
                                const uint8_t* is_divergent, float learning_rate, float b1, float b2, float target_acceptance,
                                float decay_rate, int32_t max_leapfrog_steps, float* step_size_chain, int32_t* steps_chain,
                                float* history) {
  if (!h || !state || !initial_position || !proposal_position || !proposal_momentum || !acceptance_rate || !is_divergent ||
      !step_size_chain || !steps_chain)
    return bjx_fail(h, BJX_E_INVALID, "null argument");
  if (n_ranks < 1 || (n_ranks > 1 && !nccl_comm)) return bjx_fail(h, BJX_E_INVALID, "n_ranks > 1 needs a communicator");
  AD_CUDA(cudaSetDevice(h->cfg.device));
  const int C = h->cfg.n_chains, D = h->cfg.dim;
  const int nb = n_stat_blocks(C);
  const size_t S1 = 3 + 3 * (size_t)D;
  float* w = state + kChHdr + 2 * (size_t)D;
  float* num = w + C;
  float* den = num + C;
  float* loc1 = den + C;
  float* gat1 = loc1 + (size_t)nb * S1;
  float* loc2 = gat1 + (size_t)nb * S1 * n_ranks;
  float* gat2 = loc2 + (size_t)nb * 2;
  cudaStream_t s = h->stream;
  k_chees_w<<<(C + 7) / 8, 256, 0, s>>>(C, D, proposal_position, acceptance_rate, is_divergent, w);
  k_chees_pass1<<<dim3((D + 31) / 32, nb), dim3(32, 8), 0, s>>>(C, D, proposal_position, initial_position, acceptance_rate,
                                                              is_divergent, w, loc1);
  AD_LAUNCH("k_chees_pass1");
  const float* m1 = loc1;
  if (n_ranks > 1) {
    int rc = bjx_allgather_stats(h, nccl_comm, loc1, (int64_t)(nb * S1), gat1);
    if (rc) return rc;
    m1 = gat1;
  }
  k_chees_means<<<1, 512, 0, s>>>(D, nb * n_ranks, m1, state);
  k_chees_dots<<<(C + 7) / 8, 256, 0, s>>>(C, D, proposal_position, proposal_momentum, initial_position, acceptance_rate,
                                          is_divergent, state, num, den);
  k_chees_pass2<<<nb, 256, 0, s>>>(C, num, den, loc2);
  AD_LAUNCH("k_chees_pass2");
  const float* m2 = loc2;
  if (n_ranks > 1) {
    int rc = bjx_allgather_stats(h, nccl_comm, loc2, (int64_t)(nb * 2), gat2);
    if (rc) return rc;
    m2 = gat2;
  }
  k_chees_finish<<<1, 32, 0, s>>>(nb * n_ranks, m2, state, learning_rate, b1, b2, target_acceptance, decay_rate,
                                  (float)max_leapfrog_steps, history);
  k_fill_from<<<(C + 255) / 256, 256, 0, s>>>(C, state + CH_EPS, step_size_chain);
  k_fill_steps<<<(C + 255) / 256, 256, 0, s>>>(C, state + CH_L, steps_chain);
  AD_LAUNCH("k_chees_finish");
  return 0;
}

extern "C" int bjx_chees_final(bjx_handle_t h, const float* state, float* out2) {
  if (!h || !state || !out2) return bjx_fail(h, BJX_E_INVALID, "null argument");
  AD_CUDA(cudaSetDevice(h->cfg.device));
  float hst[kChHdr];
  AD_CUDA(cudaMemcpyAsync(hst, state, sizeof(hst), cudaMemcpyDeviceToHost, h->stream));
  AD_CUDA(cudaStreamSynchronize(h->stream));
  out2[0] = expf(hst[CH_LOG_EPS_MA]);                       // step_size
  out2[1] = expf(hst[CH_LOG_T_MA] - hst[CH_LOG_EPS_MA]);    // num_leapfrog_steps (integration_steps_params)
  return 0;
}
