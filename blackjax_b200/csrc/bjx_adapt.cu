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
