// The opaque handle behind include/bjx.h (internal).
#pragma once
#include <string>

#include "../../include/bjx.h"
#include "bjx_kernels.cuh"

namespace bjx {
struct LaunchArgs;
struct BigLaunchArgs;
}
// A loaded target plug-in (bjx_plugin_load): the warp-kernel launcher (rows up to 1024 dims) and / or the big-row launcher
struct bjx_plugin_s {
  void* dl;
  int (*launch)(int kernel_id, int sc, int dm, const bjx::LaunchArgs* a);
  int (*launch_big)(int kernel_id, const bjx::BigLaunchArgs* a);
};

struct bjx_handle_s {
  bjx_config cfg;
  cudaStream_t stream;
  int sc;              // SizeClass
  int metric_kind;     // -1 until set
  bool metric_small_dense;
  const float* imm;    // caller-owned
  float* msqrt;        // owned
  size_t msqrt_elems;
  // NUTS workspace (owned, lazily allocated)
  bjx::NutsWs ws;
  void* ws_block;
  int ws_depth;
  int* h_flag;         // pinned
  int64_t last_leaf_launches, last_depth;
  // large-D dense path (bjx_dense.cu): working rows p, v, q, g [C,D] + GEMM workspace
  float* dense_block;
  size_t dense_bytes;
  void* gemm_ws;
  size_t gemm_ws_bytes;
  uint16_t* dense_mat_s[3];      // fp16-split copies of M^-1, L^-T, target precision ([D, 3D] each)
  const float* dense_mat_src[3]; // the float32 matrices they were built from
  unsigned dense_mat_ver[3];
  unsigned dense_version;        // bumped by bjx_set_metric / bjx_set_target (contents may change behind the same pointer)
  size_t dense_bytes_built;
  float* lr_block;               // low-rank metric (owned): U [D,k] | sigma [D] | 1/sigma [D] | lambda-1 [k] | 1/sqrt(lambda)-1 [k]
  int lr_k;
  void* dn_block;                // NUTS on the dense path (bjx_dense_nuts.cuh): compact rows, checkpoint velocities
  size_t dn_bytes;
  float* pool_scratch;           // scratch of bjx_pooled_stats_dense (slice partials of the D x D co-moment)
  size_t pool_scratch_bytes;
  cudaStream_t dense_stream[2];  // chain slices of the dense path run on their own streams (bjx_dense.cu)
  cudaEvent_t dense_fork, dense_join[2], dense_stagger;
  bool dense_stagger_armed;
  bool dense_streams_ready;
  int ncoef;            // integrator coefficient table (integrators.py:321-369); {0.5, 1, 0.5} = velocity Verlet
  float coef[11];
  bool general_integrator;
  int key_shared;          // see bjx_set_key_mode
  const int32_t* steps_dev;  // see bjx_set_integration_steps
  const float* ghmc_noise;   // see bjx_set_ghmc_noise
  uint32_t chain_offset;
  uint32_t* sample_keys;   // [num_steps, 2] step keys of bjx_hmc_sample
  size_t sample_keys_cap;
  std::string err;
};

int bjx_fail(bjx_handle_t h, int code, const std::string& msg);
int bjx_ensure_nuts_ws(bjx_handle_t h);  // the NUTS workspace of bjx_api.cu (shared with the dense path)
int bjx_cuda_fail(bjx_handle_t h, cudaError_t e, const char* where);


