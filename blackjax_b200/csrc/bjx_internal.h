// Internal declarations shared by the translation units of libbjx.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bjx {
void launch_prng_split(const uint32_t* keys, long long n, int num, uint32_t* out, cudaStream_t s);
void launch_prng_fold_in(const uint32_t* keys, long long n, uint32_t data, uint32_t* out, cudaStream_t s);
void launch_prng_draw(int mode, const uint32_t* keys, long long n, long long per_key, void* out, cudaStream_t s);
void launch_diag_mass_sqrt(const float* imm, long long n, float* out, cudaStream_t s);
void launch_da(int op, int C, float* st, const float* in, float target, float* eps_out, cudaStream_t s);
void launch_welford_update(long long n, const float* x, float* mean, float* m2, int count, cudaStream_t s);
void launch_welford_final(long long n, float* mean, float* m2, int count, float* imm, cudaStream_t s);
void launch_welford_dense_update(int C, int D, const float* x, float* mean, float* m2, int count, cudaStream_t s);
void launch_welford_dense_final(int C, int D, float* mean, float* m2, int count, float* imm, cudaStream_t s);
void launch_chol_linv_t(int C, int D, const float* imm, float* msqrt, cudaStream_t s);
void launch_pooled_stats(int C, int D, const float* x, const float* acc, float* out, cudaStream_t s);
void launch_pooled_stats_dense(int C, int D, const float* x, const float* acc, float* out, float* scratch, cudaStream_t s);
size_t pooled_dense_scratch_floats(int D);
void launch_rhat(int T, int C, int D, const float* hist, float* rhat, float* scratch, cudaStream_t s);
void launch_prng_randint(const uint32_t* keys, long long n, long long per_key, int minval, int maxval, int* out,
                         cudaStream_t s);
size_t ess_scratch_floats(int T, int C, int D);
void launch_ess(int T, int C, int D, const float* hist, float* ess, float* scratch, cudaStream_t s);
}  // namespace bjx
