// Microbenchmark: FP32 issue throughput of FFMA vs FFMA2 (packed f32x2) on sm_100a, per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
  asm volatile("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}" : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
template <int MODE, int ILP>
__global__ void k(float* out, float a, float b, int iters) {
  float x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fmaf(a, x[i], b);
    } else {
#pragma unroll
      for (int i = 0; i < ILP; i += 2) fma2(x[i], x[i + 1], a, a, x[i], x[i + 1], b, b);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int ILP>
void run(const char* name, int warps_per_sm) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4 * 8);
  int iters = 20000;
  dim3 grid(148), block(warps_per_sm * 32);
  k<MODE, ILP><<<grid, block>>>(out, 1.0001f, 0.5f, 10);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE, ILP><<<grid, block>>>(out, 1.0001f, 0.5f, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fmas = (double)148 * warps_per_sm * 32 * ILP * iters;
  printf("%-8s ILP=%2d warps/SM=%2d  %.3f ms  %.1f TFLOP/s  (%.1f FMA/clk/SM @1.965GHz)\n", name, ILP, warps_per_sm, ms,
         2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<0, 16>("FFMA", 4); run<0, 16>("FFMA", 8); run<0, 16>("FFMA", 16); run<0, 32>("FFMA", 8);
  run<1, 16>("FFMA2", 4); run<1, 16>("FFMA2", 8); run<1, 16>("FFMA2", 16); run<1, 32>("FFMA2", 8);
  return 0;
}
