// Batched dense linear map for the large-D dense metric / dense Gaussian target:
//   Y[C, N] = alpha * X[C, K] . A[K, N] + beta * Cin[C, N]      (A symmetric: M^-1, precision; or L^-1)
// = the reference's `linear_map(M^-1, p)` / `-P x` / `L^-T z` (blackjax/util.py:57-61, lax.dot with
// precision="highest") for all chains at once, i.e. a [C,D] x [D,D] GEMM.  It runs on the 5th-generation
// tensor cores: a warp-specialised TMA + tcgen05.mma kernel with TMEM accumulators assembled from CUTLASS
// sm100 templates, in the FastF32 operand-split mode (each float32 operand is split into three bfloat16
// terms in shared memory and the nine cross products are accumulated in float32 in TMEM) so the result is
// float32-accurate as `precision="highest"` requires -- plain TF32/BF16 MMAs (8-11 bit mantissas) would not
// meet the 1e-5 parity tolerance.  The axpy of the leapfrog is fused through the (alpha, beta) epilogue.
#include <cuda_runtime.h>

#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/dispatch_policy.hpp"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/util/packed_stride.hpp"

#include "bjx_internal.h"

namespace bjx {

using namespace cute;

using ElementA = float;
using ElementB = float;
using ElementC = float;
using ElementAcc = float;
using LayoutA = cutlass::layout::RowMajor;     // X [C, K], K contiguous
using LayoutB = cutlass::layout::ColumnMajor;  // B(k, n) = A[n*K + k]: row-major [N, K] storage (A symmetric or pre-transposed)
using LayoutC = cutlass::layout::RowMajor;
constexpr int kAlign = 4;                       // 16-byte TMA alignment

template <class MmaTileShape, class ClusterShape, class Schedule>
struct GemmCfg {
  using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTileShape, ClusterShape,
      cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementAcc, ElementC, LayoutC, kAlign, ElementC,
      LayoutC, kAlign, cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;
  using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementA, LayoutA, kAlign, ElementB, LayoutB, kAlign,
      ElementAcc, MmaTileShape, ClusterShape,
      cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
      Schedule>::CollectiveOp;
  using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue>;
  using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;
};

// CTA pair (cta_group::2): a 256x128 accumulator tile shared by two SMs, operands split between them
using Gemm = GemmCfg<Shape<_256, _128, _16>, Shape<_2, _1, _1>,
                     cutlass::gemm::KernelTmaWarpSpecialized2SmFastFP32SmemSm100>::Gemm;

size_t gemm_workspace_bytes(int M, int N, int K) {
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, 1}};
  return Gemm::get_workspace_size(args);
}

// returns 0 on success, a positive cutlass::Status code otherwise
int gemm_xa(const float* X, const float* A_nk, float* Y, const float* Cin, float alpha, float beta, int M, int N, int K,
            void* workspace, cudaStream_t stream) {
  using StrideA = typename Gemm::GemmKernel::StrideA;
  using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC;
  using StrideD = typename Gemm::GemmKernel::StrideD;
  StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, make_shape(M, K, 1));
  StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, make_shape(N, K, 1));
  StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, make_shape(M, N, 1));
  StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, make_shape(M, N, 1));
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm,
                                {M, N, K, 1},
                                {X, sa, A_nk, sb},
                                {{alpha, beta}, Cin ? Cin : Y, sc, Y, sd}};
  Gemm gemm;
  cutlass::Status st = gemm.can_implement(args);
  if (st != cutlass::Status::kSuccess) return 100 + (int)st;
  st = gemm.initialize(args, workspace, stream);
  if (st != cutlass::Status::kSuccess) return 200 + (int)st;
  st = gemm.run(stream);
  if (st != cutlass::Status::kSuccess) return 300 + (int)st;
  return 0;
}

}  // namespace bjx
