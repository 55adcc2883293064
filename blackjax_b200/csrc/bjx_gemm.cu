// Batched dense linear map for the large-D dense metric / dense Gaussian target:
//   Y[C, N] = alpha * X[C, K] . A[K, N] + beta * Cin[C, N]      (A symmetric: M^-1, precision; or L^-1)
// = the reference's `linear_map(M^-1, p)` / `-P x` / `L^-T z` (blackjax/util.py:57-61, lax.dot with
// precision="highest") for all chains at once, i.e. a [C,D] x [D,D] GEMM on the 5th-generation tensor cores.
//
// float32 accuracy on bf16 tensor cores.  TF32/BF16 MMAs (11/8-bit mantissas) cannot meet the 1e-5 parity
// tolerance, so every float32 operand is split into three bfloat16 terms x = x1 + x2 + x3 (24 mantissa bits)
// and the six cross products that sit above float32 rounding are accumulated in float32 in TMEM:
//     x.a  ~=  x1a1 + x2a1 + x3a1 + x1a2 + x2a2 + x1a3          (dropped: x2a3, x3a2 ~2^-24, x3a3 ~2^-32)
// The split is done ONCE per operand by our own streaming kernels (bjx_dense.cu k_rows_split3 /
// k_matrix_split3), which lay the planes out along K:  X' = [x1|x2|x3|x1|x2|x1],  A' = [a1|a1|a1|a2|a2|a3],
// so the whole thing is ONE plain bf16 GEMM with K' = 6K running at full tcgen05 rate (warp-specialised
// TMA + tcgen05.mma.cta_group::2 kernel from the CUTLASS sm100 collective templates; SASS UTCHMMA.2CTA,
// UTMALDG/UTMASTG, LDTM).  (A first version converted operands inside the mainloop -- CUTLASS' FastF32
// input-transform schedule -- and was bound by that shared-memory transform, not by the MMAs: 1.29 ms per
// [65536,1024]x[1024,1024] product whether 9 or 6 MMAs were issued.)
// The axpy of the leapfrog position update is fused through the (alpha, beta) epilogue.
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

using ElementA = cutlass::bfloat16_t;
using ElementB = cutlass::bfloat16_t;
using ElementC = float;
using ElementAcc = float;
using LayoutA = cutlass::layout::RowMajor;     // X' [C, 6K], K' contiguous
using LayoutB = cutlass::layout::ColumnMajor;  // B(k', n) = A'[n*6K + k']: row-major [N, 6K] storage
using LayoutC = cutlass::layout::RowMajor;
constexpr int kAlignAB = 8;                     // 16-byte TMA alignment (bf16)
constexpr int kAlignC = 4;                      // 16 bytes (float)

// CTA pair (cta_group::2): a 256x256 accumulator tile shared by two SMs, operands split between them
using MmaTileShape = Shape<_256, _256, _64>;
using ClusterShape = Shape<_2, _1, _1>;

using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTileShape, ClusterShape,
    cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementAcc, ElementC, LayoutC, kAlignC, ElementC,
    LayoutC, kAlignC, cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;

using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementA, LayoutA, kAlignAB, ElementB, LayoutB, kAlignAB,
    ElementAcc, MmaTileShape, ClusterShape,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
    cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;

using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue>;
using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;

size_t gemm_workspace_bytes(int M, int N, int K6) {
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K6, 1}};
  return Gemm::get_workspace_size(args);
}

// Y = alpha * X'.A'^T + beta * Cin with X' [M, K6] bf16, A' [N, K6] bf16 (both K'-contiguous), Y/Cin [M, N] float.
// returns 0 on success, a positive code (stage*100 + cutlass::Status) otherwise
int gemm_split(const void* Xs, const void* As, float* Y, const float* Cin, float alpha, float beta, int M, int N,
               int K6, void* workspace, cudaStream_t stream) {
  using StrideA = typename Gemm::GemmKernel::StrideA;
  using StrideB = typename Gemm::GemmKernel::StrideB;
  using StrideC = typename Gemm::GemmKernel::StrideC;
  using StrideD = typename Gemm::GemmKernel::StrideD;
  StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, make_shape(M, K6, 1));
  StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, make_shape(N, K6, 1));
  StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, make_shape(M, N, 1));
  StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, make_shape(M, N, 1));
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm,
                                {M, N, K6, 1},
                                {reinterpret_cast<const ElementA*>(Xs), sa, reinterpret_cast<const ElementB*>(As), sb},
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
