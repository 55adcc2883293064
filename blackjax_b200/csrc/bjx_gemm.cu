// Batched dense linear map for the large-D dense metric / dense Gaussian target:
//   Y[C, N] = alpha * X[C, K] . A[K, N] + beta * Cin[C, N]      (A symmetric: M^-1, precision; or L^-1)
// = the reference's `linear_map(M^-1, p)` / `-P x` / `L^-T z` (blackjax/util.py:57-61, lax.dot with
// precision="highest") for all chains at once, i.e. a [C,D] x [D,D] GEMM on the 5th-generation tensor cores.
//
// float32 accuracy on the fp16 tensor-core path.  One TF32/BF16/FP16 MMA (11/8/11-bit significands) cannot meet
// the 1e-5 parity tolerance, so every float32 operand is split into two binary16 terms x = x1 + x2 (+ a residual
// below 2^-22 |x|) and the three cross products that sit above that residual are accumulated in float32 in TMEM:
//     x.a  ~=  x1a1 + x2a1 + x1a2                      (dropped: x2a2, x1 r_a, r_x a1, each <= 2^-22 |x a|)
// -- measured 2-3e-7 of max|y| on the config-2 matrices, below the ~1e-6 summation error of a float32 SGEMM.
// binary16 has a 5-bit exponent, so the split is only that accurate when the operand sits high in its range: the
// split kernels (bjx_dense.cu) scale every activation ROW by a power of two (max|x| -> [2^13, 2^14)) and every
// constant matrix by one power of two, and hand this GEMM a per-row alpha that undoes both exactly (and carries
// the leapfrog coefficient eps_c, so per-chain step sizes fuse too).  Planes are laid out along K:
// X' = [x1|x2|x1], A' = [a1|a1|a2], so the whole thing is ONE plain fp16 GEMM with K' = 3K running at full tcgen05
// rate (warp-specialised TMA + tcgen05.mma.cta_group::2 kernel from the CUTLASS sm100 collective templates; SASS
// UTCHMMA.2CTA, UTMALDG/UTMASTG, LDTM).  History: CUTLASS' FastF32 in-mainloop conversion was bound by its
// shared-memory transform (1.29 ms per [65536,1024]x[1024,1024] product); three bf16 terms / six products
// pre-split by our kernels ran at the MMA rate (0.51 ms); two fp16 terms / three products halve the MMA work again.
// Epilogue: Y = alpha_row * acc + beta * Cin (the axpy of the leapfrog position update rides on it).
#include <cuda_runtime.h>

#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/epilogue/fusion/operations.hpp"
#include "cutlass/epilogue/thread/activation.h"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/dispatch_policy.hpp"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/util/packed_stride.hpp"

#include "bjx_internal.h"

// ---- a second epilogue: the two half kicks of consecutive leapfrog steps on the gradient product ---------------
//   D = alpha[row] * acc + (alpha[row] * acc + C)        (two separately rounded FMAs, integrators.py:134-141,235-239)
// Defined like the library's own fused operations: an operation tag plus the FusionCallbacks specialisation that
// maps it onto an epilogue visitor tree (the sm100 TMA epilogue forwards to the sm90 callbacks).
namespace cutlass::epilogue::fusion {

template <class ElementOutput_, class ElementCompute_, class ElementSource_ = ElementOutput_,
          class ElementScalar_ = ElementCompute_, FloatRoundStyle RoundStyle_ = FloatRoundStyle::round_to_nearest>
struct PerRowDoubleAxpy : LinearCombination<ElementOutput_, ElementCompute_, ElementSource_, ElementScalar_, RoundStyle_> {
  static constexpr bool IsPerRowScaleSupported = true;
};

template <class CtaTileShapeMNK, class ElementOutput, class ElementCompute, class ElementSource, class ElementScalar,
          FloatRoundStyle RoundStyle>
using Sm90PerRowDoubleAxpy =
    Sm90EVT<Sm90Compute<homogeneous_multiply_add, ElementOutput, ElementCompute, RoundStyle>,  // alpha * acc + (...)
            Sm90ColBroadcast<0, CtaTileShapeMNK, ElementScalar, ElementCompute, Stride<bool, _0, int64_t>, 1>,
            Sm90AccFetch,
            Sm90EVT<Sm90Compute<homogeneous_multiply_add, ElementCompute, ElementCompute, RoundStyle>,  // alpha * acc + C
                    Sm90ColBroadcast<0, CtaTileShapeMNK, ElementScalar, ElementCompute, Stride<bool, _0, int64_t>, 1>,
                    Sm90AccFetch,
                    Sm90SrcFetch<ElementSource>>>;

template <int StagesC, int StagesD, int FragmentSize, bool ReuseSmemC, bool DelayTmaStore, class ElementOutput,
          class ElementCompute, class ElementSource, class ElementScalar, FloatRoundStyle RoundStyle, class CtaTileShapeMNK,
          class EpilogueTile>
struct FusionCallbacks<epilogue::Sm90TmaWarpSpecialized<StagesC, StagesD, FragmentSize, ReuseSmemC, DelayTmaStore>,
                       fusion::PerRowDoubleAxpy<ElementOutput, ElementCompute, ElementSource, ElementScalar, RoundStyle>,
                       CtaTileShapeMNK, EpilogueTile>
    : Sm90PerRowDoubleAxpy<CtaTileShapeMNK, ElementOutput, ElementCompute, ElementSource, ElementScalar, RoundStyle> {
  using Impl = Sm90PerRowDoubleAxpy<CtaTileShapeMNK, ElementOutput, ElementCompute, ElementSource, ElementScalar, RoundStyle>;
  using Operation = fusion::PerRowDoubleAxpy<ElementOutput, ElementCompute, ElementSource, ElementScalar, RoundStyle>;

  struct Arguments {
    ElementScalar const* alpha_ptr = nullptr;  // [M] per-row factors

    operator typename Impl::Arguments() const {
      using StrideAlpha = Stride<bool, _0, int64_t>;
      const StrideAlpha dAlpha = {bool(1), _0{}, 0};
      return {
          {alpha_ptr, ElementScalar(0), dAlpha},  // leaf: alpha
          {},                                     // leaf: acc
          {
              {alpha_ptr, ElementScalar(0), dAlpha},  // leaf: alpha
              {},                                     // leaf: acc
              {},                                     // leaf: C
              {}                                      // multiply_add
          },
          {}  // multiply_add
      };
    }
  };

  using Impl::Impl;
};

}  // namespace cutlass::epilogue::fusion

namespace bjx {

using namespace cute;

using ElementA = cutlass::half_t;
using ElementB = cutlass::half_t;
using ElementC = float;
using ElementAcc = float;
using LayoutA = cutlass::layout::RowMajor;     // X' [C, 3K], K' contiguous
using LayoutB = cutlass::layout::ColumnMajor;  // B(k', n) = A'[n*3K + k']: row-major [N, 3K] storage
using LayoutC = cutlass::layout::RowMajor;
constexpr int kAlignAB = 8;                     // 16-byte TMA alignment (fp16)
constexpr int kAlignC = 4;                      // 16 bytes (float)

// CTA pair (cta_group::2): a 256x256 accumulator tile shared by two SMs, operands split between them
// (measured alternatives on config 2, ms per transition: 256x256x64 / cluster 2x1 48.9 -- this one; cluster 2x2 48.4-48.6;
// 256x128x64 52.0; 256x256x128 51.0-52.5: the step is power-capped, tile shape moves it by noise except where it hurts)
using MmaTileShape = Shape<_256, _256, _64>;
using ClusterShape = Shape<_2, _1, _1>;

// D = alpha[row] * acc + beta * C   (per-row alpha vector, scalar beta, no bias, identity activation)
// (scalar-aligned alpha vector: slices may start at any chain and hold any number of chains)
using FusionOp = cutlass::epilogue::fusion::PerRowLinCombPerRowBiasEltAct<cutlass::epilogue::thread::Identity, ElementC,
                                                                         ElementAcc, float, ElementC, float, 1, 1>;

using FusionKick = cutlass::epilogue::fusion::PerRowDoubleAxpy<ElementC, ElementAcc, ElementC, float>;

template <class Fusion>
struct GemmOf {
  using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, MmaTileShape, ClusterShape,
      cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementAcc, ElementC, LayoutC, kAlignC, ElementC,
      LayoutC, kAlignC, cutlass::epilogue::collective::EpilogueScheduleAuto, Fusion>::CollectiveOp;

  using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
      cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, ElementA, LayoutA, kAlignAB, ElementB, LayoutB, kAlignAB,
      ElementAcc, MmaTileShape, ClusterShape,
      cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
      cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;

  using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue>;
  using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;
};
using Gemm = GemmOf<FusionOp>::Gemm;          // Y = alpha_row * acc + beta * Cin
using GemmKick = GemmOf<FusionKick>::Gemm;    // Y = alpha_row * acc + (alpha_row * acc + Cin)

size_t gemm_workspace_bytes(int M, int N, int K3) {
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K3, 1}};
  typename GemmKick::Arguments args2{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K3, 1}};
  const size_t a = Gemm::get_workspace_size(args), b = GemmKick::get_workspace_size(args2);
  return a > b ? a : b;
}

template <class G, class SetFusion>
static int run_gemm(const void* Xs, const void* As, float* Y, const float* Cin, int M, int N, int K3, void* workspace,
                    cudaStream_t stream, SetFusion set_fusion) {
  using StrideA = typename G::GemmKernel::StrideA;
  using StrideB = typename G::GemmKernel::StrideB;
  using StrideC = typename G::GemmKernel::StrideC;
  using StrideD = typename G::GemmKernel::StrideD;
  StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, make_shape(M, K3, 1));
  StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, make_shape(N, K3, 1));
  StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, make_shape(M, N, 1));
  StrideD sd = cutlass::make_cute_packed_stride(StrideD{}, make_shape(M, N, 1));
  typename G::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm,
                             {M, N, K3, 1},
                             {reinterpret_cast<const ElementA*>(Xs), sa, reinterpret_cast<const ElementB*>(As), sb},
                             {{}, Cin ? Cin : Y, sc, Y, sd}};
  set_fusion(args.epilogue.thread);
  G gemm;
  cutlass::Status st = gemm.can_implement(args);
  if (st != cutlass::Status::kSuccess) return 100 + (int)st;
  st = gemm.initialize(args, workspace, stream);
  if (st != cutlass::Status::kSuccess) return 200 + (int)st;
  st = gemm.run(stream);
  if (st != cutlass::Status::kSuccess) return 300 + (int)st;
  return 0;
}

// X' [M, K3] fp16, A' [N, K3] fp16 (both K'-contiguous), Y/Cin [M, N] float, row_alpha [M] float (device).
//   double_kick == false:  Y[m,:] = row_alpha[m] * (X'.A'^T)[m,:] + beta * Cin[m,:]
//   double_kick == true :  Y[m,:] = row_alpha[m] * acc + (row_alpha[m] * acc + Cin[m,:])      (beta ignored)
// returns 0 on success, a positive code (stage*100 + cutlass::Status) otherwise
int gemm_split(const void* Xs, const void* As, float* Y, const float* Cin, const float* row_alpha, float beta, int M, int N,
               int K3, void* workspace, cudaStream_t stream, bool double_kick) {
  if (double_kick)
    return run_gemm<GemmKick>(Xs, As, Y, Cin, M, N, K3, workspace, stream, [&](auto& f) { f.alpha_ptr = row_alpha; });
  return run_gemm<Gemm>(Xs, As, Y, Cin, M, N, K3, workspace, stream, [&](auto& f) {
    f.alpha_ptr = row_alpha;
    f.beta = beta;
  });
}

}  // namespace bjx
