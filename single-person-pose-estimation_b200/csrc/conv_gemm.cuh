// Host-side interface of the tcgen05 convolution kernels (conv_gemm.cu).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace hgb {

// geometry of the 128-pixel TMA box of an NHWC activation tensor
struct ActBox {
  int wb, hb, nb;  // box extent along W, H, N (wb*hb*nb == 128)
};
int act_box(int H, int W, ActBox* box);

// 4-D map over an NHWC bf16 tensor: dims (C, W, H, N), box (64, wb, hb, nb), 128-byte swizzle.
int make_tmap_act(CUtensorMap* out, const void* ptr, int N, int H, int W, int C);
// 2-D map over a row-major bf16 matrix [rows][cols]; box (64 cols, box_rows rows), 128-byte swizzle.
int make_tmap_mat(CUtensorMap* out, const void* ptr, int rows, int cols, int box_rows);

// BatchNorm applied to the INPUT operand inside the kernel (deferred BN: the normalised tensor is never stored).
// Forward: statistics from `sums` (training) or the moving averages; `write` = this launch also stores the saved
// (mean, rstd) pair and updates the moving averages, exactly like the stand-alone BN pass would.
// Weight gradient: scale/shift are rebuilt from `saved`.
struct BnInput {
  const float* sums = nullptr;    // [2C] sum, sum of squares (forward, training)
  const float* gamma = nullptr;   // non-null enables the transform
  const float* beta = nullptr;
  float* moving_mean = nullptr;
  float* moving_var = nullptr;
  float* saved = nullptr;         // [2C] mean, rstd
  int mode = 0;                   // 0: batch statistics from sums (training forward), 1: moving averages (inference),
                                  // 2: saved (mean, rstd) of the forward pass (weight gradient)
  int write = 0, M = 0, C = 0;
};

// BatchNorm BACKWARD applied to the A operand of a 1x1 dgrad inside the kernel (fused bn_bwd_apply): the kernel loads the
// dz tile and the y tile of the BatchNorm that follows the convolution, computes
//     dp = [y > 0] * (A*dz + B*y + C)          (the three-coefficient form of layer_kernels.cu, bit-identical)
// in shared memory, feeds it to the MMAs, stores it (the weight gradient reads it) and accumulates the bias gradient.
// One read of dp and one launch disappear per BatchNorm (the stand-alone pass moved r dz + r y + w dp, the GEMM r dp).
struct BnBwdInput {
  const float* bsums = nullptr;   // [2C] sum dz, sum dz*y
  const float* saved = nullptr;   // [2C] mean, rstd
  const float* gamma = nullptr;   // non-null enables the fusion
  float* dgamma = nullptr;        // [C] written by CTA 0
  float* dbeta = nullptr;
  float* dbias = nullptr;         // [C] += column sums of dp (as rounded to bf16)
  int M_stat = 0;                 // rows behind the sums (global batch under sync-BN)
  float pscale = 1.f;             // 1 / ranks under sync-BN (dgamma / dbeta are identical on every rank)
  int C = 0;
};

struct ConvGemmArgs {
  int N, H, W, Cin, Cout;   // Cout = channels actually stored (multiple of 32)
  int ksize;                // 1 or 3
  int tap_sign;             // +1 forward, -1 mirrored taps (dgrad)
  int relu;
  int ldc;                  // pitch (elements) of out / res1 / res2
  const float* bias;        // [Cout] or null
  const __nv_bfloat16* res1;
  const __nv_bfloat16* res2;
  __nv_bfloat16* out;
  float* stats;             // [2*Cout] (sum, sumsq), added to; or null
  const __nv_bfloat16* bn_y; // non-null: stats = (sum out, sum out*bn_y) -- fused BatchNorm-backward reduction
  int max_ctas = 0;         // > 0: cap of the persistent grid (side lanes leave SMs to the main chain)
  BnInput bn_in;            // 1x1 only
  BnInput bn_out;           // 1x1 only, inference: out = BN(relu(conv + bias)) (+ residuals) straight from the accumulator
  BnBwdInput bn_bwd;        // 1x1 dgrad only: the input map (tmA) is dz; tmZ = y of that BatchNorm, tmDP = where dp is stored
};
// tmA: activation map of the input; tmB: weight matrix [>=Cout rows][ksize^2*Cin], box rows = block_n
int conv_gemm_block_n(int Cout);
// tmC: activation map of the output tensor (the epilogue writes the tile with TMA bulk stores)
// tmR: activation map of res1 (or null): the residual tile is then fetched by TMA into the staging buffer
// tmY: activation map of bn_y (or null): used to prefetch its tiles into L2 ahead of the statistics pass
// tmZ / tmDP: activation maps of the BatchNorm's y tensor and of the dp tensor when a.bn_bwd is enabled (else null)
int launch_conv_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap* tmR,
                     const CUtensorMap* tmY, const ConvGemmArgs& a, cudaStream_t st, const CUtensorMap* tmZ = nullptr,
                     const CUtensorMap* tmDP = nullptr, const CUtensorMap* tmB64 = nullptr);
// tmB64: the same weight matrix with 64-row boxes (3x3, 128 output channels): enables the CTA-pair (cta_group::2) kernel
// whether launch_conv_gemm can run the fused BatchNorm-backward prologue for this shape
bool conv_gemm_supports_bn_bwd(int ksize, int Cin, int Cout);

struct WgradArgs {
  int N, H, W, Cin, Cout;
  int ksize;
  int Cin_valid, Cout_valid; // 0 = all; otherwise only dw[:Cout_valid][tap][:Cin_valid] is written (padded tensors)
  float* dw;                // [Cout_valid][ksize^2*Cin_valid] fp32, added to
  BnInput bn_in;            // 1x1 only: x is normalised in shared memory (saved statistics)
  int max_ctas = 0;         // > 0: cap of the grid (the weight-gradient lane leaves SMs to the main chain)
};
// tmDY: activation map of dy (C = Cout); tmX: activation map of x (C = Cin)
int launch_conv_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradArgs& a, cudaStream_t st);

}  // namespace hgb
