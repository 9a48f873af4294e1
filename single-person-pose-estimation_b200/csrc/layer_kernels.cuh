// Launchers of the non-GEMM layer kernels of the hourglass (layer_kernels.cu).
// All activations are bf16 NHWC viewed as [M = N*H*W rows][C channels], C % 8 == 0.
#pragma once
#include "common.cuh"

namespace hgb {

typedef __nv_bfloat16 bf16;

// BatchNormalization forward (model/hourglass.py:60,80,197-201; Keras defaults momentum .99, eps 1e-3)
// out = (y - mean) * rstd * gamma + beta (+ res).  training: mean/var from `sums` (sum, sumsq over M rows,
// produced by the conv epilogue); block 0 stores (mean, rstd) in `saved` and updates the moving
// statistics (moving_var gets the unbiased batch variance, as TF's fused kernel does).  M = rows of this tensor,
// M_stat = rows behind `sums` (= M, or the global batch under sync-BN).
// the same + the MaxPool2D 2x2/2 of the output (pooled: [N][H/2][W/2][C]), bit-identical to maxpool_fwd on `out`
int bn_apply_pool_fwd(const bf16* y, const bf16* res, bf16* out, bf16* pooled, const float* sums, float* saved, const float* gamma,
                      const float* beta, float* moving_mean, float* moving_var, int M, int M_stat, int C, int H, int W, int training,
                      cudaStream_t st);
int bn_apply_fwd(const bf16* y, const bf16* res, bf16* out, const float* sums, float* saved, const float* gamma,
                 const float* beta, float* moving_mean, float* moving_var, int M, int M_stat, int C, int training, cudaStream_t st,
                 const bf16* up = nullptr, int H = 0, int W = 0);   // up: + UpSampling2D(2x, nearest) of a [N][H/2][W/2][C] tensor

// MaxPool2D 2x2/2 (hourglass.py:63,135,171-177) on [N][2h][2w][C] -> [N][h][w][C], and its gradient
// (routed to the first maximum of each window in row-major order; accumulate=1 adds into dx).
int maxpool_fwd(const bf16* x, bf16* out, int N, int h, int w, int C, cudaStream_t st);
// ybn != null: dx is the dz of a BatchNorm with pre-normalisation input ybn (same shape as dx): bsums[0:C] += sum dz,
// bsums[C:2C] += sum dz * ybn over the values as stored -- the bn_bwd_reduce pass over the tensor disappears.
int maxpool_bwd(const bf16* x, const bf16* dy, bf16* dx, int N, int h, int w, int C, int accumulate, cudaStream_t st,
                const bf16* ybn = nullptr, float* bsums = nullptr);

// UpSampling2D (nearest 2x) + Add (hourglass.py:152-154): out[N][2h][2w][C] = skip + up(low); gradient wrt low
// = sum over each 2x2 block of dout (the gradient wrt skip is dout itself).
int upsample_add_fwd(const bf16* skip, const bf16* low, bf16* out, int N, int h, int w, int C, cudaStream_t st);
// (ybn / bsums: as in maxpool_bwd, for the BatchNorm whose dz is dlow)
int upsample_add_bwd(const bf16* dout, bf16* dlow, int N, int h, int w, int C, cudaStream_t st, const bf16* ybn = nullptr,
                     float* bsums = nullptr);

// BatchNorm backward, pass 1: bsums[0:C] += sum_rows dz, bsums[C:2C] += sum_rows dz*y
int bn_bwd_reduce(const bf16* dz, const bf16* y, float* bsums, int M, int C, cudaStream_t st);
// pass 2: dp = [y > 0] * gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat));  dbias[c] += sum_rows dp;
// block 0 also writes dgamma = sum dz*xhat, dbeta = sum dz.
int bn_bwd_apply(const bf16* dz, const bf16* y, bf16* dp, const float* bsums, const float* saved, const float* gamma,
                 float* dgamma, float* dbeta, float* dbias, int M, int M_stat, float pscale, int C, cudaStream_t st);

// Convs without BN: dp = relu ? g * [y > 0] : g (in place when dp == g; dp may be null when !relu),
// dbias[c] += sum_rows dp for c < c_valid.
// dbias2 (optional): a second bias gradient that receives the same column sums.
int relu_mask_colsum(const bf16* g, const bf16* y, bf16* dp, float* dbias, int M, int C, int c_valid, int relu,
                     cudaStream_t st, float* dbias2 = nullptr);

// Depthwise half of SeparableConv2D (bottleneck_block_mobile, hourglass.py:209-231): out = dw_conv(x, w) (+res1 +res2);
// w fp32 [k*k][C] (tap-major); flip = mirrored taps (the input gradient of the same layer).  And its weight gradient.
int dwconv(const bf16* x, const float* w, const bf16* res1, const bf16* res2, bf16* out, int N, int H, int W, int C, int ksize, int flip,
           cudaStream_t st);
int dwconv_wgrad(const bf16* x, const bf16* dy, float* dw, int N, int H, int W, int C, int ksize, cudaStream_t st);

// Prediction head (hourglass.py:83): logits [M][ldl] bf16 (first K valid) -> heat [M][K] f32 = act(logits) and
// pbf [M][64] bf16 (act value, zero in the padding channels) for the re-injection conv (hourglass.py:88).
int head_act_fwd(const bf16* logits, int ldl, float* heat, bf16* pbf, int M, int K, int sigmoid, cudaStream_t st);
// dlogits [M][64] = (dLdp [M][K] f32 + g_p [M][64] bf16 (may be null)) * act'(heat); zero padding.
int head_act_bwd(const float* dLdp, const bf16* g_p, const float* heat, bf16* dlogits, int M, int K, int sigmoid,
                 cudaStream_t st);

// 7x7 stride-2 'same' patches of the f32 NHWC image (hourglass.py:59; TF pads 2 before / 3 after):
// col [N*(H/2)*(W/2)][192] bf16, K index = (ky*7 + kx)*3 + c, zero for K >= 147.
int im2col_7x7s2(const float* img, bf16* col, int N, int H, int W, cudaStream_t st);

// Keras legacy Adam on the flat parameter buffer (trainer.py:31).
int adam_step(float* w, const float* g, float* m, float* v, int64_t n, float lr_t, float b1, float b2, float eps,
              float grad_scale, cudaStream_t st);

// fp32 OHWI master kernels -> bf16 GEMM operands.  One launch for every conv of the model.
struct WeightSyncEntry {
  const float* w;   // [cout][taps][cin]
  bf16* wf;         // [cout_pad][taps*cin_pad]   forward operand (zero padded)
  bf16* wd;         // [cin_pad][taps*cout_pad]   dgrad operand, or null
  int taps, cin, cout, cin_pad, cout_pad;
};
int weight_sync(const WeightSyncEntry* entries_dev, int n_entries, int max_elems, cudaStream_t st);

}  // namespace hgb
