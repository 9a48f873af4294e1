// Non-GEMM layers of the hourglass: BatchNorm apply / backward, max-pool, nearest-upsample+add,
// prediction-head activation, 7x7 stem patch extraction, Adam and the fp32 -> bf16 weight refresh.
// All are HBM-bound streaming kernels: one thread owns 8 consecutive channels (one 16-byte access)
// of a row; per-channel reductions are accumulated in registers over the thread's rows, combined in
// shared memory and flushed with one atomic per channel per block.
#include <algorithm>

#include "layer_kernels.cuh"

namespace hgb {

static constexpr float kBnEps = 1e-3f;        // Keras BatchNormalization default
static constexpr float kBnMomentum = 0.99f;   // Keras BatchNormalization default

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 ld16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void st16(bf16* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

static inline int row_blocks(int M, int C, int rows_per_thread = 4) {
  const int R = 256 / (C / 8);
  // small tensors are latency-bound: fewer rows per thread (no serial chain of dependent loads) until the grid
  // covers the chip twice
  while (rows_per_thread > 1 && cdiv(M, R * rows_per_thread) < 2 * 148) rows_per_thread >>= 1;
  int b = cdiv(M, R * rows_per_thread);
  if (b > 148 * 8) b = 148 * 8;   // 8 resident 256-thread blocks per SM
  return b < 1 ? 1 : b;
}

// ---------------------------------------------------------------------------------- BN forward
// UNROLL rows per iteration: 4 without a residual, 2 with one -- either way 4 independent 16-byte loads in flight per
// thread at 64 registers (4 resident blocks per SM).
template <int UNROLL, bool RES>
__global__ void __launch_bounds__(256, 4) bn_apply_fwd_kernel(const bf16* __restrict__ y, const bf16* __restrict__ res,
                                                              bf16* __restrict__ out, const float* __restrict__ sums,
                                                              float* __restrict__ saved, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float* __restrict__ mm,
                                                              float* __restrict__ mv, int M, int M_stat, int C, int training,
                                                              const bf16* __restrict__ up, int lw, int lh) {
  pdl_trigger();
  pdl_wait();
  const int G = C >> 3, R = 256 / G;
  const int g = threadIdx.x % G, r0 = threadIdx.x / G;
  float sc[8], sh[8];
  const float invM = 1.f / (float)M_stat;     // rows behind the sums (the global batch under sync-BN)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    float mean, var;
    if (training) {
      mean = sums[c] * invM;
      var = fmaxf(sums[C + c] * invM - mean * mean, 0.f);
    } else {
      mean = mm[c];
      var = mv[c];
    }
    const float rstd = rsqrtf(var + kBnEps);
    sc[j] = gamma[c] * rstd;
    sh[j] = beta[c] - mean * sc[j];
    if (training && blockIdx.x == 0 && r0 == 0) {
      saved[c] = mean;
      saved[C + c] = rstd;
      const float unbiased = M_stat > 1 ? var * ((float)M_stat / (float)(M_stat - 1)) : var;
      mm[c] = mm[c] * kBnMomentum + mean * (1.f - kBnMomentum);
      mv[c] = mv[c] * kBnMomentum + unbiased * (1.f - kBnMomentum);
    }
  }
  const int stride = gridDim.x * R;
  for (int rb = blockIdx.x * R + r0; rb < M; rb += UNROLL * stride) {
    uint4 vy[UNROLL], vr[RES ? UNROLL : 1];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int r = rb + u * stride;
      if (r < M) {
        const size_t off = (size_t)r * C + g * 8;
        vy[u] = ld16(y + off);
        if (RES) vr[u] = ld16(res + off);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int r = rb + u * stride;
      if (r < M) {
        float f[8];
        unpack8(vy[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = f[j] * sc[j] + sh[j];
        if (RES) {
          float q[8];
          unpack8(vr[u], q);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += q[j];
          if (up) {     // + UpSampling2D(nearest 2x) of the lower level: pixel (n, y, x) reads (n, y/2, x/2); W = 1<<lw, H = 1<<lh
            const int x = r & ((1 << lw) - 1), t = r >> lw, y = t & ((1 << lh) - 1), n = t >> lh;
            const size_t lr = ((((size_t)n << (lh - 1)) + (size_t)(y >> 1)) << (lw - 1)) + (size_t)(x >> 1);
            unpack8(ld16(up + lr * C + g * 8), q);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += q[j];
          }
        }
        st16(out + (size_t)r * C + g * 8, pack8(f));
      }
    }
  }
}

// BatchNorm (+ residual) AND the MaxPool2D 2x2/2 that follows it (hourglass.py:63,135,171-177) in one pass: a thread owns a 2x2
// window of pixels (8 channels), writes the four normalised pixels and their maximum.  The maximum is taken over the bf16 values
// as stored, so the pooled tensor is bit-identical to what maxpool_fwd_kernel computes from `out`; what disappears is that
// kernel's re-read of `out`.  W = 1 << lw, H = 1 << lh.
template <bool RES>
__global__ void __launch_bounds__(256, 3) bn_apply_pool_fwd_kernel(const bf16* __restrict__ y, const bf16* __restrict__ res,
                                                                   bf16* __restrict__ out, bf16* __restrict__ pooled,
                                                                   const float* __restrict__ sums, float* __restrict__ saved,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   float* __restrict__ mm, float* __restrict__ mv, int M, int M_stat,
                                                                   int C, int training, int lw, int lh) {
  pdl_trigger();
  pdl_wait();
  const int G = C >> 3, R = 256 / G;
  const int g = threadIdx.x % G, r0 = threadIdx.x / G;
  float sc[8], sh[8];
  const float invM = 1.f / (float)M_stat;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    float mean, var;
    if (training) {
      mean = sums[c] * invM;
      var = fmaxf(sums[C + c] * invM - mean * mean, 0.f);
    } else {
      mean = mm[c];
      var = mv[c];
    }
    const float rstd = rsqrtf(var + kBnEps);
    sc[j] = gamma[c] * rstd;
    sh[j] = beta[c] - mean * sc[j];
    if (training && blockIdx.x == 0 && r0 == 0) {
      saved[c] = mean;
      saved[C + c] = rstd;
      const float unbiased = M_stat > 1 ? var * ((float)M_stat / (float)(M_stat - 1)) : var;
      mm[c] = mm[c] * kBnMomentum + mean * (1.f - kBnMomentum);
      mv[c] = mv[c] * kBnMomentum + unbiased * (1.f - kBnMomentum);
    }
  }
  const int W = 1 << lw, windows = M >> 2;
  for (int wi = blockIdx.x * R + r0; wi < windows; wi += gridDim.x * R) {
    const int xo = wi & ((W >> 1) - 1), t = wi >> (lw - 1), yo = t & ((1 << (lh - 1)) - 1), n = t >> (lh - 1);
    const size_t base = ((((size_t)n << lh) + (size_t)(2 * yo)) << lw) + (size_t)(2 * xo);
    const size_t rows[4] = {base, base + 1, base + W, base + W + 1};
    uint4 vy[4], vr[RES ? 4 : 1];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vy[u] = ld16(y + rows[u] * C + g * 8);
      if (RES) vr[u] = ld16(res + rows[u] * C + g * 8);
    }
    uint4 mx = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(vy[u], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = f[j] * sc[j] + sh[j];
      if (RES) {
        float q[8];
        unpack8(vr[u], q);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += q[j];
      }
      const uint4 o = pack8(f);
      st16(out + rows[u] * C + g * 8, o);
      if (u == 0) {
        mx = o;
      } else {
        const uint32_t a[4] = {mx.x, mx.y, mx.z, mx.w}, b[4] = {o.x, o.y, o.z, o.w};
        uint32_t m4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __nv_bfloat162 h = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a[k]), *reinterpret_cast<const __nv_bfloat162*>(&b[k]));
          m4[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
        mx = make_uint4(m4[0], m4[1], m4[2], m4[3]);
      }
    }
    st16(pooled + (size_t)wi * C + g * 8, mx);
  }
}

int bn_apply_pool_fwd(const bf16* y, const bf16* res, bf16* out, bf16* pooled, const float* sums, float* saved, const float* gamma,
                      const float* beta, float* moving_mean, float* moving_var, int M, int M_stat, int C, int H, int W, int training,
                      cudaStream_t st) {
  HGB_CHECK_ARG(C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, "bn_apply_pool: unsupported channel count %d", C);
  HGB_CHECK_ARG(H >= 2 && W >= 2 && (H & (H - 1)) == 0 && (W & (W - 1)) == 0 && M % (H * W) == 0,
                "bn_apply_pool: power-of-two maps required (got %dx%d)", H, W);
  if (M == 0) return HGB_OK;
  int lw = 0, lh = 0;
  while ((1 << lw) < W) ++lw;
  while ((1 << lh) < H) ++lh;
  const int blocks = row_blocks(M / 4, C, 1);
  if (res)
    launch_pdl(bn_apply_pool_fwd_kernel<true>, dim3(blocks), dim3(256), 0, st, y, res, out, pooled, sums, saved, gamma, beta,
               moving_mean, moving_var, M, M_stat, C, training, lw, lh);
  else
    launch_pdl(bn_apply_pool_fwd_kernel<false>, dim3(blocks), dim3(256), 0, st, y, res, out, pooled, sums, saved, gamma, beta,
               moving_mean, moving_var, M, M_stat, C, training, lw, lh);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

int bn_apply_fwd(const bf16* y, const bf16* res, bf16* out, const float* sums, float* saved, const float* gamma,
                 const float* beta, float* moving_mean, float* moving_var, int M, int M_stat, int C, int training, cudaStream_t st,
                 const bf16* up, int H, int W) {
  HGB_CHECK_ARG(C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, "bn_apply: unsupported channel count %d", C);
  if (M == 0) return HGB_OK;
  int lw = 0, lh = 0;
  if (up) {
    HGB_CHECK_ARG(res && H >= 2 && W >= 2 && (H & (H - 1)) == 0 && (W & (W - 1)) == 0 && M % (H * W) == 0,
                  "bn_apply: the fused upsample-add needs a residual and power-of-two maps (got %dx%d)", H, W);
    while ((1 << lw) < W) ++lw;
    while ((1 << lh) < H) ++lh;
  }
  // In inference the moving statistics are read-only, in training block 0 rewrites them after reading.
  if (res)
    launch_pdl(bn_apply_fwd_kernel<2, true>, dim3(row_blocks(M, C)), dim3(256), 0, st, y, res, out, sums, saved, gamma, beta,
               moving_mean, moving_var, M, M_stat, C, training, up, lw, lh);
  else
    launch_pdl(bn_apply_fwd_kernel<4, false>, dim3(row_blocks(M, C)), dim3(256), 0, st, y, res, out, sums, saved, gamma, beta,
               moving_mean, moving_var, M, M_stat, C, training, (const bf16*)nullptr, 0, 0);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

// Block-level combine of per-thread 8-channel partials: every thread parks its partial sums in shared memory
// ([row group][stat][channel], 2048 floats per statistic for any C), one thread per channel adds the row groups up and
// issues one global reduction.  (Shared-memory float atomicAdd compiles to a compare-and-swap spin loop: with R
// threads contending per address it cost several microseconds per block.)
template <int NSTAT>
__device__ __forceinline__ void block_channel_flush(float (&acc)[NSTAT][8], float* s_acc /*[R*NSTAT*C]*/, float* const* dst,
                                                    int C, int g, int c_valid, float* dst2 = nullptr) {
  const int G = C >> 3, R = 256 / G, r0 = threadIdx.x / G;
#pragma unroll
  for (int s = 0; s < NSTAT; ++s) {
    float4* d = reinterpret_cast<float4*>(s_acc + ((size_t)r0 * NSTAT + s) * C + g * 8);
    d[0] = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
    d[1] = make_float4(acc[s][4], acc[s][5], acc[s][6], acc[s][7]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NSTAT * C; i += blockDim.x) {
    const int s = i / C, c = i - s * C;
    if (dst[s] && c < c_valid) {
      float t = 0.f;
      for (int r = 0; r < R; ++r) t += s_acc[(size_t)r * NSTAT * C + i];
      atomicAdd(dst[s] + c, t);
      if (dst2) atomicAdd(dst2 + c, t);      // a second consumer of the same column sums (statistic 0 only callers)
    }
  }
}

// ---------------------------------------------------------------------------------- max-pool
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int64_t total,
                                                          int h, int w, int G) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    int64_t pix = i / G;
    const int ox = (int)(pix % w);
    pix /= w;
    const int oy = (int)(pix % h);
    const int n = (int)(pix / h);
    const size_t C = (size_t)G * 8;
    const size_t base = (((size_t)n * 2 * h + 2 * oy) * 2 * w + 2 * ox) * C + g * 8;
    float a[8], b[8], c[8], d[8], o[8];
    unpack8(ld16(x + base), a);
    unpack8(ld16(x + base + C), b);
    unpack8(ld16(x + base + 2 * w * C), c);
    unpack8(ld16(x + base + 2 * w * C + C), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaxf(a[j], b[j]), fmaxf(c[j], d[j]));
    st16(out + (size_t)i * 8, pack8(o));
  }
}

// STATS: the gradient this kernel writes is the dz of a BatchNorm (the closing one of the bottleneck below the pool): its
// backward statistics sum(dz), sum(dz * y) are accumulated here, over the values as stored (bf16), instead of by a
// bn_bwd_reduce pass that would read the tensor again.  A thread keeps one channel group for all its windows
// (256 % G == 0 and the grid stride is a multiple of 256), so the partial sums live in registers.
template <bool STATS>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                          bf16* __restrict__ dx, int64_t total, int h, int w, int G,
                                                          int accumulate, const bf16* __restrict__ ybn, float* __restrict__ bsums) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_acc[];
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    int64_t pix = i / G;
    const int ox = (int)(pix % w);
    pix /= w;
    const int oy = (int)(pix % h);
    const int n = (int)(pix / h);
    const size_t C = (size_t)G * 8;
    const size_t base = (((size_t)n * 2 * h + 2 * oy) * 2 * w + 2 * ox) * C + g * 8;
    const size_t offs[4] = {base, base + C, base + 2 * w * C, base + 2 * w * C + C};
    float v[4][8], gy[8];
    uint4 yv[STATS ? 4 : 1];
#pragma unroll
    for (int q = 0; q < 4; ++q) unpack8(ld16(x + offs[q]), v[q]);
    if (STATS) {
#pragma unroll
      for (int q = 0; q < 4; ++q) yv[q] = ld16(ybn + offs[q]);
    }
    unpack8(ld16(dy + (size_t)i * 8), gy);
    float o[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int best = 0;
      float bv = v[0][j];
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (v[q][j] > bv) { bv = v[q][j]; best = q; }  // strict: first maximum wins
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q][j] = (q == best) ? gy[j] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (accumulate) {
        float old[8];
        unpack8(*reinterpret_cast<const uint4*>(dx + offs[q]), old);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[q][j] += old[j];
      }
      const uint4 packed = pack8(o[q]);
      st16(dx + offs[q], packed);
      if (STATS) {
        float d[8], yy[8];
        unpack8(packed, d);
        unpack8(yv[q], yy);
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[0][j] += d[j]; acc[1][j] += d[j] * yy[j]; }
      }
    }
  }
  if (STATS) {
    const int C = G * 8;
    float* dst[2] = {bsums, bsums + C};
    block_channel_flush<2>(acc, s_acc, dst, C, threadIdx.x % G, C);
  }
}

static inline int flat_blocks(int64_t total) {
  int64_t b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return b < 1 ? 1 : (int)b;
}

int maxpool_fwd(const bf16* x, bf16* out, int N, int h, int w, int C, cudaStream_t st) {
  HGB_CHECK_ARG(C % 8 == 0, "maxpool: C %% 8");
  const int64_t total = (int64_t)N * h * w * (C / 8);
  if (total == 0) return HGB_OK;
  launch_pdl(maxpool_fwd_kernel, dim3(flat_blocks(total)), dim3(256), 0, st, x, out, total, h, w, C / 8);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}
int maxpool_bwd(const bf16* x, const bf16* dy, bf16* dx, int N, int h, int w, int C, int accumulate, cudaStream_t st,
                const bf16* ybn, float* bsums) {
  HGB_CHECK_ARG(C % 8 == 0, "maxpool: C %% 8");
  const int64_t total = (int64_t)N * h * w * (C / 8);
  if (total == 0) return HGB_OK;
  if (ybn) {
    HGB_CHECK_ARG(256 % (C / 8) == 0 && C <= 2048, "maxpool_bwd: statistics need a channel count that divides 2048 (got %d)", C);
    launch_pdl(maxpool_bwd_kernel<true>, dim3(flat_blocks(total)), dim3(256), 2 * 2048 * sizeof(float), st, x, dy, dx, total, h, w,
               C / 8, accumulate, ybn, bsums);
  } else {
    launch_pdl(maxpool_bwd_kernel<false>, dim3(flat_blocks(total)), dim3(256), 0, st, x, dy, dx, total, h, w, C / 8, accumulate,
               ybn, bsums);
  }
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

// ---------------------------------------------------------------------------------- upsample + add
__global__ void __launch_bounds__(256) upsample_add_fwd_kernel(const bf16* __restrict__ skip, const bf16* __restrict__ low,
                                                               bf16* __restrict__ out, int64_t total, int h, int w, int G) {
  pdl_trigger();
  pdl_wait();
  // one thread per (low-res pixel, channel group): reads 1 low + 4 skip vectors, writes 4
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    int64_t pix = i / G;
    const int ox = (int)(pix % w);
    pix /= w;
    const int oy = (int)(pix % h);
    const int n = (int)(pix / h);
    const size_t C = (size_t)G * 8;
    const size_t base = (((size_t)n * 2 * h + 2 * oy) * 2 * w + 2 * ox) * C + g * 8;
    const size_t offs[4] = {base, base + C, base + 2 * w * C, base + 2 * w * C + C};
    float l[8];
    unpack8(ld16(low + (size_t)i * 8), l);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float s[8];
      unpack8(ld16(skip + offs[q]), s);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] += l[j];
      st16(out + offs[q], pack8(s));
    }
  }
}

// STATS: as in maxpool_bwd_kernel -- dlow is the dz of the BatchNorm that closes the merged bottleneck one level down
template <bool STATS>
__global__ void __launch_bounds__(256) upsample_add_bwd_kernel(const bf16* __restrict__ dout, bf16* __restrict__ dlow,
                                                               int64_t total, int h, int w, int G, const bf16* __restrict__ ybn,
                                                               float* __restrict__ bsums) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_acc[];
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    int64_t pix = i / G;
    const int ox = (int)(pix % w);
    pix /= w;
    const int oy = (int)(pix % h);
    const int n = (int)(pix / h);
    const size_t C = (size_t)G * 8;
    const size_t base = (((size_t)n * 2 * h + 2 * oy) * 2 * w + 2 * ox) * C + g * 8;
    float a[8], b[8], c[8], d[8];
    uint4 yv = make_uint4(0, 0, 0, 0);
    if (STATS) yv = ld16(ybn + (size_t)i * 8);
    unpack8(ld16(dout + base), a);
    unpack8(ld16(dout + base + C), b);
    unpack8(ld16(dout + base + 2 * w * C), c);
    unpack8(ld16(dout + base + 2 * w * C + C), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (a[j] + b[j]) + (c[j] + d[j]);
    const uint4 packed = pack8(a);
    st16(dlow + (size_t)i * 8, packed);
    if (STATS) {
      float dd[8], yy[8];
      unpack8(packed, dd);
      unpack8(yv, yy);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[0][j] += dd[j]; acc[1][j] += dd[j] * yy[j]; }
    }
  }
  if (STATS) {
    const int C = G * 8;
    float* dst[2] = {bsums, bsums + C};
    block_channel_flush<2>(acc, s_acc, dst, C, threadIdx.x % G, C);
  }
}

int upsample_add_fwd(const bf16* skip, const bf16* low, bf16* out, int N, int h, int w, int C, cudaStream_t st) {
  HGB_CHECK_ARG(C % 8 == 0, "upsample_add: C %% 8");
  const int64_t total = (int64_t)N * h * w * (C / 8);
  if (total == 0) return HGB_OK;
  launch_pdl(upsample_add_fwd_kernel, dim3(flat_blocks(total)), dim3(256), 0, st, skip, low, out, total, h, w, C / 8);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}
int upsample_add_bwd(const bf16* dout, bf16* dlow, int N, int h, int w, int C, cudaStream_t st, const bf16* ybn, float* bsums) {
  HGB_CHECK_ARG(C % 8 == 0, "upsample_add: C %% 8");
  const int64_t total = (int64_t)N * h * w * (C / 8);
  if (total == 0) return HGB_OK;
  if (ybn) {
    HGB_CHECK_ARG(256 % (C / 8) == 0 && C <= 2048, "upsample_add_bwd: statistics need a channel count that divides 2048 (got %d)", C);
    launch_pdl(upsample_add_bwd_kernel<true>, dim3(flat_blocks(total)), dim3(256), 2 * 2048 * sizeof(float), st, dout, dlow, total, h,
               w, C / 8, ybn, bsums);
  } else {
    launch_pdl(upsample_add_bwd_kernel<false>, dim3(flat_blocks(total)), dim3(256), 0, st, dout, dlow, total, h, w, C / 8, ybn, bsums);
  }
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

// ---------------------------------------------------------------------------------- BN backward
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const bf16* __restrict__ dz, const bf16* __restrict__ y,
                                                            float* __restrict__ bsums, int M, int C) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_acc[];
  const int G = C >> 3, R = 256 / G;
  const int g = threadIdx.x % G, r0 = threadIdx.x / G;
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = 0.f; acc[1][j] = 0.f; }
  const int stride = gridDim.x * R;
  for (int rb = blockIdx.x * R + r0; rb < M; rb += 4 * stride) {
    uint4 va[4], vb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = rb + u * stride;
      if (r < M) {
        const size_t off = (size_t)r * C + g * 8;
        va[u] = ld16(dz + off);
        vb[u] = ld16(y + off);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (rb + u * stride < M) {
        float a[8], b[8];
        unpack8(va[u], a);
        unpack8(vb[u], b);
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[0][j] += a[j]; acc[1][j] += a[j] * b[j]; }
      }
    }
  }
  float* dst[2] = {bsums, bsums + C};
  block_channel_flush<2>(acc, s_acc, dst, C, g, C);
}

int bn_bwd_reduce(const bf16* dz, const bf16* y, float* bsums, int M, int C, cudaStream_t st) {
  HGB_CHECK_ARG(C % 8 == 0 && 256 % (C / 8) == 0, "bn_bwd_reduce: unsupported channel count %d", C);
  if (M == 0) return HGB_OK;
  launch_pdl(bn_bwd_reduce_kernel, dim3(row_blocks(M, C, 8)), dim3(256), 2 * 2048 * sizeof(float), st, dz, y, bsums, M, C);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

// dp = [y > 0] * gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)) with xhat = (y - mean)*rstd, written as
// A*dz + B*y + C with three per-channel coefficients: 24 registers of parameters instead of 40 and two fmas per element,
// which keeps the kernel at 64 registers = 4 resident blocks per SM (it ran at 112 registers / 2 blocks: 24 % of the
// warp slots, 4.3 TB/s at batch 32).
__global__ void __launch_bounds__(256, 4) bn_bwd_apply_kernel(const bf16* __restrict__ dz, const bf16* __restrict__ y,
                                                              bf16* __restrict__ dp, const float* __restrict__ bsums,
                                                              const float* __restrict__ saved, const float* __restrict__ gamma,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                              float* __restrict__ dbias, int M, int M_stat, float pscale, int C) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_acc[];
  const int G = C >> 3, R = 256 / G;
  const int g = threadIdx.x % G, r0 = threadIdx.x / G;
  const float invM = 1.f / (float)M_stat;     // rows behind the sums (the global batch under sync-BN)
  float cA[8], cB[8], cC[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    const float mean = saved[c], rstd = saved[C + c];
    const float sdz = bsums[c];
    const float sdzx = rstd * (bsums[C + c] - mean * sdz);  // sum dz * xhat
    const float a = gamma[c] * rstd;
    const float k = a * rstd * (sdzx * invM);
    cA[j] = a;
    cB[j] = -k;
    cC[j] = k * mean - a * (sdz * invM);
    if (blockIdx.x == 0 && r0 == 0) {
      dgamma[c] = sdzx * pscale;     // pscale = 1 / ranks under sync-BN: every rank holds the GLOBAL sums and the
      dbeta[c] = sdz * pscale;       // gradient bucket all-reduce adds the ranks up
    }
  }
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  const int stride = gridDim.x * R;
  for (int rb = blockIdx.x * R + r0; rb < M; rb += 2 * stride) {   // 4 independent 16-byte loads in flight per thread
    uint4 vd[2], vy[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = rb + u * stride;
      if (r < M) {
        const size_t off = (size_t)r * C + g * 8;
        vd[u] = ld16(dz + off);
        vy[u] = ld16(y + off);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = rb + u * stride;
      if (r < M) {
        float d[8], v[8], o[8];
        unpack8(vd[u], d);
        unpack8(vy[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t = fmaf(cA[j], d[j], fmaf(cB[j], v[j], cC[j]));
          o[j] = v[j] > 0.f ? t : 0.f;   // ReLU sits between the conv and the BN (hourglass.py:196-201)
        }
        const uint4 packed = pack8(o);
        st16(dp + (size_t)r * C + g * 8, packed);
        float rr[8];
        unpack8(packed, rr);              // bias gradient of the values the GEMMs will actually read
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[0][j] += rr[j];
      }
    }
  }
  float* dst[1] = {dbias};
  block_channel_flush<1>(acc, s_acc, dst, C, g, C);
}

int bn_bwd_apply(const bf16* dz, const bf16* y, bf16* dp, const float* bsums, const float* saved, const float* gamma,
                 float* dgamma, float* dbeta, float* dbias, int M, int M_stat, float pscale, int C, cudaStream_t st) {
  HGB_CHECK_ARG(C % 8 == 0 && 256 % (C / 8) == 0, "bn_bwd_apply: unsupported channel count %d", C);
  if (M == 0) return HGB_OK;
  launch_pdl(bn_bwd_apply_kernel, dim3(row_blocks(M, C, 8)), dim3(256), 2048 * sizeof(float), st, dz, y, dp, bsums, saved, gamma, dgamma, dbeta,
                                                                           dbias, M, M_stat, pscale, C);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

__global__ void __launch_bounds__(256) relu_mask_colsum_kernel(const bf16* __restrict__ gsrc, const bf16* __restrict__ y,
                                                               bf16* __restrict__ dp, float* __restrict__ dbias, int M, int C,
                                                               int c_valid, int relu, float* __restrict__ dbias2) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_acc[];
  const int G = C >> 3, R = 256 / G;
  const int g = threadIdx.x % G, r0 = threadIdx.x / G;
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  for (int r = blockIdx.x * R + r0; r < M; r += gridDim.x * R) {
    const size_t off = (size_t)r * C + g * 8;
    float d[8];
    unpack8(*reinterpret_cast<const uint4*>(gsrc + off), d);
    if (relu) {
      float v[8];
      unpack8(ld16(y + off), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = v[j] > 0.f ? d[j] : 0.f;
      st16(dp + off, pack8(d));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] += d[j];
  }
  float* dst[1] = {dbias};
  block_channel_flush<1>(acc, s_acc, dst, C, g, c_valid, dbias2);
}

int relu_mask_colsum(const bf16* g, const bf16* y, bf16* dp, float* dbias, int M, int C, int c_valid, int relu,
                     cudaStream_t st, float* dbias2) {
  HGB_CHECK_ARG(C % 8 == 0 && 256 % (C / 8) == 0, "relu_mask_colsum: unsupported channel count %d", C);
  if (M == 0) return HGB_OK;
  launch_pdl(relu_mask_colsum_kernel, dim3(row_blocks(M, C, 8)), dim3(256), 2048 * sizeof(float), st, g, y, dp, dbias, M, C, c_valid, relu,
             dbias2);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

// ---------------------------------------------------------------------------------- depthwise convolution (mobile variant)
// SeparableConv2D of bottleneck_block_mobile (model/hourglass.py:209-231) = depthwise k x k 'same' convolution (one filter per
// channel, no bias, no activation) followed by a pointwise 1x1 convolution (bias + activation), which runs on the tcgen05 GEMM
// like every other 1x1 layer.  The depthwise half is a per-channel stencil: HBM-bound, one thread per (pixel, 8 channels).
//   forward  : out[p][c] = sum_tap x[p + tap][c] * w[tap][c]                  (fp32 weights, fp32 accumulate, bf16 out)
//   backward : dx[p][c]  = sum_tap dy[p - tap][c] * w[tap][c] (+ residuals)   = the same stencil with mirrored taps (flip)
//   weights  : dw[tap][c] += sum_p dy[p][c] * x[p + tap][c]
template <int KS>
__global__ void __launch_bounds__(256) dwconv_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const bf16* __restrict__ res1,
                                                     const bf16* __restrict__ res2, bf16* __restrict__ out, int64_t total, int H, int W, int G,
                                                     int flip) {
  pdl_trigger();
  pdl_wait();
  constexpr int R = KS / 2;
  const int C = G * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % G);
    int64_t pix = i / G;
    const int px = (int)(pix % W);
    pix /= W;
    const int py = (int)(pix % H);
    const int64_t n = pix / H;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
      const int yy = py + ky - R;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const int xx = px + kx - R;
        if (xx < 0 || xx >= W) continue;
        float v[8];
        unpack8(ld16(x + ((n * H + yy) * W + xx) * C + g * 8), v);
        const int tap = flip ? (KS - 1 - ky) * KS + (KS - 1 - kx) : ky * KS + kx;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + (size_t)tap * C + g * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + (size_t)tap * C + g * 8) + 1);
        acc[0] = fmaf(v[0], w0.x, acc[0]); acc[1] = fmaf(v[1], w0.y, acc[1]); acc[2] = fmaf(v[2], w0.z, acc[2]); acc[3] = fmaf(v[3], w0.w, acc[3]);
        acc[4] = fmaf(v[4], w1.x, acc[4]); acc[5] = fmaf(v[5], w1.y, acc[5]); acc[6] = fmaf(v[6], w1.z, acc[6]); acc[7] = fmaf(v[7], w1.w, acc[7]);
      }
    }
    if (res1) {
      float q[8];
      unpack8(*reinterpret_cast<const uint4*>(res1 + (size_t)i * 8), q);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += q[j];
    }
    if (res2) {
      float q[8];
      unpack8(*reinterpret_cast<const uint4*>(res2 + (size_t)i * 8), q);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += q[j];
    }
    st16(out + (size_t)i * 8, pack8(acc));
  }
}

int dwconv(const bf16* x, const float* w, const bf16* res1, const bf16* res2, bf16* out, int N, int H, int W, int C, int ksize, int flip,
           cudaStream_t st) {
  HGB_CHECK_ARG(C % 8 == 0 && (ksize == 1 || ksize == 3), "dwconv: C %% 8 and a 1x1 or 3x3 kernel required");
  const int64_t total = (int64_t)N * H * W * (C / 8);
  if (total == 0) return HGB_OK;
  if (ksize == 1)
    launch_pdl(dwconv_kernel<1>, dim3(flat_blocks(total)), dim3(256), 0, st, x, w, res1, res2, out, total, H, W, C / 8, flip);
  else
    launch_pdl(dwconv_kernel<3>, dim3(flat_blocks(total)), dim3(256), 0, st, x, w, res1, res2, out, total, H, W, C / 8, flip);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

// dw[tap][c] += sum over pixels of dy[p][c] * x[p + tap][c].  Thread = (8 channels, a strided set of pixel rows); per-thread
// partial sums for all taps in registers, combined per block through shared memory, one global reduction per (tap, channel).
template <int KS>
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, float* __restrict__ dw,
                                                           int N, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float s_acc[];       // [R][KS*KS][C]
  constexpr int T = KS * KS, Rd = KS / 2;
  const int G = C >> 3, R = 256 / G;
  const int g = threadIdx.x % G, r0 = threadIdx.x / G;
  float acc[T][8];
#pragma unroll
  for (int t = 0; t < T; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  const int64_t M = (int64_t)N * H * W;
  for (int64_t p = (int64_t)blockIdx.x * R + r0; p < M; p += (int64_t)gridDim.x * R) {
    const int px = (int)(p % W), py = (int)((p / W) % H);
    float d[8];
    unpack8(ld16(dy + p * C + g * 8), d);
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
      const int yy = py + ky - Rd;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const int xx = px + kx - Rd;
        if (xx < 0 || xx >= W) continue;
        float v[8];
        unpack8(ld16(x + (p + (int64_t)(ky - Rd) * W + (kx - Rd)) * C + g * 8), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[ky * KS + kx][j] = fmaf(d[j], v[j], acc[ky * KS + kx][j]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    float4* dst = reinterpret_cast<float4*>(s_acc + ((size_t)r0 * T + t) * C + g * 8);
    dst[0] = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
    dst[1] = make_float4(acc[t][4], acc[t][5], acc[t][6], acc[t][7]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    float t = 0.f;
    for (int r = 0; r < R; ++r) t += s_acc[(size_t)r * T * C + i];
    atomicAdd(dw + i, t);
  }
}

int dwconv_wgrad(const bf16* x, const bf16* dy, float* dw, int N, int H, int W, int C, int ksize, cudaStream_t st) {
  HGB_CHECK_ARG(C % 8 == 0 && 256 % (C / 8) == 0 && (ksize == 1 || ksize == 3), "dwconv_wgrad: unsupported shape");
  const int64_t M = (int64_t)N * H * W;
  if (M == 0) return HGB_OK;
  const int taps = ksize * ksize;
  const size_t smem = (size_t)256 * 8 * taps * sizeof(float);      // R * T * C floats = 256 threads x 8 channels x T
  int blocks = (int)std::min<int64_t>(148 * 2, (M + 63) / 64);
  if (blocks < 1) blocks = 1;
  if (ksize == 1) {
    launch_pdl(dwconv_wgrad_kernel<1>, dim3(blocks), dim3(256), smem, st, x, dy, dw, N, H, W, C);
  } else {
    static bool attr_done = false;
    if (!attr_done) {
      HGB_CUDA(cudaFuncSetAttribute(dwconv_wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_done = true;
    }
    launch_pdl(dwconv_wgrad_kernel<3>, dim3(blocks), dim3(256), smem, st, x, dy, dw, N, H, W, C);
  }
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

// ---------------------------------------------------------------------------------- prediction head
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void __launch_bounds__(256) head_act_fwd_kernel(const bf16* __restrict__ logits, int ldl, float* __restrict__ heat,
                                                           bf16* __restrict__ pbf, int M, int K, int sigmoid) {
  pdl_trigger();
  pdl_wait();
  // thread = (pixel, 8-channel group of the 64-wide padded row)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)M * 8; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i & 7);
    const int64_t r = i >> 3;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (g * 8 < K) {
      unpack8(ld16(logits + r * ldl + g * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        if (c < K) {
          f[j] = sigmoid ? sigmoidf_(f[j]) : f[j];
          heat[r * K + c] = f[j];
        } else {
          f[j] = 0.f;
        }
      }
    }
    st16(pbf + r * 64 + g * 8, pack8(f));
  }
}

int head_act_fwd(const bf16* logits, int ldl, float* heat, bf16* pbf, int M, int K, int sigmoid, cudaStream_t st) {
  HGB_CHECK_ARG(K > 0 && K <= 64 && ldl % 8 == 0, "head_act: K must be <= 64");
  if (M == 0) return HGB_OK;
  launch_pdl(head_act_fwd_kernel, dim3(flat_blocks((int64_t)M * 8)), dim3(256), 0, st, logits, ldl, heat, pbf, M, K, sigmoid);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

__global__ void __launch_bounds__(256) head_act_bwd_kernel(const float* __restrict__ dLdp, const bf16* __restrict__ g_p,
                                                           const float* __restrict__ heat, bf16* __restrict__ dlogits, int M,
                                                           int K, int sigmoid) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)M * 8; i += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(i & 7);
    const int64_t r = i >> 3;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (g * 8 < K) {
      float gp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (g_p) unpack8(ld16(g_p + r * 64 + g * 8), gp);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        if (c < K) {
          const float p = heat[r * K + c];
          const float d = dLdp[r * K + c] + gp[j];
          f[j] = sigmoid ? d * p * (1.f - p) : d;
        }
      }
    }
    st16(dlogits + r * 64 + g * 8, pack8(f));
  }
}

int head_act_bwd(const float* dLdp, const bf16* g_p, const float* heat, bf16* dlogits, int M, int K, int sigmoid,
                 cudaStream_t st) {
  HGB_CHECK_ARG(K > 0 && K <= 64, "head_act: K must be <= 64");
  if (M == 0) return HGB_OK;
  launch_pdl(head_act_bwd_kernel, dim3(flat_blocks((int64_t)M * 8)), dim3(256), 0, st, dLdp, g_p, heat, dlogits, M, K, sigmoid);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

// ---------------------------------------------------------------------------------- stem patches
// One block = 32 consecutive output pixels of one output row.  The 7 input rows x 69 input columns x 3 channels
// they touch are staged in shared memory with coalesced loads (zeros outside the image = TF 'same' padding,
// 2 before / 3 after); for a fixed kernel row the 21 K-elements of a pixel are then CONTIGUOUS in the staged row,
// so every thread assembles its 8 K-elements from shared memory and writes one 16-byte chunk.
constexpr int kStemPix = 32;                       // output pixels per block
constexpr int kStemCols = (2 * kStemPix + 5) * 3;  // staged floats per input row (207)

__global__ void __launch_bounds__(256) im2col_7x7s2_kernel(const float* __restrict__ img, bf16* __restrict__ col, int H, int W) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[7][kStemCols + 1];
  const int ow = W / 2, oh = H / 2;
  const int xb = blockIdx.x * kStemPix, oy = blockIdx.y, n = blockIdx.z;
  const int ix0 = xb * 2 - 2;
  for (int i = threadIdx.x; i < 7 * kStemCols; i += 256) {
    const int r = i / kStemCols, cc = i - r * kStemCols;
    const int iy = oy * 2 + r - 2, ix = ix0 + cc / 3;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(img + (((size_t)n * H + iy) * W + ix0) * 3 + cc);
    tile[r][cc] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kStemPix * 24; i += 256) {   // (pixel, group of 8 K-elements); 192 = 147 + padding
    const int px = i / 24, g = i - px * 24;
    if (xb + px >= ow) continue;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = g * 8 + j;
      const int ky = kk / 21;
      f[j] = kk < 147 ? tile[ky][6 * px + (kk - ky * 21)] : 0.f;
    }
    st16(col + ((((size_t)n * oh + oy) * ow + xb + px) * 24 + g) * 8, pack8(f));
  }
}

int im2col_7x7s2(const float* img, bf16* col, int N, int H, int W, cudaStream_t st) {
  HGB_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "im2col: even image size required");
  if (N == 0) return HGB_OK;
  HGB_CHECK_ARG(N <= 65535 && H / 2 <= 65535, "im2col: batch / height exceed the grid limits");
  launch_pdl(im2col_7x7s2_kernel, dim3(cdiv(W / 2, kStemPix), H / 2, N), dim3(256), 0, st, img, col, H, W);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

// ---------------------------------------------------------------------------------- Adam
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n4, float lr_t, float b1, float b2,
                                                   float eps, float gs) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 W = reinterpret_cast<float4*>(w)[i];
    const float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 Mv = reinterpret_cast<float4*>(m)[i];
    float4 Vv = reinterpret_cast<float4*>(v)[i];
    float* pw = &W.x; const float* pg = &G.x; float* pm = &Mv.x; float* pv = &Vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = pg[j] * gs;
      pm[j] = b1 * pm[j] + (1.f - b1) * gr;
      pv[j] = b2 * pv[j] + (1.f - b2) * gr * gr;
      pw[j] -= lr_t * pm[j] / (sqrtf(pv[j]) + eps);   // epsilon outside the bias correction (Keras)
    }
    reinterpret_cast<float4*>(w)[i] = W;
    reinterpret_cast<float4*>(m)[i] = Mv;
    reinterpret_cast<float4*>(v)[i] = Vv;
  }
}

int adam_step(float* w, const float* g, float* m, float* v, int64_t n, float lr_t, float b1, float b2, float eps,
              float grad_scale, cudaStream_t st) {
  HGB_CHECK_ARG(n % 4 == 0, "adam: element count must be a multiple of 4");
  if (n == 0) return HGB_OK;
  launch_pdl(adam_kernel, dim3(flat_blocks(n / 4)), dim3(256), 0, st, w, g, m, v, n / 4, lr_t, b1, b2, eps, grad_scale);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

// ---------------------------------------------------------------------------------- weight refresh
__global__ void __launch_bounds__(256) weight_sync_kernel(const WeightSyncEntry* __restrict__ entries) {
  pdl_trigger();
  pdl_wait();
  const WeightSyncEntry e = entries[blockIdx.y];
  const int kf = e.taps * e.cin_pad;
  const int nf = e.cout_pad * kf;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) {
    const int co = i / kf, r = i - co * kf, t = r / e.cin_pad, ci = r - t * e.cin_pad;
    float v = 0.f;
    if (co < e.cout && ci < e.cin) v = e.w[((size_t)co * e.taps + t) * e.cin + ci];
    e.wf[i] = __float2bfloat16_rn(v);
  }
  if (e.wd) {
    const int kd = e.taps * e.cout_pad;
    const int nd = e.cin_pad * kd;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += gridDim.x * blockDim.x) {
      const int ci = i / kd, r = i - ci * kd, t = r / e.cout_pad, co = r - t * e.cout_pad;
      float v = 0.f;
      if (co < e.cout && ci < e.cin) v = e.w[((size_t)co * e.taps + t) * e.cin + ci];
      e.wd[i] = __float2bfloat16_rn(v);
    }
  }
}

int weight_sync(const WeightSyncEntry* entries_dev, int n_entries, int max_elems, cudaStream_t st) {
  if (n_entries == 0) return HGB_OK;
  int bx = cdiv(max_elems, 256 * 4);
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  launch_pdl(weight_sync_kernel, dim3(bx, n_entries), dim3(256), 0, st, entries_dev);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

}  // namespace hgb
