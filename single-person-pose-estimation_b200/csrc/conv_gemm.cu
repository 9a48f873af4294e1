// Convolutions of the hourglass as implicit GEMMs on the 5th-generation tensor cores.
//
//   forward / dgrad : out[pixel][cout] = sum_tap sum_cin  x[pixel + tap][cin] * w[cout][tap][cin]
//                     (model/hourglass.py:193-200 -- Conv2D 1x1 / 3x3 'same', bias, ReLU)
//   wgrad           : dw[cout][tap][cin] += sum_pixel dy[pixel][cout] * x[pixel + tap][cin]
//
// One CTA = one 128-pixel tile.  A warp-specialised pipeline:
//   warp 0   TMA producer: a 4-D tensor map over the NHWC activation tensor delivers the 128 pixels
//            of the tile shifted by the filter tap; out-of-image pixels are zero-filled by TMA,
//            which is exactly TF 'same' padding.  Weights arrive through a 2-D map.
//   warp 1   owns TMEM and issues tcgen05.mma (one elected lane); accumulators live in TMEM.
//   warps 2-5 epilogue: tcgen05.ld -> bias / ReLU / residuals -> bf16 NHWC store, plus the
//            per-channel sum and sum-of-squares the following BatchNorm needs (warp butterfly).
// Operands sit in shared memory in the canonical 128-byte-swizzled layout that both TMA and the
// tensor-core descriptors understand; nothing is staged through registers.
#include "conv_gemm.cuh"
#include "sm100_ptx.cuh"

namespace hgb {

using namespace ptx;

#ifdef HGB_KTIME
// in-kernel timeline of CTA 0 (SM clock ticks), read back by hgb_debug_ktime(): build with -DHGB_KTIME
__device__ long long g_ktime[32];
#define KT(i) do { if (blockIdx.x == 0) g_ktime[i] = clock64(); } while (0)
#else
#define KT(i) do { } while (0)
#endif

constexpr int kBlockM = 128;
constexpr int kABytes = kBlockM * 128;  // 128 pixels x 64 bf16
constexpr int kThreads = 192;       // 3x3 wgrad kernel: TMA warp, MMA warp, 4 epilogue warps
constexpr int kWgradThreads = 256;  // wgrad kernel: + 2 operand-transform warps (deferred BatchNorm)
// forward/dgrad kernel: warp 0 TMA, warp 1 MMA, warps 2-3 operand transform, then EPI (8 or 16) epilogue warps
__host__ __device__ constexpr int gemm_threads(int epi) { return 128 + 32 * epi; }
#ifndef HGB_REGSPLIT
#define HGB_REGSPLIT 1
#endif
constexpr float kBnEpsIn = 1e-3f;       // Keras BatchNormalization defaults (as in layer_kernels.cu)
constexpr float kBnMomentumIn = 0.99f;

struct GemmKernelParams {
  int M_total, H, W, HW;
  int cblk;      // Cin / 64
  int nkb;       // taps * cblk
  int tap3;      // 1 when 3x3
  int tap_sign;
  int Cout, ldc, relu;
  const float* bias;
  const __nv_bfloat16* res1;
  const __nv_bfloat16* res2;
  __nv_bfloat16* out;
  float* stats;
  const __nv_bfloat16* bn_y;  // non-null: second statistic is sum(out * bn_y) (BatchNorm backward) instead of sum(out^2)
  int res1_tma;               // res1 is fetched by TMA (tmR) into the staging buffer
  BnInput bn_in;              // gamma != null: normalise the A tile in shared memory before the MMAs read it
  BnInput bn_out;             // gamma != null (inference): scale/shift of the BatchNorm that follows, applied in the epilogue
  BnBwdInput bn_bwd;          // gamma != null (BNB kernels): the A operand is BatchNorm-backward(dz, y), computed in shared memory
  int single_store;           // debug: one thread issues all output boxes (hgb_debug_set(16, 1))
  int late_trigger;           // programmatic dependent launch is released after the producer's last load
  int early_release;          // the epilogue hands the accumulator back right after its last TMEM load (0: after staging)
  int hoist_y;                // dgrad statistics: the first batch of y rows is loaded at the top of the tile (0: inside the statistics pass)
};

template <int OFF>
__device__ __forceinline__ void bfly_step(float (&x)[32], int lane) {
  const bool hi = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < OFF; ++i) {
    const float send = hi ? x[i] : x[i + OFF];
    const float keep = hi ? x[i + OFF] : x[i];
    x[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}
// x[j] = value of column j in this lane's row.  Afterwards x[0] of lane L = sum over the 32 rows of column L.
__device__ __forceinline__ float warp_column_sums(float (&x)[32], int lane) {
  bfly_step<16>(x, lane);
  bfly_step<8>(x, lane);
  bfly_step<4>(x, lane);
  bfly_step<2>(x, lane);
  bfly_step<1>(x, lane);
  return x[0];
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

// scale / shift of a deferred BatchNorm into shared memory (same fp32 arithmetic as bn_apply_fwd_kernel, so the
// normalised operand is bit-identical to what the stand-alone pass would have stored); `writer` stores the saved
// statistics and moving averages.  Backward (sums == null): rebuilt from the saved (mean, rstd).
__device__ __forceinline__ void bn_input_setup(const BnInput& b, float* s_sc, float* s_sh, int tid, int nthreads, bool writer) {
  const float invM = 1.f / (float)b.M;
  for (int c = tid; c < b.C; c += nthreads) {
    float mean, rstd;
    if (b.mode == 2) {   // weight gradient: saved statistics
      mean = b.saved[c];
      rstd = b.saved[b.C + c];
    } else {
      float var;
      if (b.mode == 0) {
        mean = b.sums[c] * invM;
        var = fmaxf(b.sums[b.C + c] * invM - mean * mean, 0.f);
      } else {
        mean = b.moving_mean[c];
        var = b.moving_var[c];
      }
      rstd = rsqrtf(var + kBnEpsIn);
      if (b.mode == 0 && writer) {
        b.saved[c] = mean;
        b.saved[b.C + c] = rstd;
        const float unbiased = b.M > 1 ? var * ((float)b.M / (float)(b.M - 1)) : var;
        b.moving_mean[c] = b.moving_mean[c] * kBnMomentumIn + mean * (1.f - kBnMomentumIn);
        b.moving_var[c] = b.moving_var[c] * kBnMomentumIn + unbiased * (1.f - kBnMomentumIn);
      }
    }
    const float sc = b.gamma[c] * rstd;
    s_sc[c] = sc;
    s_sh[c] = b.beta[c] - mean * sc;
  }
}

// z = bf16(y * sc + sh) on one 128-pixel x 64-channel swizzled box, in place, by the 64 transform threads.
// Thread tw always meets the same logical 16-byte channel chunk (its position XOR its row phase are constant).
__device__ __forceinline__ void bn_input_transform_box(uint32_t box, const float* s_sc, const float* s_sh, int ch0, int tw) {
  const uint32_t c = (uint32_t)tw & 7u, r0 = (uint32_t)tw >> 3;          // chunk position, first row (0..7)
  const int ch = ch0 + (int)((c ^ (r0 & 7u)) << 3);
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = s_sc[ch + j]; sh[j] = s_sh[ch + j]; }
#pragma unroll 4
  for (int j = 0; j < 16; ++j) {
    const uint32_t adr = box + (r0 + 8u * (uint32_t)j) * 128u + (c << 4);
    uint4 u;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(adr));
    float f[8];
    unpack_bf16x8(u, f);
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * k] * sc[2 * k] + sh[2 * k], f[2 * k + 1] * sc[2 * k + 1] + sh[2 * k + 1]);
      w[k] = *reinterpret_cast<uint32_t*>(&h);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(adr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  }
}

// Three coefficients per channel of the fused BatchNorm backward (same algebra as bn_bwd_apply_kernel):
// dp = [y > 0] * (A*dz + B*y + C).  CTA 0 also writes dgamma / dbeta.
__device__ __forceinline__ void bn_bwd_setup(const BnBwdInput& b, float* s_cA, float* s_cB, float* s_cC, int tid, int nthreads,
                                             bool writer) {
  const float invM = 1.f / (float)b.M_stat;
  for (int c = tid; c < b.C; c += nthreads) {
    const float mean = b.saved[c], rstd = b.saved[b.C + c];
    const float sdz = b.bsums[c];
    const float sdzx = rstd * (b.bsums[b.C + c] - mean * sdz);   // sum dz * xhat
    const float a = b.gamma[c] * rstd;
    const float k = a * rstd * (sdzx * invM);
    s_cA[c] = a;
    s_cB[c] = -k;
    s_cC[c] = k * mean - a * (sdz * invM);
    if (writer) {
      b.dgamma[c] = sdzx * b.pscale;
      b.dbeta[c] = sdz * b.pscale;
    }
  }
}

// dp = bf16([y > 0] * fma(A, dz, fma(B, y, C))) on one 128-pixel x 64-channel swizzled box pair (dz box rewritten in place,
// y box read), by the 64 transform threads; acc += the values as rounded (bias gradient), packed fp32 pairs.
__device__ __forceinline__ void bn_bwd_transform_box(uint32_t box_dz, uint32_t box_y, const float* s_cA, const float* s_cB,
                                                     const float* s_cC, int ch0, int tw, uint64_t (&acc)[4]) {
  const uint32_t c = (uint32_t)tw & 7u, r0 = (uint32_t)tw >> 3;          // chunk position, first row (0..7)
  const int ch = ch0 + (int)((c ^ (r0 & 7u)) << 3);
  uint64_t cA[4], cB[4], cC[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    cA[k] = pack_f32x2(__float_as_uint(s_cA[ch + 2 * k]), __float_as_uint(s_cA[ch + 2 * k + 1]));
    cB[k] = pack_f32x2(__float_as_uint(s_cB[ch + 2 * k]), __float_as_uint(s_cB[ch + 2 * k + 1]));
    cC[k] = pack_f32x2(__float_as_uint(s_cC[ch + 2 * k]), __float_as_uint(s_cC[ch + 2 * k + 1]));
  }
#pragma unroll 4
  for (int j = 0; j < 16; ++j) {
    const uint32_t off = (r0 + 8u * (uint32_t)j) * 128u + (c << 4);
    uint32_t d[4], y[4], w[4];
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(box_dz + off));
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(y[0]), "=r"(y[1]), "=r"(y[2]), "=r"(y[3]) : "r"(box_y + off));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t d2 = pack_f32x2(d[k] << 16, d[k] & 0xffff0000u);
      const uint64_t y2 = pack_f32x2(y[k] << 16, y[k] & 0xffff0000u);
      const uint32_t t = cvt_bf16x2(fma_f32x2(cA[k], d2, fma_f32x2(cB[k], y2, cC[k])), false);
      // ReLU sits between the conv and the BN (hourglass.py:196-201): the gradient passes where y > 0
      const __nv_bfloat162 yb = *reinterpret_cast<const __nv_bfloat162*>(&y[k]);
      const uint32_t m = __hgt2_mask(yb, __float2bfloat162_rn(0.f));
      w[k] = t & m;
      acc[k] = add_f32x2(acc[k], pack_f32x2(w[k] << 16, w[k] & 0xffff0000u));
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(box_dz + off), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
  }
}

// Persistent: grid = min(#tile groups, #SMs); CTA i handles groups i, i+grid, ...  The TMA producer runs
// ahead across group boundaries; the output tile is staged in swizzled smem and written with TMA bulk stores.
//   TILES == 1: one 128-pixel tile per group, two TMEM accumulator stages (the MMAs of the next tile overlap
//               the epilogue of the current one) -- the memory-bound 1x1 convolutions.
//   TILES == 4: weight-stationary: every weight box feeds four pixel tiles (four TMEM accumulators), which
//               cuts the shared-memory fill per MAC by 37 % -- the 3x3 convolutions, whose MMA rate is
//               bounded by the bytes that fit in flight (ncu: tensor pipe 40 % with TILES == 1).
//   HALO       : 3x3 only, TILES consecutive tiles of ONE image.  The weight-stationary ring still fetches every
//               input pixel nine times from L2 (once per tap), and L2 -> shared memory bandwidth, not the tensor
//               pipe, bounds it.  Here a (64-channel block, horizontal shift) "strip" of the TILES tiles plus one
//               halo row above and below is fetched once (TILES + 1 boxes) and serves the three vertical taps
//               through descriptor offsets of whole image rows (multiples of 1024 bytes: same swizzle phase):
//               2.4x fewer activation bytes per MAC.  Strip boxes and weight boxes are recycled slot by slot
//               (tile-major MMA order), accumulators are handed over and released tile by tile.
//   EPI        : epilogue warps.  8 = two per TMEM lane quarter.  16 = four per quarter, for the N = 256 tiles of the
//               HBM-bound 1x1 layers: their epilogue (TMEM drain, bias / ReLU / residual, staging, BatchNorm statistics:
//               ~700 instructions per thread per tile) ran at IPC 0.3 per scheduler with two warps each and took 2.85 us
//               per tile against 2.2 us of HBM time; four warps per scheduler hide the TMEM / shared-memory round trips.
//   BNB        : 1x1 dgrad with the BatchNorm backward of its input gradient fused in (BnBwdInput): every stage carries
//               the dz box and the y box; the transform warps turn the dz box into dp in place, hand it to the MMAs, and
//               one of them stores it through tmDP (the weight gradient reads dp later).  A stage is recycled once the
//               MMAs AND that bulk store have read it (empty barrier count 2).
//   CTA2       : HALO only: a cluster of two CTAs (one TPC) runs every MMA as tcgen05.mma.cta_group::2 with M = 256.  CTA r keeps
//               its own four tiles (own strip ring, own accumulators in its own TMEM, own epilogue) and HALF of every
//               weight box (64 of the 128 output channels); the leader's single thread issues the MMAs for both SMs.  Every
//               weight half is read from shared memory once and reaches both tensor cores, so the operand reads per SM drop
//               from 128 to 96 B/clk -- the port that pinned the one-CTA kernel at 55 % tensor-pipe utilisation (ncu).
//               Barriers: TMA loads of both CTAs count their bytes on the LEADER's full barriers; tcgen05.commit
//               multicasts to the empty / accumulator-full barriers of both CTAs; the epilogue warps of both CTAs arrive on
//               the leader's accumulator-empty barriers.
template <int BLOCK_N, int STAGES, int TILES, int OUT_BUFS, bool HALO = false, int EPI = 8, bool BNB = false, bool CTA2 = false>
__global__ void __launch_bounds__(gemm_threads(EPI), 1) conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const __grid_constant__ CUtensorMap tmC,
                                                                const __grid_constant__ CUtensorMap tmR,
                                                                const __grid_constant__ CUtensorMap tmY,
                                                                const __grid_constant__ CUtensorMap tmZ,
                                                                const __grid_constant__ CUtensorMap tmDP,
                                                                const GemmKernelParams p) {
  constexpr int kBBytes = (CTA2 ? BLOCK_N / 2 : BLOCK_N) * 128;                  // CTA2: this CTA's half of the weight box
  constexpr int kStageBytes = TILES * kABytes + kBBytes + (BNB ? kABytes : 0);   // BNB: + the y box behind the weights
  static_assert(!CTA2 || HALO, "CTA pairs are implemented for the strip-reuse 3x3 kernel");
  static_assert(!BNB || (TILES == 1 && !HALO), "the fused BatchNorm backward is a 1x1 dgrad variant");
  constexpr int kAccStages = TILES == 1 ? 2 : 1;
  constexpr int kGemmThreads = gemm_threads(EPI), kEpiThreads = 32 * EPI;
  static_assert(EPI == 8 || EPI == 16, "two or four epilogue warps per TMEM lane quarter");
  static_assert(kAccStages * TILES * BLOCK_N <= 512, "TMEM columns");
  constexpr int kOutBytes = (BLOCK_N / 64) * kABytes;
  // HALO: strip boxes | two sets of three weight boxes.  A strip needs the rows of its TILES tiles plus one above and one below:
  // TILES + 1 boxes when a box holds at least two image rows, TILES + 2 when a box is a single row (W = 128: the 64-channel
  // front-module layers, the only ones that get the extra slot -- the 128-channel ring has no room for it and never needs it)
  constexpr int kASlots = TILES + (BLOCK_N == 64 ? 2 : 1), kBSlots = 6;
  constexpr int kRingBytes = HALO ? kASlots * kABytes + kBSlots * kBBytes : STAGES * kStageBytes;
  // full[S] | empty[S] | tmem_full[2] | tmem_empty[2] | residual
  // HALO: fullA[T+1] | emptyA[T+1] | fullB[6] | emptyB[6] | tmem_full[T] | tmem_empty[T] | residual
  // (non-HALO) ... | ready[S]: operand tile normalised by the transform warps (deferred BatchNorm) | residual
  constexpr int kNumBars = HALO ? 2 * kASlots + 2 * kBSlots + 2 * TILES + 1 : 3 * STAGES + 5;
  static_assert(!HALO || (TILES > 1 && TILES * BLOCK_N <= 512), "HALO needs one accumulator per tile");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // 128-byte swizzle atoms need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t outbase = base + kRingBytes;                 // OUT_BUFS epilogue staging buffers (1024-aligned)
  const uint32_t bar0 = outbase + OUT_BUFS * kOutBytes;
  uint8_t* tail = smem + kRingBytes + OUT_BUFS * kOutBytes + (kNumBars * 8 + 15) / 16 * 16;   // 16-byte aligned
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail);
  float* s_bias = reinterpret_cast<float*>(tail + 16);        // [BLOCK_N]
  float* s_sc = s_bias + BLOCK_N;                             // [256] scale / [256] shift of a deferred input BatchNorm
  float* s_sh = s_sc + 256;
  float* s_osc = s_sh + 256;                                  // [256] / [256]: inference BatchNorm of the OUTPUT (epilogue)
  float* s_osh = s_osc + 256;                                 // (BNB: s_sc | s_sh | s_osc hold the coefficients A | B | C)
  float* s_part = s_osh + 256;                                // BNB only: [8 row phases][4 channel blocks][64] bias-gradient partials
  // [row groups][2*BLOCK_N] = 16 KB (32 KB with 16 epilogue warps), aliases pipeline stage 0: used only after the last tile.
  // BNB: the y box of stage 0 (read by the transform warps only, long before the last accumulator is complete) -- the dz / dp
  // box may still be being read by the last dp bulk store
  // (round 2) BNB: the output staging buffer instead, once the last output store has read it -- 32 KB with 16 epilogue warps
  float* s_stats = reinterpret_cast<float*>(smem + (BNB ? kRingBytes : 0));
  static_assert(!BNB || OUT_BUFS * kOutBytes >= (EPI == 16 ? 32 : 16) * 1024, "statistics scratch in the staging buffer");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (p.M_total + kBlockM - 1) / kBlockM;
  const int num_groups = (num_tiles + TILES - 1) / TILES;
  // the next kernel may start its own setup early; it blocks in pdl_wait() until this grid completes.  late_trigger:
  // only once this CTA's producer has issued its last load, so the dependent CTAs do not park on SMs (holding their
  // shared memory) for the whole duration of this kernel and keep other lanes' kernels out
  if (!p.late_trigger) pdl_trigger();
  if (threadIdx.x == 0) KT(0);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    if (p.res1_tma) prefetch_tmap(&tmR);
    if (p.bn_y) prefetch_tmap(&tmY);
    if (BNB) { prefetch_tmap(&tmZ); prefetch_tmap(&tmDP); }
    if (HALO) {
      for (int s = 0; s < kNumBars; ++s) {
        const bool tmem_empty = s >= 2 * kASlots + 2 * kBSlots + TILES && s < kNumBars - 1;
        mbar_init(bar0 + 8 * s, tmem_empty ? (CTA2 ? 2 * EPI : EPI) : 1);   // CTA2: the epilogue warps of both CTAs arrive on the leader's
      }
    } else {
      for (int s = 0; s < 2 * STAGES + 2; ++s) mbar_init(bar0 + 8 * s, (BNB && s >= STAGES && s < 2 * STAGES) ? 2 : 1);   // BNB: empty = MMAs + dp store
      mbar_init(bar0 + 8 * (2 * STAGES + 2), EPI);  // tmem_empty: one arrival per epilogue warp
      mbar_init(bar0 + 8 * (2 * STAGES + 3), EPI);
      for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + 8 * (2 * STAGES + 4 + s), 2);  // ready: two transform warps
      mbar_init(bar0 + 8 * (3 * STAGES + 4), 1);  // residual tile landed in the staging buffer
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTA2) {
      tmem_alloc_pair(smem_u32(tmem_slot), kAccStages * TILES * BLOCK_N);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(smem_u32(tmem_slot), kAccStages * TILES * BLOCK_N);
      tmem_relinquish();
    }
  }
  if (threadIdx.x == 0) KT(1);
  pdl_wait();      // everything above overlapped the previous kernel's tail; global memory is touched only below
  if (threadIdx.x == 0) KT(2);
  const bool xform = !BNB && !HALO && TILES == 1 && p.bn_in.gamma != nullptr;
  if (xform) bn_input_setup(p.bn_in, s_sc, s_sh, threadIdx.x, kGemmThreads, blockIdx.x == 0 && p.bn_in.write != 0);
  const bool post_bn = !BNB && !HALO && TILES == 1 && p.bn_out.gamma != nullptr;
  if (post_bn) bn_input_setup(p.bn_out, s_osc, s_osh, threadIdx.x, kGemmThreads, false);
  if (BNB) bn_bwd_setup(p.bn_bwd, s_sc, s_sh, s_osc, threadIdx.x, kGemmThreads, blockIdx.x == 0);
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_arrive_wait();   // the peer's barriers are initialised before anything is signalled across the pair
  tc_fence_after();
  if (threadIdx.x == 0) KT(3);
  // the bias is needed by the epilogue only: its (cold) load overlaps the first TMA loads instead of delaying them
  if (warp >= 4) {
    for (int i = threadIdx.x - 128; i < BLOCK_N; i += kEpiThreads) s_bias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
  }
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t full0 = bar0, empty0 = bar0 + 8 * STAGES;
  const uint32_t tfull0 = HALO ? bar0 + 8 * (2 * kASlots + 2 * kBSlots) : bar0 + 16 * STAGES;
  const uint32_t tempty0 = tfull0 + (HALO ? 8 * TILES : 16);
  const uint32_t resbar = bar0 + 8 * (kNumBars - 1);
  const uint32_t ready0 = bar0 + 8 * (2 * STAGES + 4);   // non-HALO only
  // HALO barriers and buffers
  const uint32_t fullA = bar0, emptyA = bar0 + 8 * kASlots, fullB = bar0 + 16 * kASlots, emptyB = fullB + 8 * kBSlots;
  const uint32_t ringB = base + kASlots * kABytes;
  const int rpt = kBlockM / p.W;   // image rows per tile
  const int nsl = (kASlots > TILES + 1 && rpt == 1) ? TILES + 2 : TILES + 1;   // HALO: strip boxes in use

  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  // Register re-partitioning (8 epilogue warps: 384 threads x 168 registers at launch).  The epilogue is the only
  // register-hungry role -- 64 accumulator values, 32 bias values and the statistics registers live at once -- and 168 was not
  // enough to also keep the dgrad statistics' y rows in flight across the TMEM loads (they spilled, and the hoist cost more
  // than it hid).  The producer / MMA / transform warp group hands registers back, the two epilogue warp groups take them.
  constexpr bool kRegSplit = EPI == 8 && HGB_REGSPLIT != 0;
  constexpr int kRegsLow = (BNB || HALO) ? 104 : 72, kRegsHigh = (BNB || HALO) ? 200 : 216;
  static_assert(128 * kRegsLow + 256 * kRegsHigh <= 384 * 168, "register re-partitioning must fit the launch allocation");
  if (warp < 4) {      // warp group 0: producer, MMA issuer, two operand-transform warps
  if constexpr (kRegSplit) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsLow));
  if (HALO && warp == 0) {
    if (lane == 0) {
      // CTA2: bytes of both CTAs are counted on the leader's full barriers (its producer expects twice the bytes)
      const uint32_t lfullA = CTA2 ? mapa_u32(fullA, 0) : fullA, lfullB = CTA2 ? mapa_u32(fullB, 0) : fullB;
      int sc = 0;   // running strip counter
      for (int grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
        const int p0 = grp * TILES * kBlockM;
        const int n0 = p0 / p.HW;
        const int y0 = (p0 - n0 * p.HW) / p.W;
        for (int st = 0; st < 3 * p.cblk; ++st, ++sc) {
          const int cb = st / 3, dxi = st - cb * 3;
          const int dx = (dxi - 1) * p.tap_sign;
          const int set = sc & 1;
          const uint32_t bph = (sc >> 1) & 1, aph = sc & 1;
#pragma unroll
          for (int dyi = 0; dyi < 3; ++dyi) {   // the three weight boxes of this column of taps
            const int slot = set * 3 + dyi;
            mbar_wait(emptyB + 8 * slot, bph ^ 1);
            if (CTA2) {
              if (rank == 0) mbar_expect_tx(fullB + 8 * slot, 2 * kBBytes);
              tma_load_2d_pair(ringB + slot * kBBytes, &tmB, lfullB + 8 * slot, ((dyi * 3 + dxi) * p.cblk + cb) * 64, (int)rank * (BLOCK_N / 2));
            } else {
              mbar_expect_tx(fullB + 8 * slot, kBBytes);
              tma_load_2d(ringB + slot * kBBytes, &tmB, fullB + 8 * slot, ((dyi * 3 + dxi) * p.cblk + cb) * 64, 0);
            }
          }
#pragma unroll
          for (int j = 0; j < kASlots; ++j) {   // strip rows y0 - 1 ... y0 + TILES * rpt (+ slack), shifted by dx
            if (j >= nsl) break;                // (the extra slot of the 64-channel ring is used by one-row boxes only)
            mbar_wait(emptyA + 8 * j, aph ^ 1);
            if (CTA2) {
              if (rank == 0) mbar_expect_tx(fullA + 8 * j, 2 * kABytes);
              tma_load_4d_pair(base + j * kABytes, &tmA, lfullA + 8 * j, cb * 64, dx, y0 - 1 + j * rpt, n0);
            } else {
              mbar_expect_tx(fullA + 8 * j, kABytes);
              tma_load_4d(base + j * kABytes, &tmA, fullA + 8 * j, cb * 64, dx, y0 - 1 + j * rpt, n0);
            }
          }
        }
      }
    }
    if (p.late_trigger) pdl_trigger();
  } else if (HALO && warp == 1) {
    if (lane == 0 && rank == 0) {   // CTA2: the leader's thread issues the M = 256 MMAs of the pair
      constexpr uint32_t idesc = make_idesc_bf16(CTA2 ? 2 * kBlockM : kBlockM, BLOCK_N, 0, 0);
      const uint32_t row_bytes = (uint32_t)p.W * 128u;
      int sc = 0, lt = 0;
      for (int grp = blockIdx.x; grp < num_groups; grp += gridDim.x, ++lt) {
        const int nst = 3 * p.cblk;
        for (int st = 0; st < nst; ++st, ++sc) {
          const int set = sc & 1;
          const uint32_t bph = (sc >> 1) & 1, aph = sc & 1;
#pragma unroll
          for (int dyi = 0; dyi < 3; ++dyi) mbar_wait(fullB + 8 * (set * 3 + dyi), bph);
          mbar_wait(fullA, aph);
#pragma unroll
          for (int t = 0; t < TILES; ++t) {
            mbar_wait(fullA + 8 * (t + 1), aph);      // tile t reads strip boxes t and t + 1 (+ t + 2 when a box is one row)
            if (kASlots > TILES + 1 && nsl > TILES + 1) mbar_wait(fullA + 8 * (t + 2), aph);
            if (st == 0) mbar_wait(tempty0 + 8 * t, (lt & 1) ^ 1);   // the epilogue has drained this accumulator
            tc_fence_after();
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              const int dy = (dyi - 1) * p.tap_sign;
              const uint64_t adesc = make_smem_desc_sw128(base + (uint32_t)(t * rpt + dy + 1) * row_bytes, 16, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(ringB + (set * 3 + dyi) * kBBytes, 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (CTA2) umma_bf16_pair(tmem_base + (uint32_t)(t * BLOCK_N), adesc + 2 * k, bdesc + 2 * k, idesc, (st | dyi | k) != 0);
                else umma_bf16(tmem_base + (uint32_t)(t * BLOCK_N), adesc + 2 * k, bdesc + 2 * k, idesc, (st | dyi | k) != 0);
              }
            }
            auto commit = [&](uint32_t bar) { if (CTA2) umma_commit_pair(bar); else umma_commit(bar); };
            commit(emptyA + 8 * t);              // later tiles start at box t + 1
            if (t == TILES - 1) {
              commit(emptyA + 8 * TILES);
              if (kASlots > TILES + 1 && nsl > TILES + 1) commit(emptyA + 8 * (TILES + 1));
#pragma unroll
              for (int dyi = 0; dyi < 3; ++dyi) commit(emptyB + 8 * (set * 3 + dyi));
            }
            if (st == nst - 1) commit(tfull0 + 8 * t);   // accumulator of tile t complete
          }
        }
      }
    }
  } else if (warp == 0) {
    if (lane == 0) {
      int kbt = 0;  // running k-block counter: the smem ring continues across tiles
      for (int grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
        int n0[TILES], y0[TILES];
#pragma unroll
        for (int t = 0; t < TILES; ++t) {   // tiles past the end address images >= N: TMA zero-fills them
          const int p0 = (grp * TILES + t) * kBlockM;
          n0[t] = p0 / p.HW;
          y0[t] = (p0 - n0[t] * p.HW) / p.W;
        }
        for (int kb = 0; kb < p.nkb; ++kb, ++kbt) {
          const int s = kbt % STAGES;
          const uint32_t ph = (kbt / STAGES) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const int tap = kb / p.cblk, cb = kb - tap * p.cblk;
          int dy = 0, dx = 0;
          if (p.tap3) {
            dy = (tap / 3 - 1) * p.tap_sign;
            dx = (tap % 3 - 1) * p.tap_sign;
          }
          const uint32_t sa = base + s * kStageBytes;
          mbar_expect_tx(full0 + 8 * s, kStageBytes);
          tma_load_2d(sa + TILES * kABytes, &tmB, full0 + 8 * s, kb * 64, 0);
#pragma unroll
          for (int t = 0; t < TILES; ++t)
            tma_load_4d(sa + t * kABytes, &tmA, full0 + 8 * s, cb * 64, dx, y0[t] + dy, n0[t]);
          if (BNB) tma_load_4d(sa + kABytes + kBBytes, &tmZ, full0 + 8 * s, cb * 64, 0, y0[0], n0[0]);
        }
      }
    }
    if (p.late_trigger) pdl_trigger();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, 0, 0);
      int kbt = 0, lt = 0;
      for (int grp = blockIdx.x; grp < num_groups; grp += gridDim.x, ++lt) {
        const int acc = lt % kAccStages;
        const uint32_t aph = (lt / kAccStages) & 1;
        if (lt == 3 || lt == 4) KT(23 + (lt - 3) * 3);
        mbar_wait(tempty0 + 8 * acc, aph ^ 1);  // the epilogue has drained this accumulator stage
        tc_fence_after();
        if (lt == 3 || lt == 4) KT(24 + (lt - 3) * 3);
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TILES * BLOCK_N);
        for (int kb = 0; kb < p.nkb; ++kb, ++kbt) {
          const int s = kbt % STAGES;
          const uint32_t ph = (kbt / STAGES) & 1;
          mbar_wait(((xform || BNB) ? ready0 : full0) + 8 * s, ph);
          tc_fence_after();
          if (kbt == 0) KT(4);
          const uint32_t sa = base + s * kStageBytes;
          const uint64_t bdesc = make_smem_desc_sw128(sa + TILES * kABytes, 16, 1024);
#pragma unroll
          for (int t = 0; t < TILES; ++t) {
            const uint64_t adesc = make_smem_desc_sw128(sa + t * kABytes, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)  // 4 x (K = 16) per 64-channel block: +32 bytes inside the swizzle atom
              umma_bf16(d_tmem + (uint32_t)(t * BLOCK_N), adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(empty0 + 8 * s);  // frees the smem stage once these MMAs have read it
        }
        umma_commit(tfull0 + 8 * acc);
        if (lt == 0) KT(5);
        if (lt == 3 || lt == 4) KT(25 + (lt - 3) * 3);
      }
    }
  } else {
    // ---------------- operand transform (deferred BatchNorm of the input): as soon as TMA has delivered the
    // 128-pixel x 64-channel box of a stage, z = bf16(y * scale + shift) is applied in place; the MMA warp waits
    // for `ready` instead of `full`.  Zero-filled rows past the end of the tensor become `shift`: they only reach
    // accumulator rows that the store clips and the statistics skip.
    if constexpr (BNB) {
      // fused BatchNorm backward: dz box (+ y box) -> dp box, in place; MMAs and the dp store read it from there
      const int tw = threadIdx.x - 64;
      uint64_t acc[4][4];          // [channel block][packed pair]: bias-gradient partial sums of this thread's 8 channels
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[i][k] = 0ull;
      int kbt = 0, pend = -1;      // pend: stage whose dp store was issued last; its smem read is awaited one box later
      for (int grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
        const int p0 = grp * kBlockM;
        const int n0 = p0 / p.HW;
        const int y0 = (p0 - n0 * p.HW) / p.W;
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {          // 1x1: one k-block per 64-channel block, at most 4 (static acc index)
          if (kb >= p.nkb) break;
          const int s = kbt % STAGES;
          mbar_wait(full0 + 8 * s, (kbt / STAGES) & 1);
          const uint32_t sa = base + s * kStageBytes;
          bn_bwd_transform_box(sa, sa + kABytes + kBBytes, s_sc, s_sh, s_osc, kb * 64, tw, acc[kb]);
          fence_proxy_async();   // generic-proxy writes -> visible to the tensor core and to the bulk store
          asm volatile("bar.sync 2, 64;" ::: "memory");   // both transform warps are done with the box
          if (lane == 0) mbar_arrive(ready0 + 8 * s);
          if (tw == 0) {
            tma_store_4d(&tmDP, sa, kb * 64, 0, y0, n0);
            tma_store_commit();
            if (pend >= 0) { tma_store_wait_read1(); mbar_arrive(empty0 + 8 * pend); }
            pend = s;
          }
          ++kbt;
        }
      }
      if (tw == 0 && pend >= 0) { tma_store_wait_read(); mbar_arrive(empty0 + 8 * pend); }
      // bias gradient: park the partials by (row phase, channel block), one thread per channel adds the 8 row phases up
      const uint32_t cpos = (uint32_t)tw & 7u, r0 = (uint32_t)tw >> 3;
      const int lch = (int)((cpos ^ r0) << 3);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float* d = s_part + ((int)r0 * 4 + i) * 64 + lch;
#pragma unroll
        for (int k = 0; k < 4; ++k) unpack_f32x2(acc[i][k], d[2 * k], d[2 * k + 1]);
      }
      asm volatile("bar.sync 2, 64;" ::: "memory");
      for (int i = tw; i < p.cblk * 64 && i < p.bn_bwd.C; i += 64) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) t += s_part[(r * 4 + (i >> 6)) * 64 + (i & 63)];
        atomicAdd(p.bn_bwd.dbias + i, t);
      }
    } else if (xform) {
      const int tw = threadIdx.x - 64;
      int kbt = 0;
      for (int grp = blockIdx.x; grp < num_groups; grp += gridDim.x) {
        for (int kb = 0; kb < p.nkb; ++kb, ++kbt) {
          const int s = kbt % STAGES;
          mbar_wait(full0 + 8 * s, (kbt / STAGES) & 1);
          bn_input_transform_box(base + s * kStageBytes, s_sc, s_sh, (kb % p.cblk) * 64, tw);
          fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's shared-memory reads
          __syncwarp();
          if (lane == 0) mbar_arrive(ready0 + 8 * s);
        }
      }
    }
  }
  } else {
    if constexpr (kRegSplit) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsHigh));
    // ---------------- epilogue: EPI warps; TMEM lane quarter = warp % 4.  The EPI / 4 warps of a quarter split every
    // 64-channel box into its two 32-column halves (hsel) and, with 16 warps, the boxes into even and odd ones (gsel)
    const int q = warp & 3;
    const int ew = warp - 4;                         // epilogue warp index
    const int hsel = (ew >> 2) & 1;
    const int gsel = ew >> 3;                        // 0 when EPI == 8
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;                // 0 .. kEpiThreads-1
    // lane 0 of every (EPI / 4)-th epilogue warp issues (and tracks) the bulk store of output box 0, 1, 2, 3
    constexpr int kWarpsPerBox = EPI / 4;
    const int store_box = p.single_store ? (et == 0 ? 0 : -1)
                        : (lane == 0 && ew % kWarpsPerBox == 0 && ew / kWarpsPerBox < BLOCK_N / 64) ? ew / kWarpsPerBox : -1;
    // BN statistics: thread -> (16-byte chunk = 8 channels, group of kRowsPer pixel rows) of the staged tile
    constexpr int kChunks = BLOCK_N / 8;
    constexpr int kRowsPer = kBlockM / (kEpiThreads / kChunks);   // 16 / 8 / 4 rows for N = 256 / 128 / 64
    const int sch = et % kChunks, rg = et / kChunks;
    const uint32_t st_boxoff = (uint32_t)(sch >> 3) * kABytes;
    const uint32_t st_chunk = (uint32_t)(sch & 7);
    uint64_t sa2[4], sq2[4];   // packed fp32 pairs: per-channel sum and second statistic of this thread's rows
#pragma unroll
    for (int j = 0; j < 4; ++j) { sa2[j] = 0ull; sq2[j] = 0ull; }
    int lt = 0, rt = 0, nt = 0;   // groups / residual tiles / tiles processed by this CTA
    for (int grp = blockIdx.x; grp < num_groups; grp += gridDim.x, ++lt) {
      const int acc = HALO ? 0 : lt % kAccStages;
      const uint32_t aph = HALO ? (lt & 1) : (lt / kAccStages) & 1;
#pragma unroll 1
      for (int t = 0; t < TILES; ++t) {
      const int tile = grp * TILES + t;
      if (tile >= num_tiles) break;      // uniform over the CTA
      const uint32_t acc_col = (uint32_t)((acc * TILES + t) * BLOCK_N);
      const uint32_t out0 = outbase + (uint32_t)(nt & (OUT_BUFS - 1)) * kOutBytes;   // staging buffers alternate tile by tile
      const uint32_t st_box = out0 + st_boxoff;
      ++nt;
      const int p0 = tile * kBlockM;
      const int pix = p0 + row;
      const bool row_ok = pix < p.M_total;
      // the bulk store issued two tiles ago must be done reading this staging buffer (the previous tile's
      // store may still be draining from the other one); then the residual tile (if any) is fetched into
      // it while this tile's MMAs may still be running
      // (box g of every tile is stored by lane 0 of epilogue warp 2g, which also owns that bulk group: one thread
      // issuing all boxes delayed its whole warp -- and with it the tile -- by 0.35 us at N = 256)
      if (store_box >= 0) { if (OUT_BUFS == 2) tma_store_wait_read1(); else tma_store_wait_read(); }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      // BatchNorm-backward statistics (dgrad): the first batch of this thread's y rows is requested NOW; the L2 round trip used
      // to start only inside the statistics pass, fully exposed once per tile (needs the re-partitioned registers)
      constexpr int kStatBatch = kRowsPer < 8 ? kRowsPer : (EPI == 16 ? 4 : 8);   // rows in flight per thread (bounds register use)
      constexpr bool kHoistY = kRegSplit;
      uint4 yh[kHoistY ? kStatBatch : 1];
      const bool hoist = kHoistY && p.hoist_y && p.bn_y && p.stats && sch * 8 < p.Cout;
      if (hoist) {
        int rows_h = p.M_total - p0;
        if (rows_h > kBlockM) rows_h = kBlockM;
#pragma unroll
        for (int i = 0; i < (kHoistY ? kStatBatch : 1); ++i) {
          const int r = rg * kRowsPer + i;
          yh[i] = make_uint4(0, 0, 0, 0);
          if (r < rows_h) yh[i] = __ldg(reinterpret_cast<const uint4*>(p.bn_y + (size_t)(p0 + r) * p.ldc) + sch);
        }
      }
      if (p.bn_y && et == 32) {
        // BatchNorm-backward statistics stream y from global memory: pull the NEXT tile's boxes into L2 now
        // (and this tile's, the first time) so those loads are L2 hits when the statistics pass issues them
        for (int ahead = (nt == 1 ? 0 : 1); ahead <= 1; ++ahead) {
          const int tl = (ahead == 0) ? tile : ((t + 1 < TILES && tile + 1 < num_tiles) ? tile + 1 : (grp + (int)gridDim.x) * TILES);
          if (tl < num_tiles) {
            const int pp = tl * kBlockM, nn = pp / p.HW, yy0 = (pp - nn * p.HW) / p.W;
            for (int g = 0; g < BLOCK_N / 64; ++g)
              if (g * 64 < p.Cout) tma_prefetch_4d(&tmY, g * 64, 0, yy0, nn);
          }
        }
      }
      if (p.res1_tma && et == 64) {
        // same for the residual tile: the NEXT tile's boxes travel HBM -> L2 while this tile is processed, so the
        // TMA fetch into the (single) staging buffer at the top of the next tile is an L2 hit
        const int tl = (t + 1 < TILES && tile + 1 < num_tiles) ? tile + 1 : (grp + (int)gridDim.x) * TILES;
        if (tl < num_tiles) {
          const int pp = tl * kBlockM, nn = pp / p.HW, yy0 = (pp - nn * p.HW) / p.W;
          for (int g = 0; g < BLOCK_N / 64; ++g)
            if (g * 64 < p.Cout) tma_prefetch_4d(&tmR, g * 64, 0, yy0, nn);
        }
      }
      if (p.res1_tma) {
        if (et == 0) {
          const int n0 = p0 / p.HW;
          const int y0 = (p0 - n0 * p.HW) / p.W;
          int nbox = 0;
          for (int g = 0; g < BLOCK_N / 64; ++g) nbox += (g * 64 < p.Cout);
          mbar_expect_tx(resbar, nbox * kABytes);
          for (int g = 0; g < BLOCK_N / 64; ++g)
            if (g * 64 < p.Cout) tma_load_4d(out0 + (uint32_t)g * kABytes, &tmR, resbar, g * 64, 0, y0, n0);
        }
      }
      if (et == 0 && (nt == 4 || nt == 5)) KT(13 + (nt - 4) * 5);
      mbar_wait(tfull0 + 8 * (HALO ? t : acc), aph);
      tc_fence_after();
      if (et == 0 && nt == 1) KT(6);
      if (et == 0 && (nt == 4 || nt == 5)) KT(14 + (nt - 4) * 5);
      if (p.res1_tma) { mbar_wait(resbar, rt & 1); ++rt; }
      // one 32-column chunk of this thread's row: bias / ReLU / residuals -> bf16 -> swizzled staging box
      auto stage_chunk = [&](const uint32_t (&v)[32], int g) {
        const int n0c = g * 64 + hsel * 32;
        const size_t off = (size_t)pix * p.ldc + n0c;
        // staging: 128-pixel x 64-channel boxes in the 128-byte-swizzled layout TMA expects
        // (16-byte chunk index XOR (row mod 8)); a TMA-fetched residual sits at the very same addresses
        const uint32_t box = out0 + (uint32_t)g * kABytes + (uint32_t)row * 128u;
        if (!p.res1 && !p.res2 && !post_bn) {
          // no residual: bias on packed fp32 pairs, ReLU on the packed bf16 pairs -- 3 instructions per 2 channels
          // the 32 bias values of the chunk are fetched up front (8 x 16 bytes, broadcast): issued back to back, their
          // latency is paid once instead of in front of every packed add
          // (with 16 epilogue warps the register budget is 96: the bias is fetched in two halves)
          constexpr int kBiasParts = EPI == 16 ? 2 : 1;
          const uint32_t badr = smem_u32(s_bias + n0c);
#pragma unroll
          for (int part = 0; part < kBiasParts; ++part) {
            constexpr int kPer = 16 / kBiasParts;      // packed pairs per part
            uint64_t b2[kPer];
#pragma unroll
            for (int j = 0; j < kPer / 2; ++j)
              asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(b2[2 * j]), "=l"(b2[2 * j + 1]) : "r"(badr + 16u * (j + part * (kPer / 2))));
#pragma unroll
            for (int j4l = 0; j4l < kPer / 4; ++j4l) {
              const int j4 = j4l + part * (kPer / 4);
              uint32_t w[4];
#pragma unroll
              for (int j = 0; j < 4; ++j)
                w[j] = cvt_bf16x2(add_f32x2(pack_f32x2(v[8 * j4 + 2 * j], v[8 * j4 + 2 * j + 1]), b2[4 * j4l + j]), p.relu != 0);
              const uint32_t dst = box + ((((uint32_t)hsel * 4u + j4) ^ ((uint32_t)row & 7u)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                           : "memory");
            }
          }
          return;
        }
        if (p.res1_tma && !p.res2 && !p.relu) {
          // the dgrad shape (skip-gradient accumulate): residual tile already in the staging box, packed adds
          const uint64_t* b2 = reinterpret_cast<const uint64_t*>(s_bias + n0c);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint32_t adr = box + ((((uint32_t)hsel * 4u + j4) ^ ((uint32_t)row & 7u)) << 4);
            uint32_t u[4], w[4];
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "r"(adr));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint64_t x2 = pack_f32x2(v[8 * j4 + 2 * j], v[8 * j4 + 2 * j + 1]);
              if (p.bias) x2 = add_f32x2(x2, b2[4 * j4 + j]);
              x2 = add_f32x2(x2, pack_f32x2(u[j] << 16, u[j] & 0xffff0000u));
              w[j] = cvt_bf16x2(x2, false);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(adr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                         : "memory");
          }
          return;
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + s_bias[n0c + j];
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (post_bn) {   // inference: the BatchNorm that follows is a per-channel affine map of this accumulator
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = f[j] * s_osc[n0c + j] + s_osh[n0c + j];
        }
        if (p.res1_tma) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint32_t src = box + ((((uint32_t)hsel * 4u + j4) ^ ((uint32_t)row & 7u)) << 4);
            uint4 u;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(src));
            float r[8];
            unpack_bf16x8(u, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[8 * j4 + j] += r[j];
          }
        } else if (p.res1 && row_ok) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            float r[8];
            unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(p.res1 + off) + j8), r);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[8 * j8 + j] += r[j];
          }
        }
        if (p.res2 && row_ok) {
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            float r[8];
            unpack_bf16x8(__ldg(reinterpret_cast<const uint4*>(p.res2 + off) + j8), r);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[8 * j8 + j] += r[j];
          }
        }
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          uint32_t w[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[8 * j4 + 2 * j], f[8 * j4 + 2 * j + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          const uint32_t dst = box + ((((uint32_t)hsel * 4u + j4) ^ ((uint32_t)row & 7u)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                       : "memory");
        }
      };
      constexpr int kBoxes = BLOCK_N / 64;
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTA2) mbar_arrive_cluster(mapa_u32(tempty0 + 8 * t, 0));   // the leader issues the pair's MMAs
          else mbar_arrive(tempty0 + 8 * (HALO ? t : acc));
        }
      };
      const bool last_read = HALO || t == TILES - 1 || tile + 1 >= num_tiles;   // this tile ends the accumulator stage's use
      if constexpr (EPI == 8) {
        // two chunks per TMEM round trip: the second load is in flight while the first is converted
#pragma unroll 1
        for (int g = 0; g < kBoxes; g += 2) {
          if (g * 64 >= p.Cout) break;  // warp-uniform (Cout is a multiple of 64: whole boxes)
          const bool two = kBoxes > 1 && (g + 1) * 64 < p.Cout;
          uint32_t va[32], vb[32];
          const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + acc_col + (uint32_t)(g * 64 + hsel * 32);
          tmem_ld_32x32(t0, va);
          if (two) tmem_ld_32x32(t0 + 64u, vb);
          tmem_ld_wait();
          // The last columns of the accumulator are in registers: hand it back BEFORE staging, not after.  The four
          // accumulators of a strip-kernel group complete together, and the MMA warp needs accumulator t of the NEXT group
          // after t/6 of a tile time -- long before the epilogue is through with tiles 0..t-1 (forward 3x3 284 -> 257 us).
          if (p.early_release && last_read && (g + 2 >= kBoxes || (g + 2) * 64 >= p.Cout)) release_acc();
          stage_chunk(va, g);
          if (two) stage_chunk(vb, g + 1);
        }
        if (!p.early_release && last_read) release_acc();
      } else {
        // four warps per scheduler: one chunk at a time (96 registers), the other warps hide the TMEM round trip
#pragma unroll 1
        for (int g = gsel; g < kBoxes; g += 2) {
          if (g * 64 >= p.Cout) break;
          uint32_t va[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc_col + (uint32_t)(g * 64 + hsel * 32), va);
          tmem_ld_wait();
          stage_chunk(va, g);
        }
        if (last_read) release_acc();
      }
      fence_proxy_async();                            // generic-proxy smem writes -> visible to the TMA engine
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");  // the eight epilogue warps: tile fully staged
      if (et == 0 && nt == 1) KT(7);
      if (et == 0 && (nt == 4 || nt == 5)) KT(15 + (nt - 4) * 5);
      if (store_box >= 0) {
        const int n0 = p0 / p.HW;
        const int y0 = (p0 - n0 * p.HW) / p.W;
        if (p.single_store) {
          for (int g = 0; g < BLOCK_N / 64; ++g)
            if (g * 64 < p.Cout) tma_store_4d(&tmC, out0 + (uint32_t)g * kABytes, g * 64, 0, y0, n0);
        } else if (store_box * 64 < p.Cout) {
          tma_store_4d(&tmC, out0 + (uint32_t)store_box * kABytes, store_box * 64, 0, y0, n0);
        }
        tma_store_commit();
        if (et == 0 && nt == 1) KT(8);
        if (et == 0 && (nt == 4 || nt == 5)) KT(16 + (nt - 4) * 5);
      }
      if (p.stats && sch * 8 < p.Cout) {
        // per-channel sums of the values as stored (bf16), read back from the staged tile with one 16-byte
        // shared load per pixel row (a warp covers whole rows: conflict-free); second statistic is either
        // sum(v^2) (BatchNorm forward) or sum(v * y) with y streamed from global memory (BatchNorm backward).
        // All rows of a batch are issued before they are consumed; the sums run on packed fp32 pairs
        // (channel 2i in the low lane, 2i+1 in the high lane: exactly the two halves of a bf16x2 word).
        int rows = p.M_total - p0;
        if (rows > kBlockM) rows = kBlockM;
        constexpr int kBatch = kStatBatch;
#pragma unroll 1
        for (int rb = rg * kRowsPer; rb < (rg + 1) * kRowsPer; rb += kBatch) {
          uint4 vv[kBatch], yy[kBatch];
          if (p.bn_y) {
            if (kHoistY && hoist && rb == rg * kRowsPer) {
#pragma unroll
              for (int i = 0; i < kBatch; ++i) yy[i] = yh[kHoistY ? i : 0];
            } else {
#pragma unroll
              for (int i = 0; i < kBatch; ++i) {
                const int r = rb + i;
                yy[i] = make_uint4(0, 0, 0, 0);
                if (r < rows) yy[i] = __ldg(reinterpret_cast<const uint4*>(p.bn_y + (size_t)(p0 + r) * p.ldc) + sch);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < kBatch; ++i) {
            const int r = rb + i;
            const uint32_t src = st_box + (uint32_t)r * 128u + ((st_chunk ^ ((uint32_t)r & 7u)) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(vv[i].x), "=r"(vv[i].y), "=r"(vv[i].z), "=r"(vv[i].w) : "r"(src));
          }
          // rows past the end of the tensor (last tile only) hold bias/ReLU of zero-filled pixels: skipped
          const int nvalid = rows - rb >= kBatch ? kBatch : (rows - rb > 0 ? rows - rb : 0);
          if (p.bn_y) {
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
              if (i < nvalid) {
                const uint32_t w[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
                const uint32_t u[4] = {yy[i].x, yy[i].y, yy[i].z, yy[i].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t v2 = pack_f32x2(w[k] << 16, w[k] & 0xffff0000u);
                  sa2[k] = add_f32x2(sa2[k], v2);
                  sq2[k] = fma_f32x2(v2, pack_f32x2(u[k] << 16, u[k] & 0xffff0000u), sq2[k]);
                }
              }
            }
          } else if (nvalid == kBatch) {
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
              const uint32_t w[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t v2 = pack_f32x2(w[k] << 16, w[k] & 0xffff0000u);
                sa2[k] = add_f32x2(sa2[k], v2);
                sq2[k] = fma_f32x2(v2, v2, sq2[k]);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < kBatch; ++i) {
              if (i < nvalid) {
                const uint32_t w[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t v2 = pack_f32x2(w[k] << 16, w[k] & 0xffff0000u);
                  sa2[k] = add_f32x2(sa2[k], v2);
                  sq2[k] = fma_f32x2(v2, v2, sq2[k]);
                }
              }
            }
          }
        }
      }
      if (et == 0 && (nt == 4 || nt == 5)) KT(17 + (nt - 4) * 5);
      }  // tiles of the group
    }
    if (et == 0) KT(9);
    if (store_box >= 0) tma_store_wait_read();        // smem must stay valid until the last bulk store has read it
    if (et == 0) KT(10);
    if (BNB && p.stats) asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");   // the scratch is the staging buffer: every store has read it
    if (p.stats) {
      // the pipeline stages are idle now: stage 0 doubles as the cross-thread reduction scratch.  Every thread
      // parks its 16 partial sums, then one thread per channel adds the row groups up -- no shared-memory float
      // atomics (they compile to compare-and-swap spin loops: 2.5-3 us per CTA) -- and issues ONE global
      // reduction per channel per CTA for the whole layer.
      constexpr int kRowGroups = kEpiThreads / kChunks;
      if (sch * 8 < p.Cout) {
        float sa[8], sq[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { unpack_f32x2(sa2[j], sa[2 * j], sa[2 * j + 1]); unpack_f32x2(sq2[j], sq[2 * j], sq[2 * j + 1]); }
        float4* d0 = reinterpret_cast<float4*>(s_stats + (size_t)rg * 2 * BLOCK_N + sch * 8);
        float4* d1 = reinterpret_cast<float4*>(s_stats + (size_t)rg * 2 * BLOCK_N + BLOCK_N + sch * 8);
        d0[0] = make_float4(sa[0], sa[1], sa[2], sa[3]); d0[1] = make_float4(sa[4], sa[5], sa[6], sa[7]);
        d1[0] = make_float4(sq[0], sq[1], sq[2], sq[3]); d1[1] = make_float4(sq[4], sq[5], sq[6], sq[7]);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      for (int i = et; i < 2 * BLOCK_N; i += kEpiThreads) {
        const int stat = i / BLOCK_N, col = i - stat * BLOCK_N;
        if (col < p.Cout) {
          float acc = 0.f;
#pragma unroll
          for (int g = 0; g < kRowGroups; ++g) acc += s_stats[g * 2 * BLOCK_N + i];
          atomicAdd(p.stats + (size_t)stat * p.Cout + col, acc);
        }
      }
    }
    if (et == 0) KT(11);
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_arrive_wait();   // both CTAs are done with the pair's tensor memory and with each other's barriers
  if (warp == 1) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_pair(tmem_base, kAccStages * TILES * BLOCK_N);
    else tmem_dealloc(tmem_base, kAccStages * TILES * BLOCK_N);
  }
  if (threadIdx.x == 32) KT(12);
}

// ------------------------------------------------------------------------------------------
// Weight gradient.  D[cout 128][cin BLOCK_N] accumulates over the CTA's share of the pixel tiles;
// both operands are "MN-major" (the reduction axis -- pixels -- is the strided one), which the
// tensor core reads directly from the same swizzled [pixel][64 channel] boxes TMA delivers.
// ------------------------------------------------------------------------------------------
struct WgradKernelParams {
  int M_tiles, H, W, HW;
  int Cin_valid, Cout, ldw;  // valid input channels, valid output channels, row pitch of dw
  int cin_tiles;
  int tap3;
  int tiles_per_split;
  int swap_lbo_sbo;  // debug knob
  int vec4;          // dw rows are 16-byte aligned: vector reductions allowed
  int group_fast;    // > 0: blockIdx.x = (tap, channel-tile) group index (group_fast = groups per cout tile row), blockIdx.y = split.
                     // CTAs that stream the SAME pixel range (other taps / channel tiles) are then neighbours in launch order, run at
                     // the same time and meet in L2 -- with the split index fastest they ran a wave apart and every operand byte
                     // came from DRAM once per group (ncu: 3x3 weight gradient 1.09 GB read for 0.54 GB of operands)
  float* dw;
  BnInput bn_in;     // gamma != null: x is normalised in shared memory (1x1 only)
};

// MT = 128-row output-channel tiles per CTA (1 or 2, one TMEM accumulator each).  The 1x1 convolutions of the
// bottleneck are HBM-bound, so the CTA tile is made as wide as the layer -- <BLOCK_N 256, MT 1> for 256->128,
// <128, 2> for 128->256 -- and every activation byte is fetched once instead of once per tile column/row.
template <int BLOCK_N, int MT, int STAGES>
__global__ void __launch_bounds__(kWgradThreads) conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                              const __grid_constant__ CUtensorMap tmX,
                                                              const WgradKernelParams p) {
  constexpr int kBBytes = (BLOCK_N / 64) * kABytes;
  constexpr int kDyBytes = MT * 2 * kABytes;
  constexpr int kStageBytes = kDyBytes + kBBytes;
  constexpr int kCols = MT * BLOCK_N < 32 ? 32 : MT * BLOCK_N;
  static_assert(kCols <= 512, "TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t bar0 = base + STAGES * kStageBytes;   // full[S] | empty[S] | tmem_full | ready[S]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + STAGES * kStageBytes + (3 * STAGES + 1) * 8);
  float* s_sc = reinterpret_cast<float*>(tmem_slot + 4);   // [256] scale | [256] shift of a deferred input BatchNorm
  float* s_sh = s_sc + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = p.group_fast ? blockIdx.y : blockIdx.x;
  const int by = p.group_fast ? (int)blockIdx.x % p.group_fast : (int)blockIdx.y;
  const int tap = by / p.cin_tiles, cin_tile = by - tap * p.cin_tiles;
  const int cout_tile = p.group_fast ? (int)blockIdx.x / p.group_fast : (int)blockIdx.z;   // in units of MT * 128 output channels
  const bool xform = p.bn_in.gamma != nullptr;
  const int t_begin = split * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split;
  if (t_end > p.M_tiles) t_end = p.M_tiles;
  const int nkb = t_end - t_begin;  // may be <= 0 for trailing splits
  pdl_trigger();

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < 2 * STAGES + 1; ++s) mbar_init(bar0 + 8 * s, 1);
    for (int s = 0; s < STAGES; ++s) mbar_init(bar0 + 8 * (2 * STAGES + 1 + s), 2);   // ready: two transform warps
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), kCols);
    tmem_relinquish();
  }
  pdl_wait();
  if (xform) bn_input_setup(p.bn_in, s_sc, s_sh, threadIdx.x, kWgradThreads, false);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t full0 = bar0, empty0 = bar0 + 8 * STAGES, tfull = bar0 + 16 * STAGES, ready0 = tfull + 8;

  if (nkb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int dy = 0, dx = 0;
        if (p.tap3) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
        for (int kb = 0; kb < nkb; ++kb) {
          const int s = kb % STAGES;
          const uint32_t ph = (kb / STAGES) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const int p0 = (t_begin + kb) * kBlockM;
          const int n0 = p0 / p.HW;
          const int y0 = (p0 - n0 * p.HW) / p.W;
          const uint32_t sa = base + s * kStageBytes;
          mbar_expect_tx(full0 + 8 * s, kStageBytes);
#pragma unroll
          for (int i = 0; i < 2 * MT; ++i)   // boxes past the (padded) channel count are zero-filled by TMA
            tma_load_4d(sa + i * kABytes, &tmDY, full0 + 8 * s, cout_tile * (128 * MT) + i * 64, 0, y0, n0);
#pragma unroll
          for (int i = 0; i < BLOCK_N / 64; ++i)
            tma_load_4d(sa + kDyBytes + i * kABytes, &tmX, full0 + 8 * s, cin_tile * BLOCK_N + i * 64, dx, y0 + dy, n0);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);
        // MN-major, 128-byte swizzle: 64 channels contiguous (one 128-byte row per pixel), groups of
        // 8 pixels every 1024 bytes (SBO), the next 64-channel chunk one whole box later (LBO).
        const uint32_t lbo = p.swap_lbo_sbo ? 1024u : (uint32_t)kABytes;
        const uint32_t sbo = p.swap_lbo_sbo ? (uint32_t)kABytes : 1024u;
        for (int kb = 0; kb < nkb; ++kb) {
          const int s = kb % STAGES;
          const uint32_t ph = (kb / STAGES) & 1;
          mbar_wait((xform ? ready0 : full0) + 8 * s, ph);
          tc_fence_after();
          const uint32_t sa = base + s * kStageBytes;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {  // 16 pixels per MMA
              const uint64_t adesc = make_smem_desc_sw128(sa + mt * 2 * kABytes + k * 2048, lbo, sbo);
              const uint64_t bdesc = make_smem_desc_sw128(sa + kDyBytes + k * 2048, lbo, sbo);
              umma_bf16(tmem_base + (uint32_t)(mt * BLOCK_N), adesc, bdesc, idesc, (kb | k) != 0);
            }
          }
          umma_commit(empty0 + 8 * s);
        }
        umma_commit(tfull);
      }
    } else if (warp >= 6) {
      // operand transform: x boxes of every stage are normalised in place (deferred BatchNorm of the conv input)
      if (xform) {
        const int tw = threadIdx.x - 192;
        for (int kb = 0; kb < nkb; ++kb) {
          const int s = kb % STAGES;
          mbar_wait(full0 + 8 * s, (kb / STAGES) & 1);
          const uint32_t sx = base + s * kStageBytes + kDyBytes;
#pragma unroll
          for (int i = 0; i < BLOCK_N / 64; ++i)
            bn_input_transform_box(sx + i * kABytes, s_sc, s_sh, cin_tile * BLOCK_N + i * 64, tw);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(ready0 + 8 * s);
        }
      }
    } else {
      const int q = warp & 3;
      mbar_wait(tfull, 0);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int row = cout_tile * (128 * MT) + mt * 128 + q * 32 + lane;  // output channel
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * BLOCK_N + c * 32), v);
          tmem_ld_wait();
          const int col0 = cin_tile * BLOCK_N + c * 32;
          if (row < p.Cout && col0 < p.Cin_valid) {
            float* d = p.dw + (size_t)row * p.ldw + (size_t)tap * p.Cin_valid + col0;
            if (col0 + 32 <= p.Cin_valid && p.vec4) {
#pragma unroll
              for (int j = 0; j < 8; ++j)   // 16-byte vector reductions: a quarter of the atomic instructions
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4 * j), "f"(__uint_as_float(v[4 * j])),
                             "f"(__uint_as_float(v[4 * j + 1])), "f"(__uint_as_float(v[4 * j + 2])),
                             "f"(__uint_as_float(v[4 * j + 3])) : "memory");
            } else if (col0 + 32 <= p.Cin_valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) atomicAdd(d + j, __uint_as_float(v[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.Cin_valid) atomicAdd(d + j, __uint_as_float(v[j]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kCols);
  }
}

// ------------------------------------------------------------------------------------------
// 3x3 weight gradient, 128 -> 128 channels, one CTA per (horizontal tap offset dx, range of pixel tiles).
// The per-tap kernel above re-reads dy and x from L2 for each of the nine taps (128 bytes per SM per clock:
// L2-bandwidth bound at half the tensor rate).  Here the three vertical taps of one dx share everything: the
// dy tile is fetched once and multiplied with three row-shifted windows of a ROLLING strip of x (one new
// 128-pixel box per tile, rows y0-1 ... of the image, shifted by dx through the tensor map; out-of-image rows
// and columns arrive as zeros = 'same' padding).  Windows never need to be contiguous across ring slots because
// every MMA covers 16 pixels of one image row and gets its own descriptor.  42 bytes per SM per clock.
// ------------------------------------------------------------------------------------------
constexpr int kW3XSlots = 4, kW3DSlots = 3;

__global__ void __launch_bounds__(kThreads) conv_wgrad3_kernel(const __grid_constant__ CUtensorMap tmDY,
                                                               const __grid_constant__ CUtensorMap tmX,
                                                               const WgradKernelParams p) {
  constexpr int kXRing = kW3XSlots * kABytes;           // one ring per 64-channel half of x
  constexpr int kSmemX = 2 * kXRing, kSmemD = kW3DSlots * 2 * kABytes;
  constexpr int kNumBars = 2 * kW3XSlots + 2 * kW3DSlots + 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t ringX = base, ringD = base + kSmemX;
  const uint32_t bar0 = base + kSmemX + kSmemD;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmemX + kSmemD + kNumBars * 8);
  const uint32_t fullX = bar0, emptyX = bar0 + 8 * kW3XSlots, fullD = bar0 + 16 * kW3XSlots, emptyD = fullD + 8 * kW3DSlots;
  const uint32_t tfull = emptyD + 8 * kW3DSlots;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dxi = p.group_fast ? blockIdx.x : blockIdx.y, dx = dxi - 1;
  const int t_begin = (p.group_fast ? blockIdx.y : blockIdx.x) * p.tiles_per_split;
  int t_end = t_begin + p.tiles_per_split;
  if (t_end > p.M_tiles) t_end = p.M_tiles;
  const int rpt = kBlockM / p.W;            // image rows per 128-pixel tile
  const int tpi = p.H / rpt;                // tiles per image
  pdl_trigger();

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < kNumBars; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);   // three 128-column accumulators
    tmem_relinquish();
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (t_end > t_begin) {
    if (warp == 0) {
      if (lane == 0) {
        int xc = 0;   // x boxes issued so far (ring slot = xc % slots)
        auto load_x = [&](int n, int b) {
          const int s = xc % kW3XSlots;
          mbar_wait(emptyX + 8 * s, ((xc / kW3XSlots) & 1) ^ 1);
          mbar_expect_tx(fullX + 8 * s, 2 * kABytes);
          tma_load_4d(ringX + s * kABytes, &tmX, fullX + 8 * s, 0, dx, b * rpt - 1, n);
          tma_load_4d(ringX + kXRing + s * kABytes, &tmX, fullX + 8 * s, 64, dx, b * rpt - 1, n);
          ++xc;
        };
        for (int T = t_begin, dc = 0; T < t_end; ++T, ++dc) {
          const int n = T / tpi, lt = T - n * tpi;
          if (lt == 0 || T == t_begin) load_x(n, lt);   // leading box of a new image / of this CTA's range
          load_x(n, lt + 1);
          const int s = dc % kW3DSlots;
          mbar_wait(emptyD + 8 * s, ((dc / kW3DSlots) & 1) ^ 1);
          mbar_expect_tx(fullD + 8 * s, 2 * kABytes);
          tma_load_4d(ringD + s * 2 * kABytes, &tmDY, fullD + 8 * s, 0, 0, lt * rpt, n);
          tma_load_4d(ringD + s * 2 * kABytes + kABytes, &tmDY, fullD + 8 * s, 64, 0, lt * rpt, n);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, 128, 1, 1);
        int xc = 0;   // index of the window's LOW box in issue order
        bool first = true;
        for (int T = t_begin, dc = 0; T < t_end; ++T, ++dc) {
          const int n = T / tpi, lt = T - n * tpi;
          if (T != t_begin && lt != 0) ++xc;            // same image: the window slides by one box
          else if (T != t_begin) xc += 2;               // new image: both boxes are new
          const int slo = xc % kW3XSlots, shi = (xc + 1) % kW3XSlots;
          // a box is complete when its fill number matches: box i is the (i / slots)-th fill of slot i % slots
          if (T == t_begin || lt == 0) mbar_wait(fullX + 8 * slo, (xc / kW3XSlots) & 1);
          mbar_wait(fullX + 8 * shi, ((xc + 1) / kW3XSlots) & 1);
          const int sd = dc % kW3DSlots;
          mbar_wait(fullD + 8 * sd, (dc / kW3DSlots) & 1);
          tc_fence_after();
          const uint32_t da = ringD + sd * 2 * kABytes;
#pragma unroll
          for (int dyi = 0; dyi < 3; ++dyi) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // 16 pixels per MMA, always inside one image row of one box
              const int q = dyi * p.W + 16 * k;          // pixel offset of this chunk inside the two-box window
              const uint32_t xs = ringX + (uint32_t)((q >> 7) ? shi : slo) * kABytes + (uint32_t)(q & 127) * 128u;
              const uint64_t adesc = make_smem_desc_sw128(da + k * 2048, (uint32_t)kABytes, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(xs, (uint32_t)kXRing, 1024);
              umma_bf16(tmem_base + (uint32_t)(dyi * 128), adesc, bdesc, idesc, !(first && k == 0));
            }
          }
          first = false;
          // the low box is done unless the next tile of the same image reuses... it never does: the window
          // slides by one box, so the low box is free; the high box stays (it is the next tile's low box)
          umma_commit(emptyX + 8 * slo);
          const bool last_of_image = (lt == tpi - 1) || (T + 1 == t_end);
          if (last_of_image) umma_commit(emptyX + 8 * shi);
          umma_commit(emptyD + 8 * sd);
        }
        umma_commit(tfull);
      }
    } else {
      const int q = warp & 3;
      const int row = q * 32 + lane;  // output channel
      mbar_wait(tfull, 0);
      tc_fence_after();
#pragma unroll 1
      for (int dyi = 0; dyi < 3; ++dyi) {
        const int tap = dyi * 3 + dxi;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(dyi * 128 + c * 32), v);
          tmem_ld_wait();
          float* d = p.dw + (size_t)row * p.ldw + (size_t)tap * 128 + c * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4 * j), "f"(__uint_as_float(v[4 * j])),
                         "f"(__uint_as_float(v[4 * j + 1])), "f"(__uint_as_float(v[4 * j + 2])),
                         "f"(__uint_as_float(v[4 * j + 3])) : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int act_box(int H, int W, ActBox* box) {
  HGB_CHECK_ARG(W > 0 && H > 0 && (W & (W - 1)) == 0 && (H & (H - 1)) == 0 && W <= 128,
                "conv: H and W must be powers of two, W <= 128 (got %dx%d)", H, W);
  box->wb = W;
  box->hb = H < 128 / W ? H : 128 / W;
  box->nb = 128 / (box->wb * box->hb);
  return HGB_OK;
}

int make_tmap_act(CUtensorMap* out, const void* ptr, int N, int H, int W, int C) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)"); return HGB_ERR_CUDA; }
  HGB_CHECK_ARG(C % 8 == 0 && ((uintptr_t)ptr & 15) == 0, "activation tensor: C %% 8 and 16-byte alignment required");
  ActBox b;
  int rc = act_box(H, W, &b);
  if (rc) return rc;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)b.wb, (cuuint32_t)b.hb, (cuuint32_t)b.nb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activation %dx%dx%dx%d) failed: %d", N, H, W, C, (int)r); return HGB_ERR_CUDA; }
  return HGB_OK;
}

int make_tmap_mat(CUtensorMap* out, const void* ptr, int rows, int cols, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)"); return HGB_ERR_CUDA; }
  HGB_CHECK_ARG(cols % 8 == 0 && ((uintptr_t)ptr & 15) == 0 && box_rows <= 256, "weight matrix: bad shape/alignment");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(matrix %dx%d) failed: %d", rows, cols, (int)r); return HGB_ERR_CUDA; }
  return HGB_OK;
}

int conv_gemm_block_n(int Cout) { return Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256); }

static int g_num_sms = 0;

template <int BLOCK_N, int STAGES, int TILES, int OUT_BUFS, bool HALO = false, int EPI = 8, bool BNB = false, bool CTA2 = false>
static int launch_gemm_t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                         const CUtensorMap& tmY, const GemmKernelParams& kp, int tiles_m, int max_ctas, cudaStream_t st,
                         const CUtensorMap* tmZ = nullptr, const CUtensorMap* tmDP = nullptr) {
  constexpr int aslots = TILES + (BLOCK_N == 64 ? 2 : 1);      // (kASlots of the kernel)
  constexpr int ring = HALO ? aslots * kABytes + 6 * (CTA2 ? BLOCK_N / 2 : BLOCK_N) * 128
                            : STAGES * (TILES * kABytes + BLOCK_N * 128 + (BNB ? kABytes : 0));
  constexpr int nbars = HALO ? 2 * aslots + 12 + 2 * TILES + 1 : 3 * STAGES + 5;
  constexpr int smem = ring + OUT_BUFS * (BLOCK_N / 64) * kABytes + (nbars * 8 + 15) / 16 * 16 + 16 + BLOCK_N * 4 +
                       (TILES == 1 ? 4 * 256 * 4 : 0) + (BNB ? 8 * 4 * 64 * 4 : 0) + 1024;   // BatchNorm scale+shift tables only for the 1x1 variants
  static_assert(ring >= (EPI == 16 ? 32 : 16) * 1024, "stats scratch (16 / 32 KB) aliases the pipeline stages");
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool attr_done = false;
  if (!attr_done) {
    HGB_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BLOCK_N, STAGES, TILES, OUT_BUFS, HALO, EPI, BNB, CTA2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  if (!g_num_sms) {
    int dev = 0;
    HGB_CUDA(cudaGetDevice(&dev));
    HGB_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int groups = (tiles_m + TILES - 1) / TILES;
  int grid = groups < g_num_sms ? groups : g_num_sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  if (CTA2) grid &= ~1;     // whole CTA pairs (the caller guarantees an even number of groups: both CTAs of a pair loop alike)
  HGB_CUDA(launch_pdl_cluster(conv_gemm_kernel<BLOCK_N, STAGES, TILES, OUT_BUFS, HALO, EPI, BNB, CTA2>, dim3(grid), dim3(gemm_threads(EPI)), smem,
                              st, CTA2 ? 2 : 1, tmA, tmB, tmC, tmR, tmY, tmZ ? *tmZ : tmC, tmDP ? *tmDP : tmC, kp));
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

bool conv_gemm_supports_bn_bwd(int ksize, int Cin, int Cout) {
  // Cin = channels of dz / y / dp (K of the dgrad GEMM, at most 4 blocks of 64), Cout = the dgrad's output channels
  // (256 -> 256, the head convolution, is excluded: its stages are 64 KB, only two fit, and the fused kernel measured
  //  690 us against 248 + 256 us for the two-pass form at batch 256)
  return ksize == 1 && Cin % 64 == 0 && Cin <= 256 && (Cout == 128 || (Cout == 256 && Cin <= 128));
}

int launch_conv_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap* tmR,
                     const CUtensorMap* tmY, const ConvGemmArgs& a, cudaStream_t st, const CUtensorMap* tmZ,
                     const CUtensorMap* tmDP, const CUtensorMap* tmB64) {
  HGB_CHECK_ARG(a.ksize == 1 || a.ksize == 3, "conv_gemm: kernel size must be 1 or 3");
  HGB_CHECK_ARG(a.Cin % 64 == 0 && a.Cin > 0, "conv_gemm: Cin must be a multiple of 64 (got %d)", a.Cin);
  HGB_CHECK_ARG(a.Cout % 64 == 0 && a.Cout > 0, "conv_gemm: Cout must be a multiple of 64 (got %d)", a.Cout);
  HGB_CHECK_ARG(a.ldc % 8 == 0 && a.ldc >= a.Cout, "conv_gemm: bad output pitch");
  HGB_CHECK_ARG(a.tap_sign == 1 || a.tap_sign == -1, "conv_gemm: tap_sign must be +-1");
  ActBox b;
  int rc = act_box(a.H, a.W, &b);
  if (rc) return rc;
  GemmKernelParams kp;
  kp.M_total = a.N * a.H * a.W;
  kp.H = a.H; kp.W = a.W; kp.HW = a.H * a.W;
  kp.cblk = a.Cin / 64;
  kp.tap3 = a.ksize == 3;
  kp.nkb = kp.cblk * (kp.tap3 ? 9 : 1);
  kp.tap_sign = a.tap_sign;
  kp.Cout = a.Cout; kp.ldc = a.ldc; kp.relu = a.relu;
  kp.bias = a.bias; kp.res1 = a.res1; kp.res2 = a.res2; kp.out = a.out; kp.stats = a.stats;
  if (g_debug[37]) kp.stats = nullptr;      // TIMING EXPERIMENTS ONLY: drop the BatchNorm statistics of the epilogue (wrong results)
  kp.early_release = !g_debug[39];
  kp.hoist_y = !g_debug[41];
  kp.bn_y = a.bn_y;
  kp.bn_in = a.bn_in;
  kp.single_store = g_debug[16];
  kp.late_trigger = !g_debug[18];   // default on: forward pass at batch 32 9.55 -> 8.93 ms (hgb_debug_set(18, 1) = trigger at kernel start)
  kp.bn_out = a.bn_out;
  kp.bn_bwd = a.bn_bwd;
  HGB_CHECK_ARG(a.bn_bwd.gamma == nullptr || (conv_gemm_supports_bn_bwd(a.ksize, a.Cin, a.Cout) && a.bn_bwd.C == a.Cin && tmZ && tmDP &&
                                              a.bn_in.gamma == nullptr && a.bn_out.gamma == nullptr),
                "conv_gemm: the fused BatchNorm backward needs a 1x1 dgrad with <= 256 input and 128 / 256 output channels");
  HGB_CHECK_ARG(a.bn_out.gamma == nullptr || (a.ksize == 1 && a.bn_out.C == a.Cout && a.Cout <= 256 && a.bn_out.mode == 1),
                "conv_gemm: an output BatchNorm needs a 1x1 convolution with Cout <= 256 in inference mode");
  HGB_CHECK_ARG(a.bn_in.gamma == nullptr || (a.ksize == 1 && a.bn_in.C == a.Cin && a.Cin <= 256),
                "conv_gemm: an input BatchNorm needs a 1x1 convolution with Cin <= 256");
  kp.res1_tma = (a.res1 != nullptr && tmR != nullptr && !g_debug[3]) ? 1 : 0;
  const CUtensorMap& tmRr = kp.res1_tma ? *tmR : tmC;
  HGB_CHECK_ARG(a.bn_y == nullptr || tmY != nullptr, "conv_gemm: bn_y needs its tensor map");
  const CUtensorMap& tmYr = a.bn_y ? *tmY : tmC;
  if (kp.M_total == 0) return HGB_OK;
  HGB_CHECK_ARG(a.Cout <= 256, "conv_gemm: Cout must be <= 256 (one N tile per CTA), got %d", a.Cout);
  const int tiles_m = cdiv(kp.M_total, kBlockM);
  if (!g_num_sms) {
    int dev = 0;
    HGB_CUDA(cudaGetDevice(&dev));
    HGB_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // weight-stationary 4-tile groups for the 3x3 convolutions when there are enough groups to fill the chip
  const bool ws = kp.tap3 && !g_debug[5] && tiles_m >= 8 * g_num_sms;
  // strip reuse across the vertical taps: groups of 4 tiles inside one image, at least two image rows per tile
  const int rpt = kBlockM / a.W;
  // (HALO pays off as soon as there is one 4-tile group per SM; g_debug[15] overrides the tile threshold)
  const int halo_min = g_debug[15] > 0 ? g_debug[15] : 4 * g_num_sms;
  const bool halo = kp.tap3 && !g_debug[5] && tiles_m >= halo_min && !g_debug[12] && a.W <= 64 && rpt >= 2 &&
                    a.H % (4 * rpt) == 0 && a.Cout == 128;
  // the 64-channel 3x3 layers of the front module (128x128: one image row per box, three-box windows; 64x64): the per-tap ring
  // fetched every input pixel nine times from L2 and ran at a third of the tensor peak.  hgb_debug_set(45, 1) = per-tap ring.
  const bool halo64 = kp.tap3 && !g_debug[5] && !g_debug[12] && !g_debug[45] && tiles_m >= halo_min && a.Cout == 64 && a.W <= 128 &&
                      a.H % (4 * rpt) == 0 && !a.bn_bwd.gamma;
  if (a.bn_bwd.gamma) {   // 1x1 dgrad with the BatchNorm backward fused in
    if (a.Cout == 128) return launch_gemm_t<128, 3, 1, 2, false, 8, true>(tmA, tmB, tmC, tmRr, tmYr, kp, tiles_m, a.max_ctas, st, tmZ, tmDP);
    // 16 epilogue warps when the epilogue also carries the next BatchNorm's statistics (569 -> 525 us at batch 256);
    // hgb_debug_set(38, 1) = never, 2 = always
    if (g_debug[38] == 2 || (g_debug[38] == 0 && kp.stats))
      return launch_gemm_t<256, 2, 1, 1, false, 16, true>(tmA, tmB, tmC, tmRr, tmYr, kp, tiles_m, a.max_ctas, st, tmZ, tmDP);
    return launch_gemm_t<256, 2, 1, 1, false, 8, true>(tmA, tmB, tmC, tmRr, tmYr, kp, tiles_m, a.max_ctas, st, tmZ, tmDP);
  }
  // CTA pairs (tcgen05.mma.cta_group::2), OPT-IN with hgb_debug_set(30, 1): parity-green (tests/test_gpu_conv.py and the
  // batch-80 replay run it) but measured SLOWER than the one-CTA strip kernel on B200 at batch 256 -- forward 64x64 297.7 vs
  // 260.1 us, dgrad 331.8 vs 299.0 us, 32x32 94.9 vs 84.0 us (profiles/r02_ops_ab_cta_pairs.txt) -- so the one-CTA kernel
  // stays the default.  Needs the weight map with 64-row boxes and an even number of groups (with an even grid both CTAs of
  // a pair then run the same number of iterations).
  if (halo && tmB64 && g_debug[30] == 1 && (tiles_m % 8) == 0 && g_num_sms % 2 == 0 && (a.max_ctas <= 0 || a.max_ctas >= 2))
    return launch_gemm_t<128, 1, 4, 1, true, 8, false, true>(tmA, *tmB64, tmC, tmRr, tmYr, kp, tiles_m, a.max_ctas, st);
  // (16 epilogue warps measured slower in the strip kernel: forward 277 vs 257 us, dgrad 308 vs 294 us)
  if (halo) return launch_gemm_t<128, 1, 4, 1, true>(tmA, tmB, tmC, tmRr, tmYr, kp, tiles_m, a.max_ctas, st);
  if (halo64) return launch_gemm_t<64, 1, 4, 1, true>(tmA, tmB, tmC, tmRr, tmYr, kp, tiles_m, a.max_ctas, st);
  switch (conv_gemm_block_n(a.Cout)) {
    case 64: return ws ? launch_gemm_t<64, 2, 4, 2>(tmA, tmB, tmC, tmRr, tmYr, kp, tiles_m, a.max_ctas, st)
                       : launch_gemm_t<64, 6, 1, 2>(tmA, tmB, tmC, tmRr, tmYr, kp, tiles_m, a.max_ctas, st);
    case 128: return ws ? launch_gemm_t<128, 2, 4, 2>(tmA, tmB, tmC, tmRr, tmYr, kp, tiles_m, a.max_ctas, st)
                        : launch_gemm_t<128, 4, 1, 2>(tmA, tmB, tmC, tmRr, tmYr, kp, tiles_m, a.max_ctas, st);
    default:   // 64 KB staging: single.  16 epilogue warps only where they measured faster (A/B on B200, batch 256 @64x64:
      // dgrad with TMA residual + BatchNorm statistics 382 vs 399 us; plain forward tiles 171 vs 167 us: those are bound
      // by shared-memory traffic -- 448 KB per tile -- not by epilogue issue slots).  hgb_debug_set(25, 1) = always 8.
      // (round 2) 16 warps whenever the epilogue carries BatchNorm statistics -- their pass over the staged tile is shared by
      // twice the threads: 128->256 forward 235 -> 212 us, 256->256 dgrad 359 -> 303 us.  hgb_debug_set(25, 2) = the round-1 rule.
      {
        const bool epi16 = g_debug[25] == 1 ? false : g_debug[25] == 2 ? (kp.res1_tma && a.bn_y) : kp.stats != nullptr;
        return !epi16 ? launch_gemm_t<256, 3, 1, 1>(tmA, tmB, tmC, tmRr, tmYr, kp, tiles_m, a.max_ctas, st)
                      : launch_gemm_t<256, 3, 1, 1, false, 16>(tmA, tmB, tmC, tmRr, tmYr, kp, tiles_m, a.max_ctas, st);
      }
  }
}

template <int BLOCK_N, int MT, int STAGES>
static int launch_wgrad_t(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradKernelParams& kp, dim3 grid, cudaStream_t st) {
  constexpr int smem = STAGES * (2 * MT + BLOCK_N / 64) * kABytes + (3 * STAGES + 1) * 8 + 16 + 2 * 256 * 4 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool attr_done = false;
  if (!attr_done) {
    HGB_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<BLOCK_N, MT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_done = true;
  }
  HGB_CUDA(launch_pdl(conv_wgrad_kernel<BLOCK_N, MT, STAGES>, grid, dim3(kWgradThreads), smem, st, tmDY, tmX, kp));
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

int launch_conv_wgrad(const CUtensorMap& tmDY, const CUtensorMap& tmX, const WgradArgs& a, cudaStream_t st) {
  HGB_CHECK_ARG(a.ksize == 1 || a.ksize == 3, "conv_wgrad: kernel size must be 1 or 3");
  HGB_CHECK_ARG(a.Cin % 64 == 0 && a.Cin > 0, "conv_wgrad: Cin must be a multiple of 64 (got %d)", a.Cin);
  HGB_CHECK_ARG(a.Cout % 32 == 0 && a.Cout > 0, "conv_wgrad: Cout must be a multiple of 32 (got %d)", a.Cout);
  ActBox b;
  int rc = act_box(a.H, a.W, &b);
  if (rc) return rc;
  WgradKernelParams kp;
  const int M_total = a.N * a.H * a.W;
  if (M_total == 0) return HGB_OK;
  kp.M_tiles = cdiv(M_total, kBlockM);
  kp.H = a.H; kp.W = a.W; kp.HW = a.H * a.W;
  const int taps = a.ksize == 3 ? 9 : 1;
  kp.Cin_valid = a.Cin_valid > 0 ? a.Cin_valid : a.Cin;
  kp.Cout = a.Cout_valid > 0 ? a.Cout_valid : a.Cout;
  kp.ldw = taps * kp.Cin_valid;
  kp.tap3 = a.ksize == 3;
  kp.bn_in = a.bn_in;
  HGB_CHECK_ARG(a.bn_in.gamma == nullptr || (a.ksize == 1 && a.bn_in.C == a.Cin && a.Cin <= 256),
                "conv_wgrad: an input BatchNorm needs a 1x1 convolution with Cin <= 256");
  // CTA tile: the HBM-bound 1x1 layers take the widest tile that fits (every activation byte fetched once when the
  // whole layer is one tile); 3x3 layers keep 128 x 128 per tap.  Wide tiles need enough pixel tiles per CTA to
  // amortise their larger fp32 reduction epilogue.
  // 3x3, 128 -> 128, tiles inside one image: the three vertical taps share dy and a rolling strip of x
  const int rpt3 = kBlockM / a.W;
  if (a.ksize == 3 && !g_debug[13] && a.Cin == 128 && a.Cout == 128 && kp.Cin_valid == 128 && kp.Cout == 128 && a.W >= 16 &&
      a.W <= 64 && rpt3 >= 2 && a.H % rpt3 == 0 && kp.M_tiles >= 4 * 148 && (((uintptr_t)a.dw & 15) == 0)) {
    // whole waves of one CTA per SM (3 * splits <= 148 * waves); two waves only when every CTA still gets >= 64 tiles,
    // so the 192 KB fp32 reduction epilogue of a CTA stays a small fraction of its work
    int splits = (2 * 148) / 3;
    if (kp.M_tiles / splits < 64) splits = 148 / 3;
    if (a.max_ctas > 0 && 3 * splits > a.max_ctas) splits = a.max_ctas / 3 > 0 ? a.max_ctas / 3 : 1;
    kp.tiles_per_split = cdiv(kp.M_tiles, splits);
    splits = cdiv(kp.M_tiles, kp.tiles_per_split);
    kp.cin_tiles = 1; kp.swap_lbo_sbo = 0; kp.vec4 = 1; kp.dw = a.dw;
    kp.group_fast = g_debug[34] ? 0 : 3;          // hgb_debug_set(34, 1): the round-1 launch order (split index fastest)
    constexpr int smem = (2 * kW3XSlots + 2 * kW3DSlots) * kABytes + (2 * kW3XSlots + 2 * kW3DSlots + 1) * 8 + 16 + 1024;
    static_assert(smem <= 227 * 1024, "shared memory budget");
    static bool attr_done = false;
    if (!attr_done) {
      HGB_CUDA(cudaFuncSetAttribute(conv_wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_done = true;
    }
    HGB_CUDA(launch_pdl(conv_wgrad3_kernel, kp.group_fast ? dim3(3, splits, 1) : dim3(splits, 3, 1), dim3(kThreads), smem, st, tmDY, tmX, kp));
    HGB_LAUNCH_CHECK();
    return HGB_OK;
  }
  int bn = (a.Cin % 128 == 0) ? 128 : 64, mt = 1;
  if (a.ksize == 1 && g_debug[11] != 1 && kp.M_tiles >= (g_debug[22] > 0 ? g_debug[22] : 8 * 148)) {
    if (a.Cin % 256 == 0 && (a.Cout <= 128 || g_debug[11] == 2)) bn = 256;
    else if (a.Cout % 256 == 0 && a.Cin == 128) mt = 2;   // (256 -> 256 measured slower with wide tiles: two stages only)
  }
  kp.cin_tiles = a.Cin / bn;
  const int cout_tiles = cdiv(a.Cout, 128 * mt);
  const int groups = taps * kp.cin_tiles * cout_tiles;
  // split-K over pixel tiles: enough CTAs to fill the chip, but at least 8 tiles per CTA so the fp32
  // atomic epilogue (a full 128 x N tile per CTA) is amortised -- it dominated the low-resolution levels
  int splits = cdiv(2 * 148, groups);
  if (a.max_ctas > 0 && splits * groups > a.max_ctas) splits = a.max_ctas / groups;
  if (splits > kp.M_tiles / 8) splits = kp.M_tiles / 8;
  if (splits < 1) splits = 1;
  kp.tiles_per_split = cdiv(kp.M_tiles, splits);
  splits = cdiv(kp.M_tiles, kp.tiles_per_split);
  kp.swap_lbo_sbo = g_debug[1];
  kp.vec4 = (kp.ldw % 4 == 0) && (kp.Cin_valid % 4 == 0) && (((uintptr_t)a.dw & 15) == 0);
  kp.dw = a.dw;
  kp.group_fast = (g_debug[34] || groups == 1) ? 0 : taps * kp.cin_tiles;
  dim3 grid(splits, taps * kp.cin_tiles, cout_tiles);
  if (kp.group_fast) grid = dim3(groups, splits, 1);
  if (bn == 256) return launch_wgrad_t<256, 1, 2>(tmDY, tmX, kp, grid, st);
  if (mt == 2) return launch_wgrad_t<128, 2, 2>(tmDY, tmX, kp, grid, st);
  if (bn == 128) return launch_wgrad_t<128, 1, 3>(tmDY, tmX, kp, grid, st);
  return launch_wgrad_t<64, 1, 4>(tmDY, tmX, kp, grid, st);
}

}  // namespace hgb

// ------------------------------------------------------------------------------------------
// Stand-alone C entry points (tensor maps are built per call; the model caches them)
// ------------------------------------------------------------------------------------------
using namespace hgb;

extern "C" int hgb_conv_gemm(const void* in, const void* w, const float* bias, const void* res1, const void* res2, void* out,
                             float* stats, int N, int H, int W, int Cin, int Cout, int ksize, int relu, int ldc,
                             int tap_sign, void* stream) {
  HGB_CHECK_ARG(in && w && out, "hgb_conv_gemm: null pointer");
  HGB_CHECK_ARG(N > 0, "hgb_conv_gemm: empty batch");
  CUtensorMap tmA, tmB, tmC;
  int rc = make_tmap_act(&tmA, in, N, H, W, Cin);
  if (rc) return rc;
  rc = make_tmap_act(&tmC, out, N, H, W, ldc);
  if (rc) return rc;
  const int bn = conv_gemm_block_n(Cout);
  rc = make_tmap_mat(&tmB, w, Cout, ksize * ksize * Cin, bn);
  if (rc) return rc;
  ConvGemmArgs a;
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = ksize; a.tap_sign = tap_sign; a.relu = relu; a.ldc = ldc;
  a.bias = bias; a.res1 = (const __nv_bfloat16*)res1; a.res2 = (const __nv_bfloat16*)res2; a.out = (__nv_bfloat16*)out;
  a.stats = stats;
  a.bn_y = nullptr;
  CUtensorMap tmR;
  if (res1) {
    rc = make_tmap_act(&tmR, res1, N, H, W, ldc);
    if (rc) return rc;
  }
  CUtensorMap tmB64;
  const bool pair_ok = ksize == 3 && Cout == 128;
  if (pair_ok) {
    rc = make_tmap_mat(&tmB64, w, Cout, ksize * ksize * Cin, 64);
    if (rc) return rc;
  }
  return launch_conv_gemm(tmA, tmB, tmC, res1 ? &tmR : nullptr, nullptr, a, (cudaStream_t)stream, nullptr, nullptr, pair_ok ? &tmB64 : nullptr);
}

#ifdef HGB_KTIME
extern "C" int hgb_debug_ktime(long long* out32) {
  return cudaMemcpyFromSymbol(out32, hgb::g_ktime, sizeof(long long) * 32) == cudaSuccess ? 0 : -2;
}
#endif

extern "C" int hgb_conv_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout, int ksize,
                              void* stream) {
  HGB_CHECK_ARG(x && dy && dw, "hgb_conv_wgrad: null pointer");
  HGB_CHECK_ARG(N > 0, "hgb_conv_wgrad: empty batch");
  CUtensorMap tmDY, tmX;
  int rc = make_tmap_act(&tmDY, dy, N, H, W, Cout);
  if (rc) return rc;
  rc = make_tmap_act(&tmX, x, N, H, W, Cin);
  if (rc) return rc;
  WgradArgs a;
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ksize = ksize; a.dw = dw;
  a.Cin_valid = 0; a.Cout_valid = 0;
  return launch_conv_wgrad(tmDY, tmX, a, (cudaStream_t)stream);
}
