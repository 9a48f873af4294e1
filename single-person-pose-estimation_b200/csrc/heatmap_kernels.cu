// Heatmap-path kernels: target rendering, one-pass losses, decode, PCK / OKS.
// All are HBM- or latency-bound byte/float streaming kernels: 16-byte coalesced
// accesses, per-joint state kept in registers by making the per-thread element stride a
// multiple of the joint count (so a thread always sees the same joints), warp-shuffle
// reductions, no tensor cores.
//
// Reference arithmetic (relative to the reference tree):
//   render  : dataset_builder.py:220-235, utilities/data_utils.py:187-211
//   losses  : loss.py:2-36, trainer.py:231-233
//   decode  : utilities/data_utils.py:100-183
//   PCK     : eval.py:62-88 ; OKS: public COCO keypoint similarity (pycocotools computeOks)
#include "common.cuh"
#include "sm100_ptx.cuh"
#include <math_constants.h>

namespace hgb {

// ------------------------------------------------------------------------------------
// Target rendering
// ------------------------------------------------------------------------------------
// float32(exp(-d2/2)) for d2 = dx^2+dy^2, |dx|,|dy| <= 3 -- bit patterns of the values numpy
// produces (float64 exp, rounded on assignment into the float32 map; data_utils.py:202,209).
__constant__ uint32_t c_gauss_lut[19] = {
    0x3f800000u /*0*/, 0x3f1b4598u /*1*/, 0x3ebc5ab2u /*2*/, 0u, 0x3e0a9555u /*4*/, 0x3da81c2eu /*5*/, 0u, 0u,
    0x3c960aaeu /*8*/, 0x3c360282u /*9*/, 0x3bdcc9ffu /*10*/, 0u, 0u, 0x3ac50f0cu /*13*/, 0u, 0u, 0u, 0u,
    0x39016791u /*18*/};

// grid: (chunks, B).  A block owns one CONTIGUOUS range of a sample's float4 vectors.  Nearly every byte of a target map is
// zero, so the block first streams zeros over its range (three instructions per 16 bytes: the first version evaluated the
// 7x7 window test per element and was instruction-issue bound at 0.47 of the HBM peak), then -- after a block barrier,
// which orders the two writes to an address -- the K joints x 49 window values that fall inside the range are stored on
// top (they land in L2 lines that are still dirty from the zero fill: no extra DRAM traffic).  Joint k owns channel k, so
// windows of different joints never touch the same element; within a window every element is written once ('assigned',
// data_utils.py:209-210).
__global__ void __launch_bounds__(256) render_targets_kernel(const float* __restrict__ kx, const float* __restrict__ ky,
                                                             const int32_t* __restrict__ kv, int H, int W, int K,
                                                             int vec_per_sample, int vec_per_block, float* __restrict__ out) {
  extern __shared__ int s_joint[];  // [K][2]: cx, cy  (cx < 0 when the joint is not drawn)
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float fx = kx[(size_t)b * K + k], fy = ky[(size_t)b * K + k];
    const int x = (int)fx, y = (int)fy;  // Python int(): truncate toward zero (dataset_builder.py:229-230)
    const bool ok = (0 < x) && (x < W) && (0 < y) && (y < H) && (kv[(size_t)b * K + k] > 0);
    s_joint[2 * k] = ok ? x : -(1 << 28);
    s_joint[2 * k + 1] = y;
  }
  float* ob = out + (size_t)b * H * W * K;
  float4* o4 = reinterpret_cast<float4*>(ob);
  const int v_begin = blockIdx.x * vec_per_block;
  const int v_end = min(v_begin + vec_per_block, vec_per_sample);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  int v = v_begin + threadIdx.x;
  for (; v + 3 * 256 < v_end; v += 4 * 256) {
    __stcs(o4 + v, zero);
    __stcs(o4 + v + 256, zero);
    __stcs(o4 + v + 512, zero);
    __stcs(o4 + v + 768, zero);
  }
  for (; v < v_end; v += 256) __stcs(o4 + v, zero);
  __syncthreads();
  const int e_begin = v_begin * 4, e_end = v_end * 4;
  for (int i = threadIdx.x; i < K * 49; i += blockDim.x) {
    const int k = i / 49, pos = i - k * 49;
    const int wy = pos / 7, wx = pos - wy * 7;
    const int cx = s_joint[2 * k];
    if (cx < 0) continue;
    const int px = cx + wx - 3, py = s_joint[2 * k + 1] + wy - 3;
    if (px < 0 || px >= W || py < 0 || py >= H) continue;          // the patch is clipped to the image (data_utils.py:204-208)
    const int e = (py * W + px) * K + k;
    if (e < e_begin || e >= e_end) continue;
    const int dx = wx - 3, dy = wy - 3;
    ob[e] = __uint_as_float(c_gauss_lut[dx * dx + dy * dy]);
  }
}

// ------------------------------------------------------------------------------------
// Losses
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
struct Vec4 {};
template <>
struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&r)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&r)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
  }
};
template <>
struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&r)[4]) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&v.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&v.y);
    r[0] = __low2float(a); r[1] = __high2float(a); r[2] = __low2float(b); r[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&r)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(r[0], r[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(r[2], r[3]);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&a);
    v.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = v;
  }
};

// Per-(b,k) sums.  Block = 16*K threads, one block per sample; thread t always owns joints
// (4t + j) % K because the vector stride 4*blockDim is a multiple of K.
//   mode 0: s0 = sum t                        (weighted_keypoint_mse, loss.py:32)
//   mode 1: s0 = sum t*p, s1 = sum t*t, s2 = sum p*p   (IOU, loss.py:25-26)
template <typename TP, int MODE>
__global__ void bk_sums_kernel(const float* __restrict__ yt, const TP* __restrict__ yp, int HWK, int K,
                               float* __restrict__ ws /* [B][K][4] */) {
  extern __shared__ float s_part[];  // [3][S*4]
  const int b = blockIdx.x, t = threadIdx.x, S = 16 * K;  // blockDim = S rounded up to a warp multiple
  const float* tb = yt + (size_t)b * HWK;
  const TP* pb = yp + (size_t)b * HWK;
  float a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
  for (int v = t; v < HWK / 4 && t < S; v += S) {
    float tv[4], pv[4];
    Vec4<float>::load(tb + 4 * (size_t)v, tv);
    if (MODE == 1) Vec4<TP>::load(pb + 4 * (size_t)v, pv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (MODE == 0) {
        a0[j] += tv[j];
      } else {
        a0[j] += tv[j] * pv[j];
        a1[j] += tv[j] * tv[j];
        a2[j] += pv[j] * pv[j];
      }
    }
  }
  if (t < S) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s_part[0 * 4 * S + 4 * t + j] = a0[j];
      if (MODE == 1) {
        s_part[1 * 4 * S + 4 * t + j] = a1[j];
        s_part[2 * 4 * S + 4 * t + j] = a2[j];
      }
    }
  }
  __syncthreads();
  // candidate c = 4t+j belongs to joint c % K; 64 candidates per joint
  const int warp = t >> 5, lane = t & 31, nwarp = blockDim.x >> 5;
  for (int k = warp; k < K; k += nwarp) {
    double r0 = 0, r1 = 0, r2 = 0;
    for (int m = lane; m < 64; m += 32) {
      const int c = k + K * m;
      r0 += (double)s_part[c];
      if (MODE == 1) {
        r1 += (double)s_part[4 * S + c];
        r2 += (double)s_part[8 * S + c];
      }
    }
    r0 = warp_sum_d(r0);
    if (MODE == 1) { r1 = warp_sum_d(r1); r2 = warp_sum_d(r2); }
    if (lane == 0) {
      float* o = ws + ((size_t)b * K + k) * 4;
      if (MODE == 0) {
        o[0] = (float)r0;
      } else {
        // IoU pieces: a = 1/(U+eps), c = (I+eps)/(U+eps)^2, iou
        const double eps = 1e-7;
        const double U = r1 + r2 - r0;
        const double a = 1.0 / (U + eps);
        const double iou = (r0 + eps) * a;
        o[0] = (float)a;
        o[1] = (float)(iou * a);
        o[2] = (float)iou;
      }
    }
  }
}

// One pass: loss partial sums + gradient.  kind 0 weighted_mse, 1 mse, 2 iou, 3 keypoint mse.
template <typename TP, typename TG, int KIND>
__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ yt, const TP* __restrict__ yp,
                                                        TG* __restrict__ grad, const float* __restrict__ ws,
                                                        int64_t nvec, int HWK, int K, double inv_count,
                                                        double* __restrict__ loss_acc) {
  const float gscale = (float)(2.0 * inv_count);
  const float iou_gscale = (float)inv_count;
  double part = 0.0;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    float tv[4], pv[4], gv[4];
    Vec4<float>::load(yt + 4 * v, tv);
    Vec4<TP>::load(yp + 4 * v, pv);
    int b = 0, k = 0;
    if (KIND >= 2) {
      const int64_t e0 = 4 * v;
      b = (int)(e0 / HWK);
      k = (int)((e0 - (int64_t)b * HWK) % K);
    }
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = pv[j] - tv[j];
      if (KIND == 0) {
        const float w = tv[j] > 0.f ? 82.f : 1.f;  // loss.py:14
        acc += w * d * d;
        gv[j] = gscale * w * d;
      } else if (KIND == 1) {
        acc += d * d;
        gv[j] = gscale * d;
      } else if (KIND == 3) {
        const float kw = ws[((size_t)b * K + k) * 4] == 0.f ? 0.f : 1.f;  // loss.py:34
        acc += kw * d * d;
        gv[j] = gscale * kw * d;
      } else {  // IoU: d iou/dp = t*a - c*(2p - t); loss = sum(1 - iou)/(B*K)
        const float* c = ws + ((size_t)b * K + k) * 4;
        gv[j] = -iou_gscale * (tv[j] * c[0] - c[1] * (2.f * pv[j] - tv[j]));
      }
      if (KIND >= 2) {
        if (++k == K) k = 0;  // HWK % 4 == 0 so a vector never crosses a sample
      }
    }
    part += (double)acc;
    if (grad) Vec4<TG>::store(grad + 4 * v, gv);
  }
  if (KIND == 2) return;  // IoU loss scalar is summed by iou_loss_sum_kernel
  part = warp_sum_d(part);
  __shared__ double s_w[8];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += s_w[i];
    atomicAdd(loss_acc, s * inv_count);
  }
}

__global__ void iou_loss_sum_kernel(const float* __restrict__ ws, int BK, double inv_count, double* loss_acc) {
  double part = 0;
  for (int i = threadIdx.x; i < BK; i += blockDim.x) part += 1.0 - (double)ws[(size_t)i * 4 + 2];
  part = warp_sum_d(part);
  __shared__ double s_w[32];
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += s_w[i];
    atomicAdd(loss_acc, s * inv_count);
  }
}

// (B,H,W) map the reference loss functions return (mean over the joint axis).
template <int KIND>
__global__ void loss_map_kernel(const float* __restrict__ yt, const float* __restrict__ yp, const float* __restrict__ ws,
                                int64_t npix, int HW, int K, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int b = (int)(i / HW);
  const float* t = yt + i * K;
  const float* p = yp + i * K;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const float d = t[k] - p[k];
    float w = 1.f;
    if (KIND == 0) w = t[k] > 0.f ? 82.f : 1.f;
    if (KIND == 3) w = ws[((size_t)b * K + k) * 4] == 0.f ? 0.f : 1.f;
    acc += d * d * w;
  }
  out[i] = acc / (float)K;
}

__global__ void iou_vec_kernel(const float* __restrict__ ws, int B, int K, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += ws[((size_t)b * K + k) * 4 + 2];
  out[b] = 1.f - s / (float)K;
}

// ------------------------------------------------------------------------------------
// Decode
// ------------------------------------------------------------------------------------
// numpy argmax order: a NaN beats everything, then larger value, then lower flat index.
__device__ __forceinline__ bool better(float a, int ia, float b, int ib) {
  const bool an = a != a, bn = b != b;
  if (an || bn) return an && (!bn || ia < ib);
  return a > b || (a == b && ia < ib);
}

template <typename T, int VEC>
struct DecLoad {};
template <>
struct DecLoad<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&r)[4]) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
  }
  static __device__ __forceinline__ void unpack(const uint4& v, float (&r)[4]) {
    r[0] = __uint_as_float(v.x); r[1] = __uint_as_float(v.y); r[2] = __uint_as_float(v.z); r[3] = __uint_as_float(v.w);
  }
  static __device__ __forceinline__ uint4 neg_inf() { return make_uint4(0xff800000u, 0xff800000u, 0xff800000u, 0xff800000u); }
};
template <>
struct DecLoad<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void unpack(const uint4& v, float (&r)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r[2 * i] = __uint_as_float(w[i] << 16);
      r[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&r)[8]) {
    unpack(__ldcs(reinterpret_cast<const uint4*>(p)), r);
  }
  static __device__ __forceinline__ uint4 neg_inf() { return make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u); }
};

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// Order-preserving 64-bit key of a (value, flat index) candidate: atomicMax over keys picks numpy's argmax -- a NaN beats
// everything, then the larger value (-0.0 == +0.0), then the LOWER index.  Every real key is > 0 (the identity).
__device__ __forceinline__ unsigned long long decode_key(float v, int idx) {
  uint32_t ord;
  if (v != v) {
    ord = 0xffffffffu;
  } else {
    uint32_t b = __float_as_uint(v == 0.f ? 0.f : v);
    ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  }
  return ((unsigned long long)ord << 32) | (unsigned long long)(0xffffffffu - (uint32_t)idx);
}

// 1-D bulk copy global -> shared memory (TMA engine), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kDecStages = 2;   // bulk copies in flight per CTA (default; hgb_debug_set(28, ...) varies ring geometry)

// `split` independent CTAs (16*K threads each) per sample: CTA r scans the r-th contiguous share of the map and merges its
// per-joint maxima into the sample's keys with one 64-bit atomicMax per joint; the CTA that arrives last (a counter, the
// threadfence-reduction pattern) finishes the sample: confidence, clipped 3x3 window, outputs.  The keys and the counter
// live in out_idx itself (words 0-1 of every (sample, joint) row; word 2 of joint 0), zeroed by the launcher, so no
// workspace is needed.
// The share is STREAMED THROUGH SHARED MEMORY by the TMA engine: a ring of kDecStages chunks of kDecIters*16*K vectors
// (17 KB at K = 17), refilled by one thread with cp.async.bulk as soon as a chunk has been consumed.  With register loads
// the bytes in flight are bounded by the register file (4 loads x 272 threads x 4 CTAs = 70 KB per SM, and none while a
// CTA compares or reduces): one CTA per sample reached 0.54 / 0.35 of the HBM peak (f32 / bf16) at batch 1024, four
// register-load CTAs per sample 0.57 / 0.28, a DSMEM cluster per sample less.  The ring keeps ~52 KB per CTA (3 CTAs per
// SM) in flight regardless of what the threads are doing: 0.59 / 0.32 at batch 1024, 0.84 / 0.50 at batch 4096, 0.93-1.05 /
// 0.62-0.77 at 128x128.  (A persistent variant -- 2 CTAs per SM looping over units, the ring running ahead across unit
// boundaries -- measured WORSE, 0.46 / 0.24 at batch 1024: fewer resident CTAs, same per-chunk barrier cost.)
// The vector stride VEC*16*K, the chunk and the share are multiples of K, so slot j of thread t always carries joint
// (VEC*t + j) % K: VEC running (value,index) pairs in registers, no dynamic indexing.
template <typename T, int VEC, int kDecIters>
__global__ void __launch_bounds__(320) decode_kernel(const T* __restrict__ hm, int H, int W, int K, double thr, int version,
                                                     int32_t* __restrict__ out_idx, float* __restrict__ out_kp, int stages) {
  extern __shared__ __align__(128) unsigned char s_raw[];
  const int S = 16 * K, t = threadIdx.x, b = blockIdx.y;  // blockDim = S rounded up to a warp multiple
  const int split = gridDim.x, rank = blockIdx.x;
  const int chunk_vec = kDecIters * S;
  const uint32_t chunk_bytes = (uint32_t)chunk_vec * 16u;
  const uint32_t ring = ptx::smem_u32(s_raw), bars = ring + (uint32_t)stages * chunk_bytes;
  float* s_val = reinterpret_cast<float*>(s_raw);         // [S*VEC]  (aliases the ring once the share is consumed)
  int* s_idx = reinterpret_cast<int*>(s_raw) + S * VEC;   // [S*VEC]
  __shared__ int s_last;
  const int HWK = H * W * K;
  const int nvec = HWK / VEC / split;                     // vectors of this CTA's share
  const int v0 = rank * nvec;
  const int nchunks = (nvec + chunk_vec - 1) / chunk_vec;
  const T* base = hm + (size_t)b * HWK;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(out_idx + (size_t)b * K * 4);   // stride 2 per joint
  int* counter = out_idx + (size_t)b * K * 4 + 2;

  auto issue = [&](int c) {      // one thread: chunk c -> ring slot c % kDecStages
    const int nv = min(chunk_vec, nvec - c * chunk_vec);
    const uint32_t slot = (uint32_t)(c % stages);
    ptx::mbar_expect_tx(bars + 8u * slot, (uint32_t)nv * 16u);
    bulk_load_1d(ring + slot * chunk_bytes, base + (size_t)(v0 + c * chunk_vec) * VEC, (uint32_t)nv * 16u, bars + 8u * slot);
  };
  if (t == 0) {
    for (int i = 0; i < stages; ++i) ptx::mbar_init(bars + 8u * i, 1);
    ptx::fence_barrier_init();
    for (int c = 0; c < stages && c < nchunks; ++c) issue(c);
  }
  __syncthreads();

  float bv[VEC];
  int bi[VEC];
  // Fast path.  ncu on the first ring version: 18 thread-instructions per element (compare + two selects + index add + NaN
  // test on EVERY element), issue slots 48 % busy -- instruction-bound at 0.59 of the HBM peak.  Now the common case costs
  // ~3: a thread takes G vectors of a chunk at a time (16 values per slot group), forms the per-slot maximum with fmaxf,
  // and only when that beats the slot's running maximum (rare once the maximum has settled) looks for the FIRST vector that
  // holds it -- vectors arrive in increasing index order, so "first maximum" is a strict greater-than against the running
  // value.  NaN / Inf are caught by one fused multiply-add per vector group: sum * 0 is NaN iff some value was NaN or
  // infinite (numpy: a NaN beats everything; an all -inf slot keeps its first element) and sends the CTA's share through
  // the exact comparison below (from global memory: the rare path).
  constexpr int G = 16 / VEC;             // vectors per group: 4 (f32) / 2 (bf16)
  static_assert(kDecIters % G == 0, "chunk = whole vector groups per thread");
  float nanacc = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) { bv[j] = -CUDART_INF_F; bi[j] = t < S ? (v0 + t) * VEC + j : 0x7fffffff; }
  for (int c = 0; c < nchunks; ++c) {
    const uint32_t slot = (uint32_t)(c % stages);
    ptx::mbar_wait(bars + 8u * slot, (uint32_t)(c / stages) & 1u);
    const int nv = min(chunk_vec, nvec - c * chunk_vec);
    if (t < S) {
      uint4 raw[kDecIters];
#pragma unroll
      for (int u = 0; u < kDecIters; ++u) {
        const int vi = u * S + t;
        raw[u] = vi < nv ? *reinterpret_cast<const uint4*>(s_raw + slot * chunk_bytes + (size_t)vi * 16) : DecLoad<T, VEC>::neg_inf();
      }
#pragma unroll
      for (int g0 = 0; g0 < kDecIters; g0 += G) {
        float r[G][VEC];
        float sum = 0.f;
#pragma unroll
        for (int u = 0; u < G; ++u) {
          DecLoad<T, VEC>::unpack(raw[g0 + u], r[u]);
          if (g0 * S + u * S + t < nv) {               // (padding vectors of a ragged last chunk hold -inf: not a NaN signal)
#pragma unroll
            for (int j = 0; j < VEC; ++j) sum += r[u][j];
          }
        }
        nanacc = fmaf(sum, 0.f, nanacc);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float m = r[0][j];
#pragma unroll
          for (int u = 1; u < G; ++u) m = fmaxf(m, r[u][j]);
          if (m > bv[j]) {                              // rare: find the first vector of the group that holds the maximum
            int uf = G - 1;
#pragma unroll
            for (int u = G - 2; u >= 0; --u) uf = (r[u][j] == m) ? u : uf;
            bv[j] = m;
            bi[j] = (v0 + c * chunk_vec + (g0 + uf) * S + t) * VEC + j;
          }
        }
      }
    }
    __syncthreads();                                    // the chunk is consumed: its slot can be refilled
    if (t == 0 && c + stages < nchunks) issue(c + stages);
  }
  const int nan_seen = nanacc != nanacc;
  int v;
  if (__syncthreads_or(nan_seen)) {   // block-uniform: exact numpy order (NaN first, then value, then lower index)
#pragma unroll
    for (int j = 0; j < VEC; ++j) { bv[j] = -CUDART_INF_F; bi[j] = 0x7fffffff; }
    for (v = t < S ? t : nvec; v < nvec; v += S) {
      float r[VEC];
      DecLoad<T, VEC>::load(base + (size_t)(v0 + v) * VEC, r);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int e = (v0 + v) * VEC + j;
        if (better(r[j], e, bv[j], bi[j])) { bv[j] = r[j]; bi[j] = e; }
      }
    }
  }
  if (t < S) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      s_val[t * VEC + j] = bv[j];
      s_idx[t * VEC + j] = bi[j];
    }
  }
  __syncthreads();

  const int warp = t >> 5, lane = t & 31, nwarp = blockDim.x >> 5;
  const int ncand = 16 * VEC;  // candidates per joint: c = k + K*m
  // confidence, clipped 3x3 window and outputs of joint k from its argmax element ci (whole warp)
  auto finish = [&](int k, int ci) {
    const int index = ci / K;  // flat pixel index (row-major)
    const int x = index % W;   // data_utils.py:121
    const int y = index / H;   // data_utils.py:122 (height; square maps only)
    int pidx = 0;
    float conf = 0.f;
    if (lane == 31) conf = to_f32<T>(base[ci]);   // the element itself: exact bits (signed zero, NaN payload)
    if (version == 2) {        // data_utils.py:160-169: one lane per element of the clipped 3x3 window (parallel loads)
      const int x1 = max(x - 1, 0), x2 = min(x + 2, W), y1 = max(y - 1, 0), y2 = min(y + 2, H);
      const int pw = x2 - x1, ph = y2 - y1;
      float pb = -CUDART_INF_F;
      int pbi = 0x7fffffff;
      if (lane < pw * ph) {
        const int r = lane / pw, c = lane - r * pw;
        pb = (r == 1 && c == 1) ? 0.f : to_f32<T>(base[((size_t)(y1 + r) * W + (x1 + c)) * K + k]);   // :166 reads as 0
        pbi = lane;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, pb, o);
        const int oi = __shfl_xor_sync(0xffffffffu, pbi, o);
        if (better(ov, oi, pb, pbi)) { pb = ov; pbi = oi; }
      }
      pidx = __shfl_sync(0xffffffffu, pbi, 0);
    }
    conf = __shfl_sync(0xffffffffu, conf, 31);
    __syncwarp();
    if (lane == 0) {
      const int px = pidx % 3, py = pidx / 3;  // always 3 (data_utils.py:168-169)
      int32_t* oi = out_idx + ((size_t)b * K + k) * 4;
      oi[0] = index; oi[1] = x; oi[2] = y; oi[3] = pidx;   // (split > 1: overwrites the key and, for joint 0, the counter)
      float* ok = out_kp + ((size_t)b * K + k) * 3;
      // numpy >= 2 (NEP 50) compares the float32 confidence with float32(threshold); numpy 1.x
      // promoted to float64.  They differ only when conf == float32(thr) rounds above thr.
      if (conf > (float)thr) {
        ok[0] = (float)x + 0.25f * (float)px;
        ok[1] = (float)y + 0.25f * (float)py;
        ok[2] = conf;
      } else {
        ok[0] = 0.f; ok[1] = 0.f; ok[2] = 0.f;
      }
    }
  };
  for (int k = warp; k < K; k += nwarp) {   // this CTA's maximum of joint k
    float cv = -CUDART_INF_F;
    int ci = 0x7fffffff;
    for (int m = lane; m < ncand; m += 32) {
      const float a = s_val[k + K * m];
      const int ia = s_idx[k + K * m];
      if (better(a, ia, cv, ci)) { cv = a; ci = ia; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, cv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, ci, o);
      if (better(ov, oi, cv, ci)) { cv = ov; ci = oi; }
    }
    if (split == 1) {          // the CTA saw the whole map: no merge, no keys, no counter
      finish(k, ci);
    } else if (lane == 0) {
      atomicMax(keys + 2 * k, decode_key(cv, ci));
      __threadfence();       // the key must be visible before this CTA's arrival is counted
    }
  }
  if (split == 1) return;
  __syncthreads();
  if (t == 0) s_last = atomicAdd(counter, 1) == split - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // the last CTA of the sample: every share has merged its maxima
  for (int k = warp; k < K; k += nwarp) {
    const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(keys + 2 * k);
    finish(k, (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull)));
  }
}

// ------------------------------------------------------------------------------------
// PCK / OKS
// ------------------------------------------------------------------------------------
__global__ void pck_kernel(const double* __restrict__ xp, const double* __restrict__ yp, const double* __restrict__ xg,
                           const double* __restrict__ yg, const int32_t* __restrict__ vs,
                           const double* __restrict__ bbox_wh, int N, int K, double pck_thr, int32_t* counts) {
  extern __shared__ int s_cnt[];  // [2K]
  for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N * K; i += gridDim.x * blockDim.x) {
    const int n = i / K, k = i - n * K;
    if (vs[i] > 0) {
      const double bw = bbox_wh[2 * n], bh = bbox_wh[2 * n + 1];
      const double diam = sqrt(__dadd_rn(__dmul_rn(bw, bw), __dmul_rn(bh, bh)));  // eval.py:70
      const double th = __dmul_rn(pck_thr, diam);                                  // eval.py:71
      const double dx = __dsub_rn(xg[i], xp[i]), dy = __dsub_rn(yg[i], yp[i]);
      const double dist = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));  // eval.py:85
      atomicAdd(&s_cnt[K + k], 1);
      if (dist <= th) atomicAdd(&s_cnt[k], 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * K; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(&counts[i], s_cnt[i]);
}

__constant__ double c_coco_sigmas[17] = {.026, .025, .025, .035, .035, .079, .079, .072, .072,
                                         .062, .062, .107, .107, .087, .087, .089, .089};

__global__ void oks_kernel(const double* __restrict__ xp, const double* __restrict__ yp, const double* __restrict__ xg,
                           const double* __restrict__ yg, const int32_t* __restrict__ vs, const double* __restrict__ area,
                           const double* __restrict__ bb, int N, int K, double* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int k1 = 0;
  for (int k = 0; k < K; ++k) k1 += vs[(size_t)n * K + k] > 0;
  const double a = area[n] + 2.220446049250313e-16;  // np.spacing(1)
  double s = 0;
  int cnt = 0;
  for (int k = 0; k < K; ++k) {
    const size_t i = (size_t)n * K + k;
    double dx, dy;
    if (k1 > 0) {
      if (!(vs[i] > 0)) continue;
      dx = xp[i] - xg[i];
      dy = yp[i] - yg[i];
    } else {
      const double x0 = bb[4 * n] - bb[4 * n + 2], x1 = bb[4 * n] + bb[4 * n + 2] * 2;
      const double y0 = bb[4 * n + 1] - bb[4 * n + 3], y1 = bb[4 * n + 1] + bb[4 * n + 3] * 2;
      dx = fmax(0.0, x0 - xp[i]) + fmax(0.0, xp[i] - x1);
      dy = fmax(0.0, y0 - yp[i]) + fmax(0.0, yp[i] - y1);
    }
    const double sg = c_coco_sigmas[k] * 2;
    const double e = (dx * dx + dy * dy) / (sg * sg) / a / 2;
    s += exp(-e);
    ++cnt;
  }
  out[n] = cnt ? s / cnt : 0.0;
}

}  // namespace hgb

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
using namespace hgb;

extern "C" int hgb_render_targets(const float* kps_x, const float* kps_y, const int32_t* kps_v, int B, int H, int W,
                                  int K, float* out, void* stream) {
  HGB_CHECK_ARG(kps_x && kps_y && kps_v && out, "hgb_render_targets: null pointer");
  HGB_CHECK_ARG(B >= 0 && H > 0 && W > 0 && K > 0 && K <= 1024, "hgb_render_targets: bad shape");
  HGB_CHECK_ARG(((int64_t)H * W * K) % 4 == 0, "hgb_render_targets: H*W*K must be a multiple of 4");
  if (B == 0) return HGB_OK;
  const int vec = H * W * K / 4;
  // blocks of >= 16 KB, and at least four waves of 8 resident blocks per SM where the batch allows it
  int chunks = cdiv(148 * 8 * 4, B);
  const int max_chunks = cdiv(vec, 1024);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int per = cdiv(vec, chunks);
  chunks = cdiv(vec, per);
  HGB_CHECK_ARG(B <= 65535, "hgb_render_targets: batch exceeds the grid limit");
  dim3 grid(chunks, B);
  render_targets_kernel<<<grid, 256, 2 * K * sizeof(int), (cudaStream_t)stream>>>(kps_x, kps_y, kps_v, H, W, K, vec, per, out);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

extern "C" int64_t hgb_loss_workspace_bytes(int B, int K) { return (int64_t)B * K * 4 * sizeof(float) + 256; }

template <typename TP>
static int launch_bk_sums(int kind, const float* yt, const TP* yp, int B, int HWK, int K, float* ws, cudaStream_t st) {
  const int threads = (16 * K + 31) / 32 * 32;
  const size_t smem = 3 * 4 * (16 * K) * sizeof(float);
  if (kind == HGB_LOSS_IOU)
    bk_sums_kernel<TP, 1><<<B, threads, smem, st>>>(yt, yp, HWK, K, ws);
  else
    bk_sums_kernel<TP, 0><<<B, threads, smem, st>>>(yt, yp, HWK, K, ws);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

template <typename TP, typename TG>
static int launch_loss(int kind, const float* yt, const TP* yp, TG* grad, const float* ws, int64_t nvec, int HWK, int K,
                       double inv_count, double* loss_acc, cudaStream_t st) {
  int blocks = (int)((nvec + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  switch (kind) {
    case 0: loss_grad_kernel<TP, TG, 0><<<blocks, 256, 0, st>>>(yt, yp, grad, ws, nvec, HWK, K, inv_count, loss_acc); break;
    case 1: loss_grad_kernel<TP, TG, 1><<<blocks, 256, 0, st>>>(yt, yp, grad, ws, nvec, HWK, K, inv_count, loss_acc); break;
    case 2: loss_grad_kernel<TP, TG, 2><<<blocks, 256, 0, st>>>(yt, yp, grad, ws, nvec, HWK, K, inv_count, loss_acc); break;
    default: loss_grad_kernel<TP, TG, 3><<<blocks, 256, 0, st>>>(yt, yp, grad, ws, nvec, HWK, K, inv_count, loss_acc); break;
  }
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

extern "C" int hgb_loss_fwd_bwd(int kind, const float* y_true, const void* y_pred, int pred_dtype, int B, int H, int W,
                                int K, double inv_count, double* loss_acc, void* grad, int grad_dtype, void* workspace,
                                void* stream) {
  HGB_CHECK_ARG(kind >= 0 && kind <= 3, "hgb_loss_fwd_bwd: unknown loss kind %d", kind);
  HGB_CHECK_ARG(y_true && y_pred && loss_acc, "hgb_loss_fwd_bwd: null pointer");
  HGB_CHECK_ARG(B >= 0 && H > 0 && W > 0 && K > 0 && K <= 64, "hgb_loss_fwd_bwd: bad shape");
  HGB_CHECK_ARG(((int64_t)H * W * K) % 4 == 0, "hgb_loss_fwd_bwd: H*W*K must be a multiple of 4");
  HGB_CHECK_ARG(pred_dtype == HGB_F32 || pred_dtype == HGB_BF16, "hgb_loss_fwd_bwd: bad pred dtype");
  HGB_CHECK_ARG(grad_dtype == HGB_F32 || grad_dtype == HGB_BF16, "hgb_loss_fwd_bwd: bad grad dtype");
  HGB_CHECK_ARG(kind < 2 || workspace, "hgb_loss_fwd_bwd: this loss needs a workspace");
  if (B == 0) return HGB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int HWK = H * W * K;
  const int64_t nvec = (int64_t)B * HWK / 4;
  float* ws = (float*)workspace;
  int rc = HGB_OK;
  if (kind >= 2) {
    rc = pred_dtype == HGB_F32 ? launch_bk_sums<float>(kind, y_true, (const float*)y_pred, B, HWK, K, ws, st)
                               : launch_bk_sums<__nv_bfloat16>(kind, y_true, (const __nv_bfloat16*)y_pred, B, HWK, K, ws, st);
    if (rc) return rc;
    if (kind == HGB_LOSS_IOU) {
      iou_loss_sum_kernel<<<1, 1024, 0, st>>>(ws, B * K, inv_count, loss_acc);
      HGB_LAUNCH_CHECK();
      if (!grad) return HGB_OK;
    }
  }
  if (pred_dtype == HGB_F32) {
    if (grad_dtype == HGB_F32)
      rc = launch_loss<float, float>(kind, y_true, (const float*)y_pred, (float*)grad, ws, nvec, HWK, K, inv_count, loss_acc, st);
    else
      rc = launch_loss<float, __nv_bfloat16>(kind, y_true, (const float*)y_pred, (__nv_bfloat16*)grad, ws, nvec, HWK, K, inv_count, loss_acc, st);
  } else {
    if (grad_dtype == HGB_F32)
      rc = launch_loss<__nv_bfloat16, float>(kind, y_true, (const __nv_bfloat16*)y_pred, (float*)grad, ws, nvec, HWK, K, inv_count, loss_acc, st);
    else
      rc = launch_loss<__nv_bfloat16, __nv_bfloat16>(kind, y_true, (const __nv_bfloat16*)y_pred, (__nv_bfloat16*)grad, ws, nvec, HWK, K, inv_count, loss_acc, st);
  }
  return rc;
}

extern "C" int hgb_loss_map(int kind, const float* y_true, const float* y_pred, int B, int H, int W, int K, float* out,
                            void* workspace, void* stream) {
  HGB_CHECK_ARG(kind >= 0 && kind <= 3, "hgb_loss_map: unknown loss kind %d", kind);
  HGB_CHECK_ARG(y_true && y_pred && out, "hgb_loss_map: null pointer");
  HGB_CHECK_ARG(B >= 0 && H > 0 && W > 0 && K > 0 && K <= 64, "hgb_loss_map: bad shape");
  HGB_CHECK_ARG(kind < 2 || (workspace && ((int64_t)H * W * K) % 4 == 0), "hgb_loss_map: workspace / shape");
  if (B == 0) return HGB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int HWK = H * W * K;
  float* ws = (float*)workspace;
  if (kind >= 2) {
    int rc = launch_bk_sums<float>(kind, y_true, y_pred, B, HWK, K, ws, st);
    if (rc) return rc;
  }
  const int64_t npix = (int64_t)B * H * W;
  const int blocks = (int)((npix + 255) / 256);
  switch (kind) {
    case 0: loss_map_kernel<0><<<blocks, 256, 0, st>>>(y_true, y_pred, ws, npix, H * W, K, out); break;
    case 1: loss_map_kernel<1><<<blocks, 256, 0, st>>>(y_true, y_pred, ws, npix, H * W, K, out); break;
    case 3: loss_map_kernel<3><<<blocks, 256, 0, st>>>(y_true, y_pred, ws, npix, H * W, K, out); break;
    default: iou_vec_kernel<<<(B + 127) / 128, 128, 0, st>>>(ws, B, K, out); break;
  }
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

extern "C" int hgb_decode(const void* heatmaps, int dtype, int B, int H, int W, int K, double conf_threshold,
                          int version, int32_t* out_idx, float* out_kpts, void* stream) {
  HGB_CHECK_ARG(heatmaps && out_idx && out_kpts, "hgb_decode: null pointer");
  HGB_CHECK_ARG(version == 1 || version == 2, "hgb_decode: version must be 1 or 2");
  HGB_CHECK_ARG(H == W, "hgb_decode: reference divides the flat index by height (data_utils.py:122); H must equal W");
  HGB_CHECK_ARG(B >= 0 && H >= 2 && K > 0 && K <= 64, "hgb_decode: bad shape");
  HGB_CHECK_ARG(dtype == HGB_F32 || dtype == HGB_BF16, "hgb_decode: bad dtype");
  const int vec = dtype == HGB_F32 ? 4 : 8;
  HGB_CHECK_ARG(((int64_t)H * W * K) % vec == 0, "hgb_decode: H*W*K must be a multiple of %d", vec);
  HGB_CHECK_ARG((int64_t)H * W * K < (1ll << 31), "hgb_decode: map too large");
  if (B == 0) return HGB_OK;
  HGB_CHECK_ARG(B <= 65535, "hgb_decode: batch exceeds the grid limit");
  const int threads = (16 * K + 31) / 32 * 32;
  HGB_CHECK_ARG(threads <= 320, "hgb_decode: at most 20 joints per launch configuration");
  HGB_CHECK_ARG((((uintptr_t)heatmaps | (uintptr_t)out_idx) & 15) == 0, "hgb_decode: heatmaps / out_idx must be 16-byte aligned");
  // ring of `stages` chunks of `iters` vectors per thread + their mbarriers; the (value, index) staging of the block
  // reduction aliases the ring.  hgb_debug_set(28, 10 * stages + iters) varies the geometry (iters 4 or 8).
  int stages = kDecStages, iters = 8;      // A/B on B200 (tools_decode_ab.py): 2 x 35 KB beat 4 x 17 KB and every larger ring
  if (g_debug[28] > 0) { stages = g_debug[28] / 10; iters = g_debug[28] % 10; }
  HGB_CHECK_ARG(stages >= 2 && stages <= 8 && (iters == 4 || iters == 8), "hgb_decode: ring geometry");
  const size_t ring = (size_t)stages * iters * (16 * K) * 16;
  HGB_CHECK_ARG(ring >= (size_t)(16 * K) * vec * 8, "hgb_decode: staging does not fit the ring");
  const size_t smem = ring + stages * 8;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_done = false;
  if (!attr_done) {
    HGB_CUDA(cudaFuncSetAttribute(decode_kernel<float, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    HGB_CUDA(cudaFuncSetAttribute(decode_kernel<__nv_bfloat16, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    HGB_CUDA(cudaFuncSetAttribute(decode_kernel<float, 4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    HGB_CUDA(cudaFuncSetAttribute(decode_kernel<__nv_bfloat16, 8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  HGB_CHECK_ARG(smem <= 200 * 1024, "hgb_decode: too many joints for the shared-memory ring");
  // CTAs per sample: one whole map per CTA amortises the reduction and the key merge best (A/B: split 1 >= split 2 > 4 > 8
  // at batch 1024 and 4096); small batches are split so that every SM still gets its 3 resident CTAs.  Shares must be a
  // whole number of 16-byte vectors and of pixels (K elements).  hgb_debug_set(24, n) overrides.
  const int64_t nvec = (int64_t)H * W * K / vec;
  int split = 1;
  while (split < 8 && (int64_t)B * split < 148 * 3) split <<= 1;
  if (g_debug[24] > 0) split = g_debug[24];
  while (split > 1 && (nvec % split != 0 || (nvec / split * vec) % K != 0)) split >>= 1;
  // keys + arrival counters live in out_idx (see decode_kernel); a whole map per CTA needs neither
  if (split > 1) HGB_CUDA(cudaMemsetAsync(out_idx, 0, (size_t)B * K * 4 * sizeof(int32_t), st));
  if (dtype == HGB_F32) {
    if (iters == 4) decode_kernel<float, 4, 4><<<dim3(split, B), threads, smem, st>>>((const float*)heatmaps, H, W, K, conf_threshold, version, out_idx, out_kpts, stages);
    else decode_kernel<float, 4, 8><<<dim3(split, B), threads, smem, st>>>((const float*)heatmaps, H, W, K, conf_threshold, version, out_idx, out_kpts, stages);
  } else {
    if (iters == 4) decode_kernel<__nv_bfloat16, 8, 4><<<dim3(split, B), threads, smem, st>>>((const __nv_bfloat16*)heatmaps, H, W, K, conf_threshold, version, out_idx, out_kpts, stages);
    else decode_kernel<__nv_bfloat16, 8, 8><<<dim3(split, B), threads, smem, st>>>((const __nv_bfloat16*)heatmaps, H, W, K, conf_threshold, version, out_idx, out_kpts, stages);
  }
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

extern "C" int hgb_pck_reduce(const double* xs_pred, const double* ys_pred, const double* xs_gt, const double* ys_gt,
                              const int32_t* vs, const double* bbox_wh, int N, int K, double pck_threshold,
                              int32_t* counts, void* stream) {
  HGB_CHECK_ARG(xs_pred && ys_pred && xs_gt && ys_gt && vs && bbox_wh && counts, "hgb_pck_reduce: null pointer");
  HGB_CHECK_ARG(N >= 0 && K > 0 && K <= 1024, "hgb_pck_reduce: bad shape");
  if (N == 0) return HGB_OK;
  int blocks = cdiv((int64_t)N * K, 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  pck_kernel<<<blocks, 256, 2 * K * sizeof(int), (cudaStream_t)stream>>>(xs_pred, ys_pred, xs_gt, ys_gt, vs, bbox_wh, N, K,
                                                                         pck_threshold, counts);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

extern "C" int hgb_oks_similarity(const double* xs_pred, const double* ys_pred, const double* xs_gt, const double* ys_gt,
                                  const int32_t* vs, const double* area, const double* bbox_xywh, int N, int K,
                                  double* oks_out, void* stream) {
  HGB_CHECK_ARG(xs_pred && ys_pred && xs_gt && ys_gt && vs && area && bbox_xywh && oks_out, "hgb_oks_similarity: null pointer");
  HGB_CHECK_ARG(N >= 0 && K > 0 && K <= 17, "hgb_oks_similarity: K must be in [1,17] (COCO sigmas)");
  if (N == 0) return HGB_OK;
  oks_kernel<<<cdiv(N, 128), 128, 0, (cudaStream_t)stream>>>(xs_pred, ys_pred, xs_gt, ys_gt, vs, area, bbox_xywh, N, K, oks_out);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}
