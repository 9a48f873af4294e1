// Input-path kernels: everything between a decoded frame and the network's
// (B,256,256,3) float32 input / the renderer's keypoints, on the GPU.
//   crop_resize       : tf.image.convert_image_dtype(uint8->f32) + crop_and_pad (utilities/data_utils.py:48-98) +
//                       tf.image.resize bilinear, half-pixel centres (dataset_builder.py:99,133; demo.py:44-50), one pass
//   augment_affine    : Fliplr + Affine(scale, rotate) image warp of dataset_builder.py:163-172 = cv2.warpAffine
//                       (INTER_LINEAR, constant border 0) with its 10-bit fixed-point coordinates / 5-bit fractions
//   augment_keypoints : the same flip (x -> W - x, left/right label swap) and affine on the keypoints, visibility filter
//                       (dataset_builder.py:143-185, 270-300)
//   color_augment     : brightness, contrast, saturation, hue, min-max normalisation (dataset_builder.py:190-204)
// Measured (B200, batch 256, tools_input_bench.py): with one thread per pixel recomputing its coordinates the two gather
// kernels ran at 2.6-2.8 TB/s with DRAM traffic equal to the algorithmic bytes -- bound by coordinate arithmetic and
// strided stores, not by HBM (a variant with 16-byte stores but 3x the coordinate work was 1.7x slower).  Hence the
// column-in-registers / rows-in-shared-memory tiling and the warp-level store exchange below.
// All of it streams over 786 KB per image; no arithmetic is contracted into FMAs where the reference's
// CPU kernels round each operation (explicit __fmul_rn / __fadd_rn / __dmul_rn), so results match the oracle bit for
// bit wherever the reference arithmetic is deterministic.
#include "common.cuh"

namespace hgb {

// ------------------------------------------------------------------------------------------------ crop + resize
struct Interp {
  int lo, hi;
  float lerp;
};

__device__ __forceinline__ Interp interp_of(int o, int out_size, int in_size) {
  // ResizeBilinear, half_pixel_centers: in = (o + 0.5) * (in_size / out_size) - 0.5, every step rounded to float32
  float scale = __fdiv_rn((float)in_size, (float)out_size);
  float pos = __fsub_rn(__fmul_rn(__fadd_rn((float)o, 0.5f), scale), 0.5f);
  float fl = floorf(pos);
  Interp r;
  r.lo = max((int)fl, 0);
  r.hi = min((int)ceilf(pos), in_size - 1);
  r.lerp = __fsub_rn(pos, fl);
  return r;
}

template <typename T>
__device__ __forceinline__ float load_px(const T* p);
template <>
__device__ __forceinline__ float load_px<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_px<uint8_t>(const uint8_t* p) {
  // (float)u through the mantissa of 2^23 (exact for u < 2^23) instead of I2F: ncu showed the conversion pipe 40 % busy
  const float f = __fsub_rn(__uint_as_float(0x4B000000u | (unsigned)__ldg(p)), 8388608.0f);
  return __fmul_rn(f, 0.00392156862745098f);
}

// A block covers a tile of 256 output columns x kRows rows of one image.  A thread owns one column: its horizontal
// interpolation entry (two IEEE divisions, floor, ceil) is computed ONCE and kept in registers, the kRows vertical entries
// of the tile are computed once per block into shared memory, so the per-pixel work is four gathers and three lerps
// (measured: computing the coordinates per pixel cost a third of the kernel's time).  The three floats of a pixel are
// exchanged through shared memory so that a warp stores 384 contiguous bytes with three fully coalesced instructions.
constexpr int kRows = 16;

__device__ __forceinline__ void store_pixels_coalesced(float* __restrict__ row_out, int x_first, int n_valid, float r0, float r1, float r2,
                                                       float* __restrict__ stage) {
  const int lane = threadIdx.x & 31;
  stage[3 * lane] = r0;
  stage[3 * lane + 1] = r1;
  stage[3 * lane + 2] = r2;
  __syncwarp();
  float* o = row_out + (size_t)x_first * 3;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (lane + 32 * k < 3 * n_valid) o[lane + 32 * k] = stage[lane + 32 * k];
  __syncwarp();
}

template <typename T>
__global__ void __launch_bounds__(256) crop_resize_kernel(const void* const* __restrict__ src_ptrs, const int32_t* __restrict__ src_hw,
                                                          const int32_t* __restrict__ crop_xywh, int out_h, int out_w,
                                                          float* __restrict__ out) {
  __shared__ Interp rows[kRows];
  __shared__ float stage[8][96];
  const int n = blockIdx.z;
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy0 = blockIdx.y * kRows;
  const T* src = (const T*)src_ptrs[n];
  const int sh = src_hw[2 * n], sw = src_hw[2 * n + 1];
  int x0 = 0, y0 = 0, cw = sw, ch = sh;
  if (crop_xywh) {
    x0 = crop_xywh[4 * n];
    y0 = crop_xywh[4 * n + 1];
    cw = crop_xywh[4 * n + 2];
    ch = crop_xywh[4 * n + 3];
  }
  if (threadIdx.x < kRows && oy0 + threadIdx.x < out_h) rows[threadIdx.x] = interp_of(oy0 + threadIdx.x, out_h, ch);
  __syncthreads();
  const bool live = ox < out_w;
  const Interp ix = interp_of(live ? ox : 0, out_w, cw);
  const int xs[2] = {ix.lo + x0, ix.hi + x0};
  const bool xok[2] = {live && xs[0] >= 0 && xs[0] < sw, live && xs[1] >= 0 && xs[1] < sw};
  const int xoff[2] = {xs[0] * 3, xs[1] * 3};
  const int warp_x = blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  const int n_valid = max(0, min(32, out_w - warp_x));
  if (n_valid == 0) return;                                 // whole warp beyond the image (warp-uniform)
  float* img_out = out + (size_t)n * out_h * out_w * 3;
  const int nrows = min(kRows, out_h - oy0);
  for (int r = 0; r < nrows; ++r) {
    const Interp iy = rows[r];
    const int ys[2] = {iy.lo + y0, iy.hi + y0};
    float v[2][2][3];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const bool yok = ys[a] >= 0 && ys[a] < sh;            // zero padding of crop_and_pad
      const T* row = src + (size_t)ys[a] * sw * 3;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const T* p = row + xoff[b];
#pragma unroll
        for (int c = 0; c < 3; ++c) v[a][b][c] = (yok && xok[b]) ? load_px<T>(p + c) : 0.f;
      }
    }
    float res[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float top = __fadd_rn(v[0][0][c], __fmul_rn(__fsub_rn(v[0][1][c], v[0][0][c]), ix.lerp));
      const float bot = __fadd_rn(v[1][0][c], __fmul_rn(__fsub_rn(v[1][1][c], v[1][0][c]), ix.lerp));
      res[c] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), iy.lerp));
    }
    store_pixels_coalesced(img_out + (size_t)(oy0 + r) * out_w * 3, warp_x, n_valid, res[0], res[1], res[2], stage[threadIdx.x >> 5]);
  }
}

// ------------------------------------------------------------------------------------------------ flip + affine warp
// Same tiling as crop_resize.  OpenCV itself splits the fixed-point source coordinate into a per-column part
// (adelta, bdelta) and a per-row part (X0, Y0): a thread keeps its column's pair in registers, the block computes the kRows
// row pairs once into shared memory, and the per-pixel coordinate is two 64-bit adds and shifts.
__global__ void __launch_bounds__(256) augment_affine_kernel(const float* __restrict__ in, const double* __restrict__ inv_mats,
                                                             const int32_t* __restrict__ flip, int H, int W, float* __restrict__ out) {
  __shared__ int row_x0[kRows], row_y0[kRows];              // OpenCV keeps these in 32-bit ints (saturate_cast<int>) as well
  __shared__ float stage[8][96];
  const int n = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y0 = blockIdx.y * kRows;
  const double* M = inv_mats + 6 * n;
  if (threadIdx.x < kRows) {
    const double y = (double)(y0 + threadIdx.x);
    // cv2.warpAffine: 10-bit fixed point, +16 rounds to the 1/32-pixel grid
    row_x0[threadIdx.x] = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[1], y), M[2]), 1024.0)) + 16;
    row_y0[threadIdx.x] = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(M[4], y), M[5]), 1024.0)) + 16;
  }
  __syncthreads();
  const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(M[0], (double)x), 1024.0));
  const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(M[3], (double)x), 1024.0));
  const bool fl = flip[n] != 0;
  const int warp_x = blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  const int n_valid = max(0, min(32, W - warp_x));
  if (n_valid == 0) return;
  const float* img = in + (size_t)n * H * W * 3;
  float* img_out = out + (size_t)n * H * W * 3;
  const int nrows = min(kRows, H - y0);
  for (int r = 0; r < nrows; ++r) {
    const int X = (row_x0[r] + adelta) >> 5, Y = (row_y0[r] + bdelta) >> 5;
    const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    const float fx = __fmul_rn(__fsub_rn(__uint_as_float(0x4B000000u | (unsigned)(X & 31)), 8388608.0f), 0.03125f);
    const float fy = __fmul_rn(__fsub_rn(__uint_as_float(0x4B000000u | (unsigned)(Y & 31)), 8388608.0f), 0.03125f);
    const float wx[2] = {__fsub_rn(1.f, fx), fx}, wy[2] = {__fsub_rn(1.f, fy), fy};
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int yy = sy + a, xx = sx + b;
        const bool ok = x < W && yy >= 0 && yy < H && xx >= 0 && xx < W;
        const float w = __fmul_rn(wy[a], wx[b]);
        const float* p = img + (yy * W + (fl ? W - 1 - xx : xx)) * 3;     // 32-bit index: H * W * 3 < 2^31 (checked by the host)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float t = __fmul_rn(ok ? __ldg(p + c) : 0.f, w);
          acc[c] = (a == 0 && b == 0) ? t : __fadd_rn(acc[c], t);
        }
      }
    store_pixels_coalesced(img_out + (size_t)(y0 + r) * W * 3, warp_x, n_valid, acc[0], acc[1], acc[2], stage[threadIdx.x >> 5]);
  }
}

__global__ void augment_keypoints_kernel(const float* __restrict__ kx, const float* __restrict__ ky, const int32_t* __restrict__ kv,
                                         const int32_t* __restrict__ flip, const double* __restrict__ fwd_mats,
                                         const int32_t* __restrict__ partner, int N, int K, float label_w,
                                         float* __restrict__ ox, float* __restrict__ oy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int n = i / K, k = i - n * K;
  const bool fl = flip[n] != 0;
  const int src = fl ? partner[k] : k;                 // after the label swap, slot k holds its partner's joint
  const int v = kv[n * K + src];
  float x = v > 0 ? kx[n * K + src] : 0.f, y = v > 0 ? ky[n * K + src] : 0.f;
  if (fl) x = __fsub_rn(label_w, x);
  const double* M = fwd_mats + 6 * n;
  const double xa = __dadd_rn(__dadd_rn(__dmul_rn(M[0], (double)x), __dmul_rn(M[1], (double)y)), M[2]);
  const double ya = __dadd_rn(__dadd_rn(__dmul_rn(M[3], (double)x), __dmul_rn(M[4], (double)y)), M[5]);
  ox[i] = v > 0 ? (float)xa : 0.f;
  oy[i] = v > 0 ? (float)ya : 0.f;
}

// ------------------------------------------------------------------------------------------------ colour
// workspace per image: double sum[3] | uint32 min_key, max_key (order-preserving float keys)
struct ColorWs {
  double sum[3];
  unsigned int min_key, max_key;
};

__device__ __forceinline__ unsigned int float_key(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void color_init_kernel(ColorWs* ws, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  ws[n].sum[0] = ws[n].sum[1] = ws[n].sum[2] = 0.0;
  ws[n].min_key = 0xffffffffu;
  ws[n].max_key = 0u;
}

__global__ void __launch_bounds__(256) color_sum_kernel(const float* __restrict__ img, const float* __restrict__ params, int HW,
                                                        ColorWs* __restrict__ ws) {
  const int n = blockIdx.y;
  const float delta = params[4 * n];
  const float* p = img + (size_t)n * HW * 3;
  const int total = HW * 3;
  double s[3] = {0.0, 0.0, 0.0};
  const int stride = gridDim.x * blockDim.x;
  if ((total & 3) == 0) {                                   // 16-byte loads over the flat image; element j belongs to channel j % 3
    const float4* p4 = reinterpret_cast<const float4*>(p);
#pragma unroll 4
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total / 4; q += stride) {
      const float4 v = __ldg(p4 + q);
      const int r = q % 3;                                   // (4q) % 3
      const double a = (double)__fadd_rn(v.x, delta) + (double)__fadd_rn(v.w, delta), b = (double)__fadd_rn(v.y, delta),
                   c = (double)__fadd_rn(v.z, delta);
      // channels of (x, y, z, w): r, r+1, r+2, r  (mod 3)
      s[0] += r == 0 ? a : (r == 1 ? c : b);
      s[1] += r == 0 ? b : (r == 1 ? a : c);
      s[2] += r == 0 ? c : (r == 1 ? b : a);
    }
  } else {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += stride) {
      const double v = (double)__fadd_rn(__ldg(p + j), delta);
      const int c = j % 3;
      s[0] += c == 0 ? v : 0.0;
      s[1] += c == 1 ? v : 0.0;
      s[2] += c == 2 ? v : 0.0;
    }
  }
  __shared__ double sh[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
    if ((threadIdx.x & 31) == 0) sh[c][threadIdx.x >> 5] = s[c];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
    atomicAdd(&ws[n].sum[threadIdx.x], t);
  }
}

__device__ __forceinline__ void rgb_to_hsv(float r, float g, float b, float& h, float& s, float& v) {
  const float vv = fmaxf(r, fmaxf(g, b));
  const float range = __fsub_rn(vv, fminf(r, fminf(g, b)));
  s = vv > 0.f ? __fdiv_rn(range, vv) : 0.f;
  const float norm = __fdiv_rn(1.0f, __fmul_rn(6.0f, range));
  float hh;
  if (r == vv) hh = __fmul_rn(norm, __fsub_rn(g, b));
  else if (g == vv) hh = __fadd_rn(__fmul_rn(norm, __fsub_rn(b, r)), (float)(2.0 / 6.0));
  else hh = __fadd_rn(__fmul_rn(norm, __fsub_rn(r, g)), (float)(4.0 / 6.0));
  if (range <= 0.f) hh = 0.f;
  if (hh < 0.f) hh = __fadd_rn(hh, 1.f);
  h = hh;
  v = vv;
}

__device__ __forceinline__ void hsv_to_rgb(float h, float s, float v, float& r, float& g, float& b) {
  const float c = __fmul_rn(s, v), m = __fsub_rn(v, c), dh = __fmul_rn(h, 6.f);
  const int cat = (int)dh;
  float fm = dh;
  for (int i = 0; i < 4; ++i) {
    if (fm <= 0.f) fm = __fadd_rn(fm, 2.f);
    if (fm >= 2.f) fm = __fsub_rn(fm, 2.f);
  }
  const float x = __fmul_rn(c, __fsub_rn(1.f, fabsf(__fsub_rn(fm, 1.f))));
  float rr = 0.f, gg = 0.f, bb = 0.f;
  switch (cat) {
    case 0: rr = c; gg = x; break;
    case 1: rr = x; gg = c; break;
    case 2: gg = c; bb = x; break;
    case 3: gg = x; bb = c; break;
    case 4: rr = x; bb = c; break;
    case 5: rr = c; bb = x; break;
    default: break;
  }
  r = __fadd_rn(rr, m);
  g = __fadd_rn(gg, m);
  b = __fadd_rn(bb, m);
}

__device__ __forceinline__ void adjust_hue(float& r, float& g, float& b, float delta) {
  float vmax, vmid, vmin;
  int cat;
  if (r < g) {
    if (b < r) { vmax = g; vmid = r; vmin = b; cat = 1; }
    else if (b > g) { vmax = b; vmid = g; vmin = r; cat = 3; }
    else { vmax = g; vmid = b; vmin = r; cat = 2; }
  } else {
    if (b < g) { vmax = r; vmid = g; vmin = b; cat = 0; }
    else if (b > r) { vmax = b; vmid = r; vmin = g; cat = 4; }
    else { vmax = r; vmid = b; vmin = g; cat = 5; }
  }
  float h = 0.f;
  if (vmax != vmin) {
    const float ratio = __fdiv_rn(__fsub_rn(vmid, vmin), __fsub_rn(vmax, vmin));
    h = __fadd_rn((float)cat, (cat & 1) == 0 ? ratio : __fsub_rn(1.f, ratio));
  }
  h = __fadd_rn(h, __fmul_rn(delta, 6.f));
  for (int i = 0; i < 3; ++i) {
    if (h < 0.f) h = __fadd_rn(h, 6.f);
    if (h >= 6.f) h = __fsub_rn(h, 6.f);
  }
  const int c2 = (int)h;
  float ratio2 = __fsub_rn(h, (float)c2);
  if (c2 & 1) ratio2 = __fsub_rn(1.f, ratio2);
  const float mid = __fadd_rn(vmin, __fmul_rn(ratio2, __fsub_rn(vmax, vmin)));
  switch (c2) {
    case 0: r = vmax; g = mid; b = vmin; break;
    case 1: r = mid; g = vmax; b = vmin; break;
    case 2: r = vmin; g = vmax; b = mid; break;
    case 3: r = vmin; g = mid; b = vmax; break;
    case 4: r = mid; g = vmin; b = vmax; break;
    default: r = vmax; g = vmin; b = mid; break;
  }
}

// Pixels are processed whole (HSV needs the three channels), but global memory is touched only with coalesced 16-byte
// accesses: a block stages tiles of 1024 pixels (12 KB) through shared memory; a thread owns pixels t, t+256, ... of the
// tile (stride-3 word accesses: conflict-free).
constexpr int kColorTile = 1024;

__global__ void __launch_bounds__(256) color_apply_kernel(float* __restrict__ img, const float* __restrict__ params, int HW,
                                                          ColorWs* __restrict__ ws) {
  __shared__ __align__(16) float tile[kColorTile * 3];
  const int n = blockIdx.y;
  const float delta = params[4 * n], contrast = params[4 * n + 1], sat = params[4 * n + 2], hue = params[4 * n + 3];
  float mean[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) mean[c] = (float)(ws[n].sum[c] / (double)HW);
  float* p = img + (size_t)n * HW * 3;
  const bool vec = ((HW * 3) & 3) == 0;
  float lo = INFINITY, hi = -INFINITY;
  for (int base = blockIdx.x * kColorTile; base < HW; base += gridDim.x * kColorTile) {
    const int npx = min(kColorTile, HW - base), nfl = npx * 3;
    float* g = p + (size_t)base * 3;                        // base * 3 floats: 16-byte aligned because kColorTile * 3 % 4 == 0
    if (vec && (nfl & 3) == 0) {
      for (int q = threadIdx.x; q < nfl / 4; q += blockDim.x) reinterpret_cast<float4*>(tile)[q] = reinterpret_cast<const float4*>(g)[q];
    } else {
      for (int q = threadIdx.x; q < nfl; q += blockDim.x) tile[q] = g[q];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < npx; t += blockDim.x) {
      float q[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float x = __fadd_rn(tile[t * 3 + c], delta);
        q[c] = __fadd_rn(__fmul_rn(__fsub_rn(x, mean[c]), contrast), mean[c]);
      }
      float h, s, v;
      rgb_to_hsv(q[0], q[1], q[2], h, s, v);
      s = fminf(1.f, fmaxf(0.f, __fmul_rn(s, sat)));
      hsv_to_rgb(h, s, v, q[0], q[1], q[2]);
      adjust_hue(q[0], q[1], q[2], hue);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        tile[t * 3 + c] = q[c];
        lo = fminf(lo, q[c]);
        hi = fmaxf(hi, q[c]);
      }
    }
    __syncthreads();
    if (vec && (nfl & 3) == 0) {
      for (int q = threadIdx.x; q < nfl / 4; q += blockDim.x) reinterpret_cast<float4*>(g)[q] = reinterpret_cast<const float4*>(tile)[q];
    } else {
      for (int q = threadIdx.x; q < nfl; q += blockDim.x) g[q] = tile[q];
    }
    __syncthreads();
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomicMin(&ws[n].min_key, float_key(lo));
    atomicMax(&ws[n].max_key, float_key(hi));
  }
}

__global__ void __launch_bounds__(256) color_normalize_kernel(float* __restrict__ img, int HW, const ColorWs* __restrict__ ws) {
  const int n = blockIdx.y;
  const float lo = key_float(ws[n].min_key), hi = key_float(ws[n].max_key);
  const float span = __fsub_rn(hi, lo);
  float* p = img + (size_t)n * HW * 3;
  const int total = HW * 3, stride = gridDim.x * blockDim.x;
  if ((total & 3) == 0) {
    float4* p4 = reinterpret_cast<float4*>(p);
#pragma unroll 4
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total / 4; q += stride) {
      float4 v = p4[q];
      v.x = __fdiv_rn(__fsub_rn(v.x, lo), span);
      v.y = __fdiv_rn(__fsub_rn(v.y, lo), span);
      v.z = __fdiv_rn(__fsub_rn(v.z, lo), span);
      v.w = __fdiv_rn(__fsub_rn(v.w, lo), span);
      p4[q] = v;
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) p[i] = __fdiv_rn(__fsub_rn(p[i], lo), span);
  }
}

}  // namespace hgb

using namespace hgb;

extern "C" int hgb_crop_resize(const void* const* src_ptrs, const int32_t* src_hw, int src_dtype, const int32_t* crop_xywh, int N,
                               int out_h, int out_w, float* out, void* stream) {
  HGB_CHECK_ARG(src_ptrs && src_hw && out, "hgb_crop_resize: null pointer");
  HGB_CHECK_ARG(src_dtype == HGB_F32 || src_dtype == HGB_U8, "hgb_crop_resize: source dtype must be HGB_F32 or HGB_U8");
  HGB_CHECK_ARG(N >= 0 && N <= 65535 && out_h > 0 && out_w > 0, "hgb_crop_resize: bad sizes (N <= 65535)");
  if (N == 0) return HGB_OK;
  HGB_CHECK_ARG(cdiv(out_h, kRows) <= 65535, "hgb_crop_resize: out_h too large");
  dim3 grid(cdiv(out_w, 256), cdiv(out_h, kRows), N);
  if (src_dtype == HGB_U8)
    crop_resize_kernel<uint8_t><<<grid, 256, 0, (cudaStream_t)stream>>>(src_ptrs, src_hw, crop_xywh, out_h, out_w, out);
  else
    crop_resize_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(src_ptrs, src_hw, crop_xywh, out_h, out_w, out);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

extern "C" int hgb_augment_affine(const float* images, const double* inv_mats, const int32_t* flip, int N, int H, int W, float* out,
                                  void* stream) {
  HGB_CHECK_ARG(images && inv_mats && flip && out, "hgb_augment_affine: null pointer");
  HGB_CHECK_ARG(images != out, "hgb_augment_affine: the warp is a gather and cannot run in place");
  HGB_CHECK_ARG(N >= 0 && N <= 65535 && H > 0 && W > 0 && H <= 32767 && W <= 32767 && (int64_t)H * W * 3 < (1LL << 31),
                "hgb_augment_affine: bad sizes (H * W * 3 must stay below 2^31)");
  if (N == 0) return HGB_OK;
  augment_affine_kernel<<<dim3(cdiv(W, 256), cdiv(H, kRows), N), 256, 0, (cudaStream_t)stream>>>(images, inv_mats, flip, H, W, out);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

extern "C" int hgb_augment_keypoints(const float* kps_x, const float* kps_y, const int32_t* kps_v, const int32_t* flip,
                                     const double* fwd_mats, const int32_t* flip_partner, int N, int K, int label_w, float* out_x,
                                     float* out_y, void* stream) {
  HGB_CHECK_ARG(kps_x && kps_y && kps_v && flip && fwd_mats && flip_partner && out_x && out_y, "hgb_augment_keypoints: null pointer");
  HGB_CHECK_ARG(N >= 0 && K > 0 && label_w > 0, "hgb_augment_keypoints: bad sizes");
  if (N == 0) return HGB_OK;
  augment_keypoints_kernel<<<cdiv((int64_t)N * K, 128), 128, 0, (cudaStream_t)stream>>>(kps_x, kps_y, kps_v, flip, fwd_mats, flip_partner, N,
                                                                                      K, (float)label_w, out_x, out_y);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}

extern "C" int64_t hgb_color_workspace_bytes(int N) { return (int64_t)sizeof(ColorWs) * (N > 0 ? N : 0); }

extern "C" int hgb_color_augment(float* images, const float* params, int N, int H, int W, void* workspace, void* stream) {
  HGB_CHECK_ARG(images && params && workspace, "hgb_color_augment: null pointer");
  HGB_CHECK_ARG(N >= 0 && N <= 65535 && H > 0 && W > 0, "hgb_color_augment: bad sizes");
  HGB_CHECK_ARG(((uintptr_t)images & 15) == 0 && ((uintptr_t)workspace & 7) == 0, "hgb_color_augment: images must be 16-byte aligned");
  if (N == 0) return HGB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ColorWs* ws = (ColorWs*)workspace;
  const int HW = H * W;
  // enough blocks per image to fill 148 SMs at small N, grid-stride beyond that
  // ~8 resident blocks per SM in total; every thread keeps several 16-byte accesses in flight (grid-stride, unrolled)
  const int per_img = max(1, min(cdiv(HW * 3, 4 * 256), cdiv(148 * 8, N)));
  const int tiles = max(1, min(cdiv(HW, kColorTile), cdiv(148 * 8, N)));
  color_init_kernel<<<cdiv(N, 128), 128, 0, st>>>(ws, N);
  color_sum_kernel<<<dim3(per_img, N), 256, 0, st>>>(images, params, HW, ws);
  color_apply_kernel<<<dim3(tiles, N), 256, 0, st>>>(images, params, HW, ws);
  color_normalize_kernel<<<dim3(per_img, N), 256, 0, st>>>(images, HW, ws);
  HGB_LAUNCH_CHECK();
  return HGB_OK;
}
