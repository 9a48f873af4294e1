// Host-side IO helpers of the input path: TFRecord framing checksum and JPEG decode.
//   hgb_crc32c        : CRC-32C (Castagnoli) of the TFRecord framing tf.data.TFRecordDataset verifies
//                       (dataset_builder.py:39,48,63), slicing-by-8 on the host
//   hgb_tfrecord_scan : walks the record framing of a whole (memory-mapped) TFRecord file, verifying both CRCs per record
//   hgb_example_parse : one serialized tf.train.Example -> feature table + decoded float / int64 values
//                       (tf.io.parse_single_example, dataset_builder.py:262), a protobuf wire-format walk on the host
//   hgb_jpeg_info     : size / component count from the SOF marker (what tf.image.decode_image reads first, :263)
//   hgb_jpeg_decode   : baseline / progressive JPEG -> interleaved RGB uint8 in DEVICE memory through nvJPEG
//                       (library decode = plumbing, like cuBLAS for a plain GEMM); libnvjpeg is opened at first use so
//                       libhgb200.so itself has no load-time dependency on it.  nvJPEG's default backend runs the
//                       Huffman stage on the host, so a batch is spread over host threads, each with its own decoder
//                       state and CUDA stream; the caller's stream is ordered before and after them with events.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace hgb {

// ------------------------------------------------------------------------------------------------ CRC-32C
static uint32_t g_crc_table[8][256];
static std::once_flag g_crc_once;

static void crc_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xff];
}

// ------------------------------------------------------------------------------------------------ nvJPEG binding
struct NvJpeg {
  void* dso = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  nvjpegStatus_t (*create)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*state_create)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                           cudaStream_t) = nullptr;
  std::string error;
  bool ok = false;
  struct Worker {
    nvjpegJpegState_t state = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
  };
  std::vector<Worker> workers;
  cudaEvent_t begin = nullptr;
  int device = -1;
};
static NvJpeg g_nvjpeg;
static std::mutex g_nvjpeg_mutex;

static bool nvjpeg_ready() {
  NvJpeg& j = g_nvjpeg;
  if (j.ok) return true;
  if (!j.error.empty()) return false;
  const char* names[] = {"libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so", "libnvjpeg.so"};
  for (const char* n : names) {
    j.dso = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (j.dso) break;
  }
  if (!j.dso) {
    j.error = "libnvjpeg.so.12 not found (CUDA toolkit library); JPEG decode is unavailable";
    return false;
  }
  j.create = (decltype(j.create))dlsym(j.dso, "nvjpegCreateSimple");
  j.state_create = (decltype(j.state_create))dlsym(j.dso, "nvjpegJpegStateCreate");
  j.decode = (decltype(j.decode))dlsym(j.dso, "nvjpegDecode");
  if (!j.create || !j.state_create || !j.decode) {
    j.error = "libnvjpeg lacks nvjpegCreateSimple / nvjpegJpegStateCreate / nvjpegDecode";
    return false;
  }
  nvjpegStatus_t s = j.create(&j.handle);
  if (s == NVJPEG_STATUS_SUCCESS) s = j.state_create(j.handle, &j.state);
  if (s != NVJPEG_STATUS_SUCCESS) {
    j.error = "nvjpeg initialisation failed with status " + std::to_string((int)s);
    return false;
  }
  j.ok = true;
  return true;
}

}  // namespace hgb

using namespace hgb;

#if defined(__x86_64__) && defined(__GNUC__)
// The CRC32 instruction of SSE4.2 computes exactly this polynomial (what TensorFlow's own crc32c uses when available);
// taken when the CPU reports it, the table walk below is the portable path and the definition both are tested against.
__attribute__((target("sse4.2"))) static uint32_t crc32c_hw(const uint8_t* p, int64_t len) {
  uint64_t c = 0xffffffffu;
  while (len > 0 && ((uintptr_t)p & 7)) {
    c = __builtin_ia32_crc32qi((uint32_t)c, *p++);
    --len;
  }
  while (len >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    c = __builtin_ia32_crc32di(c, w);
    p += 8;
    len -= 8;
  }
  while (len-- > 0) c = __builtin_ia32_crc32qi((uint32_t)c, *p++);
  return (uint32_t)c ^ 0xffffffffu;
}
static bool crc32c_has_hw() {
  static const bool yes = __builtin_cpu_supports("sse4.2");
  return yes;
}
#else
static uint32_t crc32c_hw(const uint8_t*, int64_t) { return 0; }
static bool crc32c_has_hw() { return false; }
#endif

extern "C" uint32_t hgb_crc32c_portable(const void* data, int64_t len);

extern "C" uint32_t hgb_crc32c(const void* data, int64_t len) {
  if (crc32c_has_hw()) return crc32c_hw((const uint8_t*)data, len);
  return hgb_crc32c_portable(data, len);
}

extern "C" uint32_t hgb_crc32c_portable(const void* data, int64_t len) {
  std::call_once(g_crc_once, crc_init);
  const uint8_t* p = (const uint8_t*)data;
  uint32_t c = 0xffffffffu;
  while (len > 0 && ((uintptr_t)p & 7)) {
    c = g_crc_table[0][(c ^ *p++) & 0xff] ^ (c >> 8);
    --len;
  }
  while (len >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = g_crc_table[7][w & 0xff] ^ g_crc_table[6][(w >> 8) & 0xff] ^ g_crc_table[5][(w >> 16) & 0xff] ^ g_crc_table[4][(w >> 24) & 0xff] ^
        g_crc_table[3][(w >> 32) & 0xff] ^ g_crc_table[2][(w >> 40) & 0xff] ^ g_crc_table[1][(w >> 48) & 0xff] ^ g_crc_table[0][w >> 56];
    p += 8;
    len -= 8;
  }
  while (len-- > 0) c = g_crc_table[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return c ^ 0xffffffffu;
}

extern "C" int hgb_jpeg_info(const uint8_t* data, int64_t len, int32_t* height, int32_t* width, int32_t* components) {
  HGB_CHECK_ARG(data && height && width && components, "hgb_jpeg_info: null pointer");
  HGB_CHECK_ARG(len >= 4 && data[0] == 0xFF && data[1] == 0xD8, "hgb_jpeg_info: not a JPEG stream (no SOI marker)");
  int64_t i = 2;
  while (i + 4 <= len) {
    if (data[i] != 0xFF) { ++i; continue; }
    const uint8_t m = data[i + 1];
    if (m == 0xFF) { ++i; continue; }                                   // fill byte
    if (m == 0xD8 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { i += 2; continue; }   // markers without a length
    const int64_t seg = ((int64_t)data[i + 2] << 8) | data[i + 3];
    const bool sof = m >= 0xC0 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC;
    if (sof) {
      HGB_CHECK_ARG(i + 10 <= len, "hgb_jpeg_info: truncated frame header");
      *height = (data[i + 5] << 8) | data[i + 6];
      *width = (data[i + 7] << 8) | data[i + 8];
      *components = data[i + 9];
      HGB_CHECK_ARG(*height > 0 && *width > 0, "hgb_jpeg_info: empty frame");
      return HGB_OK;
    }
    if (m == 0xDA || m == 0xD9) break;                                   // start of scan / end of image before any frame
    i += 2 + seg;
  }
  set_error("hgb_jpeg_info: no frame header found");
  return HGB_ERR_INVALID;
}

extern "C" int hgb_jpeg_decode(const uint8_t* const* datas, const int64_t* lens, int N, uint8_t* const* outs, const int32_t* hw,
                               void* stream) {
  HGB_CHECK_ARG(N >= 0, "hgb_jpeg_decode: negative count");
  if (N == 0) return HGB_OK;
  HGB_CHECK_ARG(datas && lens && outs && hw, "hgb_jpeg_decode: null pointer");
  std::lock_guard<std::mutex> lock(g_nvjpeg_mutex);
  if (!nvjpeg_ready()) {
    set_error("%s", g_nvjpeg.error.c_str());
    return HGB_ERR_STATE;
  }
  for (int n = 0; n < N; ++n) {
    HGB_CHECK_ARG(datas[n] && outs[n] && lens[n] > 0, "hgb_jpeg_decode: image %d: null pointer / empty stream", n);
    int32_t h, w, c;
    if (int rc = hgb_jpeg_info(datas[n], lens[n], &h, &w, &c)) return rc;
    HGB_CHECK_ARG(h == hw[2 * n] && w == hw[2 * n + 1], "hgb_jpeg_decode: image %d is %dx%d, the output buffer was sized for %dx%d", n, h, w,
                  hw[2 * n], hw[2 * n + 1]);
    HGB_CHECK_ARG(c == 1 || c == 3, "hgb_jpeg_decode: image %d has %d components (grey and YCbCr/RGB only)", n, c);
  }
  NvJpeg& j = g_nvjpeg;
  int device = 0;
  HGB_CUDA(cudaGetDevice(&device));
  if (j.device != device) {                      // worker streams / events belong to one device
    HGB_CHECK_ARG(j.device < 0, "hgb_jpeg_decode: the decoder was initialised on device %d, called on device %d", j.device, device);
    j.device = device;
    HGB_CUDA(cudaEventCreateWithFlags(&j.begin, cudaEventDisableTiming));
  }
  int want = (int)std::thread::hardware_concurrency();
  if (const char* e = getenv("HGB_JPEG_THREADS")) want = atoi(e);
  const int T = std::max(1, std::min(std::min(want, 32), N));
  while ((int)j.workers.size() < T) {
    NvJpeg::Worker w;
    nvjpegStatus_t s = j.state_create(j.handle, &w.state);
    if (s != NVJPEG_STATUS_SUCCESS) {
      set_error("hgb_jpeg_decode: nvjpegJpegStateCreate failed with status %d", (int)s);
      return HGB_ERR_CUDA;
    }
    // Highest stream priority: the decoder's kernels are tiny and nvjpegDecode waits for them per image.  At the default
    // (lowest) priority they starve behind a training step that keeps every SM busy with persistent CTAs from
    // higher-priority lanes -- the background prefetcher of round 1 measured 49.8 img/s for exactly this reason.
    int prio_lo = 0, prio_hi = 0;
    HGB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    HGB_CUDA(cudaStreamCreateWithPriority(&w.stream, cudaStreamNonBlocking, prio_hi));
    HGB_CUDA(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
    j.workers.push_back(w);
  }
  cudaStream_t caller = (cudaStream_t)stream;
  HGB_CUDA(cudaEventRecord(j.begin, caller));      // the output buffers may still be in use by earlier work on the caller's stream
  std::atomic<int> failed_image{-1}, failed_status{0};
  auto work = [&](int t) {
    cudaSetDevice(device);
    NvJpeg::Worker& w = j.workers[t];
    cudaStreamWaitEvent(w.stream, j.begin, 0);
    for (int n = t; n < N && failed_image.load() < 0; n += T) {
      nvjpegImage_t img = {};
      img.channel[0] = outs[n];
      img.pitch[0] = (size_t)hw[2 * n + 1] * 3;
      nvjpegStatus_t s = j.decode(j.handle, w.state, datas[n], (size_t)lens[n], NVJPEG_OUTPUT_RGBI, &img, w.stream);
      if (s != NVJPEG_STATUS_SUCCESS) {
        failed_status.store((int)s);
        failed_image.store(n);
      }
    }
    cudaEventRecord(w.done, w.stream);
  };
  if (T == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t) pool.emplace_back(work, t);
    for (auto& th : pool) th.join();
  }
  for (int t = 0; t < T; ++t) HGB_CUDA(cudaStreamWaitEvent(caller, j.workers[t].done, 0));
  if (failed_image.load() >= 0) {
    set_error("hgb_jpeg_decode: nvjpegDecode failed on image %d with status %d", failed_image.load(), failed_status.load());
    return HGB_ERR_CUDA;
  }
  return HGB_OK;
}

// ------------------------------------------------------------------------------------------------ tf.train.Example
namespace {
struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (p >= end) { ok = false; return 0; }
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  // next field: returns false at the end; for wire type 2 `sub` is the payload, for 0 `value`, for 1 / 5 `sub` spans the fixed bytes
  bool next(uint32_t& number, uint32_t& wire, uint64_t& value, Cursor& sub) {
    if (p >= end || !ok) return false;
    const uint64_t key = varint();
    number = (uint32_t)(key >> 3);
    wire = (uint32_t)(key & 7);
    if (!ok) return false;
    if (wire == 0) {
      value = varint();
    } else if (wire == 2 || wire == 1 || wire == 5) {
      const uint64_t n = wire == 2 ? varint() : (wire == 1 ? 8 : 4);
      if (!ok || n > (uint64_t)(end - p)) { ok = false; return false; }
      sub.p = p;
      sub.end = p + n;
      sub.ok = true;
      p += n;
    } else {
      ok = false;
    }
    return ok;
  }
};
}  // namespace

extern "C" int hgb_example_parse(const uint8_t* data, int64_t len, int max_features, int64_t* table, float* fvals, int64_t fcap,
                                 int64_t* ivals, int64_t icap) {
  HGB_CHECK_ARG(data && table && fvals && ivals && len >= 0 && max_features > 0, "hgb_example_parse: bad arguments");
  Cursor ex{data, data + len};
  int nfeat = 0;
  int64_t nf = 0, ni = 0;
  uint32_t num, wire;
  uint64_t val;
  Cursor features{nullptr, nullptr}, entry{nullptr, nullptr}, field{nullptr, nullptr}, list{nullptr, nullptr}, item{nullptr, nullptr};
  while (ex.next(num, wire, val, features)) {
    if (num != 1 || wire != 2) continue;                                  // Example.features
    while (features.next(num, wire, val, entry)) {
      if (num != 1 || wire != 2) continue;                                // Features.feature map entry
      int64_t name_off = -1, name_len = 0, kind = 0, start = 0, count = 0, extra = 0;
      while (entry.next(num, wire, val, field)) {
        if (num == 1 && wire == 2) {
          name_off = field.p - data;
          name_len = field.end - field.p;
        } else if (num == 2 && wire == 2) {                               // Feature: oneof bytes_list / float_list / int64_list
          while (field.next(num, wire, val, list)) {
            if (wire != 2 || num < 1 || num > 3) continue;
            kind = num;
            start = num == 2 ? nf : ni;
            count = 0;
            while (list.next(num, wire, val, item)) {
              if (num != 1) continue;
              if (kind == 1) {                                            // bytes: offset / length of the first value, count of values
                if (wire != 2) { list.ok = false; break; }
                if (count == 0) { start = item.p - data; extra = item.end - item.p; }
                ++count;
              } else if (kind == 2) {                                     // floats: packed chunk(s) or single fixed32 values
                if (wire != 2 && wire != 5) { list.ok = false; break; }
                const int64_t n = (item.end - item.p) / 4;
                if ((item.end - item.p) % 4 || nf + n > fcap) { set_error("hgb_example_parse: float capacity exceeded / bad packing"); return HGB_ERR_STATE; }
                memcpy(fvals + nf, item.p, (size_t)n * 4);
                nf += n;
                count += n;
              } else {                                                    // int64: packed varints or single varints
                if (wire == 0) {
                  if (ni + 1 > icap) { set_error("hgb_example_parse: int64 capacity exceeded"); return HGB_ERR_STATE; }
                  ivals[ni++] = (int64_t)val;
                  ++count;
                } else if (wire == 2) {
                  while (item.p < item.end && item.ok) {
                    const uint64_t v = item.varint();
                    if (!item.ok) break;
                    if (ni + 1 > icap) { set_error("hgb_example_parse: int64 capacity exceeded"); return HGB_ERR_STATE; }
                    ivals[ni++] = (int64_t)v;
                    ++count;
                  }
                  if (!item.ok) list.ok = false;
                } else {
                  list.ok = false;
                }
              }
            }
            if (!list.ok) field.ok = false;
          }
          if (!field.ok) entry.ok = false;
        }
      }
      if (!entry.ok) features.ok = false;
      if (name_off >= 0 && features.ok) {
        if (nfeat >= max_features) { set_error("hgb_example_parse: more than %d features", max_features); return HGB_ERR_STATE; }
        int64_t* row = table + 6 * nfeat++;
        row[0] = name_off; row[1] = name_len; row[2] = kind; row[3] = start; row[4] = count; row[5] = extra;
      }
    }
    if (!features.ok) ex.ok = false;
  }
  HGB_CHECK_ARG(ex.ok, "hgb_example_parse: malformed tf.train.Example");
  return nfeat;
}

// ------------------------------------------------------------------------------------------------ TFRecord framing
static inline uint32_t masked_crc(const uint8_t* p, int64_t n) {
  const uint32_t c = hgb_crc32c(p, n);
  return ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

extern "C" int64_t hgb_tfrecord_scan(const uint8_t* data, int64_t len, int64_t start, int verify, int64_t* offsets, int64_t* lengths,
                                     int64_t cap, int64_t* next) {
  if (!data || !offsets || !lengths || !next || len < 0 || start < 0 || start > len || cap <= 0) {
    set_error("hgb_tfrecord_scan: bad arguments");
    return HGB_ERR_INVALID;
  }
  int64_t at = start, n = 0;
  while (at < len && n < cap) {
    if (len - at < 12) { set_error("truncated record header at byte %lld", (long long)at); return HGB_ERR_INVALID; }
    uint64_t size;
    uint32_t crc;
    memcpy(&size, data + at, 8);
    memcpy(&crc, data + at + 8, 4);
    if (verify && masked_crc(data + at, 8) != crc) { set_error("corrupted record length at byte %lld", (long long)at); return HGB_ERR_INVALID; }
    if (size > (uint64_t)(len - at - 12) || (uint64_t)(len - at - 12) - size < 4) {
      set_error("truncated record at byte %lld", (long long)at);
      return HGB_ERR_INVALID;
    }
    const uint8_t* payload = data + at + 12;
    memcpy(&crc, payload + size, 4);
    if (verify && masked_crc(payload, (int64_t)size) != crc) { set_error("corrupted record payload at byte %lld", (long long)at); return HGB_ERR_INVALID; }
    offsets[n] = at + 12;
    lengths[n] = (int64_t)size;
    ++n;
    at += 12 + (int64_t)size + 4;
  }
  *next = at;
  return n;
}
