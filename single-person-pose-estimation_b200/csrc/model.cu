// The stacked-hourglass network (model/hourglass.py:5-206) as a static execution plan.
//
// hgb_model_create() walks the same construction code path as the reference (front module,
// num_stacks hourglass modules, heads) and records
//   * the parameter table (Keras layer names, creation order; kernels stored OHWI so they are
//     the K-major GEMM operand directly),
//   * every activation tensor (bf16 NHWC) with its offset in one caller-owned arena,
//   * a forward op list and, per segment (front, stack 0..S-1), a backward op list.
// Nothing is allocated or traced at run time; forward/backward replay the lists on a stream.
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "comm.cuh"
#include "conv_gemm.cuh"
#include "layer_kernels.cuh"

using namespace hgb;

namespace {

constexpr size_t kAlign = 1024;
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Act {
  int n, h, w, c;
  size_t off;  // bytes into the arena
  CUtensorMap tmap;
  bool has_tmap = false;
  // Deferred BatchNorm: the BN output is never materialised.  `bn_of` >= 0 marks a virtual tensor = BN(bn_of)(acts[src]);
  // its only consumers are 1x1 convolutions (forward GEMM and weight gradient), which normalise the operand tile in
  // shared memory right after TMA delivers it (bit-identical to the stand-alone BN pass: same fp32 fma, same rounding).
  int bn_of = -1, src = -1;
};

struct ConvL {
  std::string name;
  int ksize;            // GEMM view: 1 or 3 (the 7x7 stem runs as a 1x1 over extracted patches)
  int taps;             // ksize^2
  int cin, cout;        // valid channels of the GEMM view (stem: 147)
  int cin_pad, cout_pad;
  int real_k, real_cin; // Keras kernel shape (stem: 7, 3)
  int relu;
  int h, w;             // spatial size the GEMM runs at
  int64_t w_off, b_off; // float offsets into the parameter buffer
  size_t wf_off, wd_off;
  bool has_wd;
  CUtensorMap tm_wf, tm_wd;
  CUtensorMap tm_wf64, tm_wd64;   // 64-row boxes of the same operands: halves of a CTA pair (3x3, 128 -> 128)
  bool has_w64 = false;
  double flops;
  int seg;
  int y_act = -1, in_act = -1, z_act = -1, in_bn = -1;
  int dw = -1;          // mobile: index of the depthwise stage that feeds this (pointwise) convolution
};

// depthwise stage of a SeparableConv2D (mobile bottleneck): per-channel k x k stencil, fp32 weights [k*k][c], no bias
struct DwL {
  std::string name;
  int k, c, h, w;
  int64_t w_off;        // float offset into the parameter buffer
  int in_act, out_act;
};

struct BNL {
  std::string name;
  int c;
  bool has_writer = false;   // deferred BN: one consumer (the first) stores saved statistics / moving averages
  int64_t gamma_off, beta_off, mm_off, mv_off;  // float offsets into the parameter buffer
  size_t sums_off, bsums_off, saved_off;        // byte offsets into the arena
};

struct ParamT {
  std::string name;
  int rank;
  int64_t dims[4];
  int64_t off;
  int trainable;
};

enum OpType {
  F_IM2COL, F_CONV, F_BN, F_POOL, F_UPADD, F_HEAD,
  B_BN_REDUCE, B_BN_APPLY, B_WGRAD, B_DGRAD, B_RELU_MASK, B_COLSUM, B_POOL, B_UPADD, B_HEAD,
  F_DW, B_DW_DGRAD, B_DW_WGRAD      // mobile variant: depthwise stencil (conv = index into dws)
};

struct Op {
  OpType type;
  int conv = -1, bn = -1;
  int a0 = -1, a1 = -1, a2 = -1, a3 = -1;
  int flag = 0;
  int lane = -1;   // execution lane (CUDA stream); -1 = the lane current at emission
  // B_DGRAD with the BatchNorm backward of its input gradient fused in (the stand-alone B_BN_APPLY is gone): the GEMM
  // reads dz = act fz and y = act fy of BatchNorm fbn, computes dp in shared memory and stores it into a0
  int fbn = -1, fz = -1, fy = -1;
};

// Execution lanes.  The hourglass is not a chain: the skip ("short") bottleneck of every level is independent
// of the whole deeper sub-hourglass, and weight gradients are leaves of the backward graph.  Every op carries a
// lane; lanes map to CUDA streams and cross-lane dependencies (derived from the ops' read/write sets when the
// plan is built) become events.  At small per-GPU batches the low-resolution levels are latency-bound chains
// of tiny kernels; running them beside the 64x64 skip branch and beside the weight gradients hides them.
constexpr int kLaneMain = 0;    // the critical chain: down path, bottom, merges, heads
constexpr int kLaneWgrad = 1;   // weight / bias gradients of the main chain
constexpr int kLaneShort0 = 2;  // + level (f1, f2, f4, f8): the skip bottleneck of that level
constexpr int kLaneWgradX = 6;  // + i: extra weight-gradient lanes (leaf ops of the main chain are dealt round-robin)
constexpr int kMaxWgradLanes = 3;
constexpr int kNumLanes = 6 + kMaxWgradLanes - 1;
inline bool is_skip_lane(int l) { return l >= kLaneShort0 && l < kLaneShort0 + 4; }
inline bool is_wgrad_lane(int l) { return l == kLaneWgrad || l >= kLaneWgradX; }
constexpr int kScratchKinds = 8;                  // dp3 | dz(mid) | dp(mid, conv2) | dp(mid, conv1) | mobile: dt of conv3 / conv2 / conv1 / skip
constexpr int kScratchSets = 2 + 4;               // main chain ping-pong + one per skip lane

struct Range { int space; size_t lo, hi; };       // space 0: arena bytes, 1: gradient buffer (floats)
struct Dep { int lane, idx; };
struct SchedOp {
  int seg, idx;                 // position in fwd_ops / bwd_ops
  std::vector<Dep> deps;        // cross-lane: wait for the event of op `idx` (sequence index) of lane `lane`
  bool signal = false;          // some later op of another lane waits for this one
};

struct BneckRec {
  int x, out;
  int skip_conv, s_act;  // -1 when identity
  int c1, c2, c3, bn1, bn2, bn3;
  int y1, z1, y2, z2, y3;
};

struct StackRec {
  BneckRec down[4], bottom[3], shortb[4], merged[4];
  int pool_in[4], pool_out[4];   // pools after f1,f2,f4 and the bottom pool after f8
  int up_a[4];                   // upsample+add outputs (f8,f4,f2,f1 order)
  int up_low[4];                 // the lower-resolution input of each merge
  int conv_h, bn_h, y_h, z_h, conv_p, logits, pbf, conv2, conv3, t1, x_in, next;
};

}  // namespace

struct hgb_model {
  hgb_model_config cfg;
  int device;
  int S, C, K, B;
  int hm_h, hm_w;

  std::vector<Act> acts;
  std::vector<ConvL> convs;
  std::vector<DwL> dws;
  std::vector<BNL> bns;
  std::vector<ParamT> params;
  std::vector<std::vector<Op>> fwd_ops, bwd_ops;  // per segment
  std::vector<int64_t> seg_begin, seg_end;        // trainable float ranges

  int64_t n_train = 0, n_nontrain = 0;            // exact scalar counts
  int64_t train_floats = 0, nontrain_floats = 0;  // padded region sizes
  size_t arena_bytes = 0;
  size_t zero_off = 0, zero_bytes = 0;
  std::vector<size_t> heat_off, dldp_off;
  size_t lossws_off = 0, sync_off = 0;
  int sync_max_elems = 0;
  int col_act = -1;

  float* p_params = nullptr;
  float* p_grads = nullptr;
  float* p_m = nullptr;
  float* p_v = nullptr;
  uint8_t* p_arena = nullptr;
  bool maps_ready = false;
  bool fwd_training_done = false;
  int64_t launches = 0;
  // live timing of one conv class (bench.py roofline): CUDA event pairs around matching launches
  int prof_on = 0, prof_type = -1, prof_k = 0, prof_cin = 0, prof_cout = 0, prof_h = 0;
  std::vector<cudaEvent_t> prof_ev;
  size_t prof_used = 0;
  double prof_flops = 0;
  int prof_all = 0;
  std::vector<Op> prof_ops;

  // ---- lanes: forward / backward sequences with cross-lane dependencies (build_schedule), streams + events
  std::vector<SchedOp> fwd_seq, bwd_seq;           // fwd: segments 0..S; bwd: segments S..0
  std::vector<int> bwd_seq_begin;                  // first sequence index of each segment's backward ops
  cudaStream_t lane_stream[kNumLanes] = {nullptr};
  std::vector<cudaEvent_t> fwd_ev, bwd_ev;         // one per sequence op that signals another lane
  cudaEvent_t fork_ev = nullptr, join_ev[kNumLanes] = {nullptr};
  bool lanes_ready = false;
  bool pdl_suppressed = false;   // set while a multi-lane backward pass is being issued
  int bwd_unjoined_from = 1 << 30;   // first backward sequence index issued since the lanes were last joined into the caller
  bool lane_dirty[kNumLanes] = {false};   // lane has work the caller's stream has not been ordered after yet
  int num_sms = 0;

  // ---- data parallelism (hgb_model_set_comm): gradient buckets are all-reduced over `comm`; with sync_bn the BatchNorm
  // sums of every layer are all-reduced too, so N ranks compute exactly the single-device step of the concatenated batch
  hgb_comm* comm = nullptr;
  int sync_bn = 0;
  int stat_ranks() const { return (comm && sync_bn) ? comm->nranks : 1; }

  // ---- build-time state
  size_t arena_cur = 0;
  int bn_counter = 0;
  int cur_seg = 0;
  int cur_lane = kLaneMain, cur_set = 0, main_set = 0;
  size_t scratch_off[kScratchSets * kScratchKinds] = {0}, scratch_bytes[kScratchSets * kScratchKinds] = {0};
  size_t grad_base = 0, grad_cur = 0, grad_max = 0;
  bool sizing_pass = true;

  size_t arena_alloc(size_t bytes) {
    const size_t off = arena_cur;
    arena_cur += align_up(bytes, kAlign);
    return off;
  }
  int new_act(int n, int h, int w, int c) {
    Act a;
    a.n = n; a.h = h; a.w = w; a.c = c;
    a.off = arena_alloc((size_t)n * h * w * c * 2);
    acts.push_back(a);
    return (int)acts.size() - 1;
  }
  int alias_act(int n, int h, int w, int c, size_t off) {
    Act a;
    a.n = n; a.h = h; a.w = w; a.c = c; a.off = off;
    acts.push_back(a);
    return (int)acts.size() - 1;
  }
  int64_t add_param(const std::string& name, int rank, const int64_t* dims, int trainable) {
    ParamT p;
    p.name = name; p.rank = rank; p.trainable = trainable;
    int64_t n = 1;
    for (int i = 0; i < 4; ++i) { p.dims[i] = i < rank ? dims[i] : 1; if (i < rank) n *= dims[i]; }
    if (trainable) {
      p.off = train_floats;
      train_floats += (n + 3) / 4 * 4;
      n_train += n;
    } else {
      p.off = nontrain_floats;  // relocated behind the trainable region in finalize()
      nontrain_floats += (n + 3) / 4 * 4;
      n_nontrain += n;
    }
    params.push_back(p);
    return p.off;
  }

  int add_conv(const std::string& name, int in_act, int real_k, int real_cin, int cout, int relu, bool need_dgrad,
               const char* kernel_leaf = "/kernel") {
    const Act& in = acts[in_act];
    ConvL c;
    c.name = name;
    c.real_k = real_k; c.real_cin = real_cin;
    if (real_k == 7) { c.ksize = 1; c.taps = 1; c.cin = 147; c.cin_pad = 192; }
    else { c.ksize = real_k; c.taps = real_k * real_k; c.cin = real_cin; c.cin_pad = (real_cin + 63) / 64 * 64; }
    c.cout = cout;
    c.cout_pad = (cout + 63) / 64 * 64;
    c.relu = relu;
    c.h = in.h; c.w = in.w;
    const int64_t kd[4] = {real_k, real_k, real_cin, cout};
    c.w_off = add_param(name + kernel_leaf, 4, kd, 1);
    const int64_t bd[1] = {cout};
    c.b_off = add_param(name + "/bias", 1, bd, 1);
    c.wf_off = arena_alloc((size_t)c.cout_pad * c.taps * c.cin_pad * 2);
    c.has_wd = need_dgrad;
    c.wd_off = need_dgrad ? arena_alloc((size_t)c.cin_pad * c.taps * c.cout_pad * 2) : 0;
    c.flops = 2.0 * (double)in.n * in.h * in.w * (double)real_k * real_k * real_cin * cout;
    c.seg = cur_seg;
    sync_max_elems = std::max(sync_max_elems, c.cout_pad * c.taps * c.cin_pad);
    convs.push_back(c);
    return (int)convs.size() - 1;
  }
  int add_bn(int c) {
    BNL b;
    b.name = bn_counter == 0 ? "batch_normalization" : "batch_normalization_" + std::to_string(bn_counter);
    ++bn_counter;
    b.c = c;
    const int64_t d[1] = {c};
    b.gamma_off = add_param(b.name + "/gamma", 1, d, 1);
    b.beta_off = add_param(b.name + "/beta", 1, d, 1);
    b.mm_off = add_param(b.name + "/moving_mean", 1, d, 0);
    b.mv_off = add_param(b.name + "/moving_variance", 1, d, 0);
    b.sums_off = b.bsums_off = b.saved_off = 0;  // assigned in finalize()
    bns.push_back(b);
    return (int)bns.size() - 1;
  }
  void emit_f(Op o) { if (o.lane < 0) o.lane = cur_lane; fwd_ops[cur_seg].push_back(o); }
  void emit_b(Op o) { if (o.lane < 0) o.lane = cur_lane; bwd_ops[cur_seg].push_back(o); }
  // weight / bias gradients are leaves: the main chain hands them to its side lane
  // (the leaf ops of the main chain are independent of each other: dealt round-robin over `wgrad_lanes` streams, so the short,
  //  latency-bound weight gradients of the low-resolution levels overlap instead of queueing behind one another)
  int wgrad_lanes = 1, leaf_rr = 0;
  int leaf_lane() {
    if (cur_lane != kLaneMain) return cur_lane;
    const int i = leaf_rr++ % wgrad_lanes;
    return i == 0 ? kLaneWgrad : kLaneWgradX + i - 1;
  }

  // conv (+ReLU) [+ BN] -> returns the tensor the next layer consumes; y/z report both stages
  int conv_unit(const std::string& name, int in_act, int real_k, int real_cin, int cout, int relu, bool bn, bool need_dgrad,
                int bn_res, int* conv_idx, int* bn_idx, int* y_act, int res1 = -1, int res2 = -1, bool defer_bn = false,
                bool separable = false) {
    int dwi = -1;
    if (separable) {
      // SeparableConv2D (hourglass.py:218-226): depthwise k x k on the input channels, then the pointwise 1x1 (bias, activation).
      // Keras variable order: depthwise_kernel (k,k,cin,1), pointwise_kernel (1,1,cin,cout), bias.
      DwL d;
      d.name = name; d.k = real_k; d.c = acts[in_act].c; d.h = acts[in_act].h; d.w = acts[in_act].w;
      const int64_t kd[4] = {real_k, real_k, real_cin, 1};
      d.w_off = add_param(name + "/depthwise_kernel", 4, kd, 1);
      d.in_act = in_act;
      d.out_act = new_act(acts[in_act].n, d.h, d.w, d.c);
      dws.push_back(d);
      dwi = (int)dws.size() - 1;
      Op o; o.type = F_DW; o.conv = dwi; o.a0 = in_act; o.a1 = d.out_act; emit_f(o);
      in_act = d.out_act;
      real_k = 1;
    }
    const int ci = add_conv(name, in_act, real_k, real_cin, cout, relu, need_dgrad, separable ? "/pointwise_kernel" : "/kernel");
    convs[ci].dw = dwi;
    const Act in = acts[in_act];
    const int y = new_act(in.n, in.h, in.w, convs[ci].cout_pad);
    Op o;
    // flag = 1 + index of the BatchNorm applied to the INPUT tile inside the kernel (0 = none); a0 = its source tensor
    o.type = F_CONV; o.conv = ci; o.a0 = in.bn_of >= 0 ? in.src : in_act; o.a1 = y; o.a2 = res1; o.a3 = res2;
    o.flag = in.bn_of >= 0 ? 1 + in.bn_of : 0;
    if (in.bn_of >= 0 && !bns[in.bn_of].has_writer) { o.flag |= 0x10000; bns[in.bn_of].has_writer = true; }
    int bi = -1, out = y;
    if (bn) { bi = add_bn(cout); o.bn = bi; }
    emit_f(o);
    convs[ci].y_act = y;
    convs[ci].in_act = in.bn_of >= 0 ? in.src : in_act;
    convs[ci].in_bn = in.bn_of;
    if (bn && defer_bn && !hgb::g_debug[14]) {
      Act v;   // virtual: dims only
      v.n = in.n; v.h = in.h; v.w = in.w; v.c = cout; v.off = 0; v.bn_of = bi; v.src = y;
      acts.push_back(v);
      out = (int)acts.size() - 1;
      convs[ci].z_act = -1;
    } else if (bn) {
      out = new_act(in.n, in.h, in.w, cout);
      Op b;
      b.type = F_BN; b.bn = bi; b.a0 = y; b.a1 = bn_res; b.a2 = out;
      emit_f(b);
      convs[ci].z_act = out;
    }
    if (conv_idx) *conv_idx = ci;
    if (bn_idx) *bn_idx = bi;
    if (y_act) *y_act = y;
    return out;
  }

  // model/hourglass.py:184-206
  BneckRec bottleneck(int x, int cout, const std::string& name) {
    BneckRec r;
    const int cin = acts[x].c;
    r.x = x;
    r.skip_conv = -1; r.s_act = -1;
    int skip = x;
    const bool sep = cfg.mobile != 0;   // bottleneck_block_mobile: every convolution of the block is a SeparableConv2D
    if (cin != cout) {
      skip = conv_unit(name + "_skip", x, 1, cin, cout, 1, false, true, -1, &r.skip_conv, nullptr, nullptr, -1, -1, false, sep);
      r.s_act = skip;
    }
    r.z1 = conv_unit(name + "_conv_1x1_1", x, 1, cin, cout / 2, 1, true, true, -1, &r.c1, &r.bn1, &r.y1, -1, -1, false, sep);
    r.z2 = conv_unit(name + "_conv_3x3_2", r.z1, 3, cout / 2, cout / 2, 1, true, true, -1, &r.c2, &r.bn2, &r.y2, -1, -1,
                     /*defer_bn=*/!sep, sep);   // only consumer: the 1x1 conv_1x1_3 (a depthwise stage cannot normalise its operand)
    r.out = conv_unit(name + "_conv_1x1_3", r.z2, 1, cout / 2, cout, 1, true, true, skip, &r.c3, &r.bn3, &r.y3, -1, -1, false, sep);
    return r;
  }
  int pool(int x) {
    const Act a = acts[x];
    const int o = new_act(a.n, a.h / 2, a.w / 2, a.c);
    // MaxPool2D folded into the BatchNorm that produces its input (same lane, the op emitted last): that kernel writes the
    // pooled tensor next to its output (F_BN flag bit 0: a3 is the pooled OUTPUT) and the pool kernel's re-read of a whole
    // 256-channel tensor disappears, together with one launch on the main chain.  hgb_debug_set(43, 1) keeps the pool kernel.
    Op* last = fwd_ops[cur_seg].empty() ? nullptr : &fwd_ops[cur_seg].back();
    // Training plans only: in inference the BatchNorm runs inside the preceding convolution's epilogue (fuse_inference_bn) and
    // the pre-BatchNorm tensor is never stored -- folding the pool into a BatchNorm op would bring that tensor back (batch 128,
    // 8 stacks: 21.3 vs 21.7 ms; hgb_debug_set(43, 2) folds there too).
    if ((cfg.training || hgb::g_debug[43] == 2) && hgb::g_debug[43] != 1 && last && last->type == F_BN && last->a2 == x && last->a3 < 0 && last->lane == cur_lane &&
        (a.h & (a.h - 1)) == 0 && (a.w & (a.w - 1)) == 0 && a.h >= 2 && a.w >= 2) {
      last->a3 = o;
      last->flag |= 1;
      return o;
    }
    Op p;
    p.type = F_POOL; p.a0 = x; p.a1 = o;
    emit_f(p);
    return o;
  }

  // ---------------- backward emission
  int scratch_act(int which, int n, int h, int w, int c) {
    const int id = cur_set * kScratchKinds + which;
    scratch_bytes[id] = std::max(scratch_bytes[id], (size_t)n * h * w * c * 2);
    return alias_act(n, h, w, c, scratch_off[id]);
  }
  int grad_act(int n, int h, int w, int c) {
    const size_t bytes = align_up((size_t)n * h * w * c * 2, kAlign);
    const size_t off = grad_base + grad_cur;
    grad_cur += bytes;
    grad_max = std::max(grad_max, grad_cur);
    return alias_act(n, h, w, c, off);
  }
  // mobile: the gradient of a separable convolution passes through its depthwise stage: the pointwise dgrad writes dt (gradient
  // wrt the depthwise output, scratch `kind`), the depthwise stencil with mirrored taps turns it into the input gradient
  // (+ residuals), and the depthwise weights get their own reduction.  Returns true when it handled the dgrad / wgrad pair.
  bool separable_bwd(int conv, int dp, int dgrad_out, int res1, int res2, int kind) {
    const int dwi = convs[conv].dw;
    if (dwi < 0) return false;
    const DwL& d = dws[dwi];
    const Act t = acts[d.out_act];
    const int dt = scratch_act(kind, t.n, t.h, t.w, t.c);
    Op o;
    o.type = B_DGRAD; o.conv = conv; o.a0 = dp; o.a1 = dt; emit_b(o);
    emit_wgrad(conv, dp, d.out_act);
    if (dgrad_out >= 0) {
      o = Op(); o.type = B_DW_DGRAD; o.conv = dwi; o.a0 = dt; o.a1 = dgrad_out; o.a2 = res1; o.a3 = res2; emit_b(o);
    }
    o = Op(); o.type = B_DW_WGRAD; o.conv = dwi; o.a0 = dt; o.a1 = d.in_act; o.lane = leaf_lane(); emit_b(o);
    return true;
  }
  void bn_conv_bwd(int bn, int conv, int dz, int y, int dp, int x_in, int dgrad_out, int res1, int res2, int dt_kind = 4) {
    Op o;
    o = Op(); o.type = B_BN_REDUCE; o.bn = bn; o.a0 = dz; o.a1 = y; emit_b(o);
    o = Op(); o.type = B_BN_APPLY; o.bn = bn; o.conv = conv; o.a0 = dz; o.a1 = y; o.a2 = dp; emit_b(o);
    if (separable_bwd(conv, dp, dgrad_out, res1, res2, dt_kind)) return;
    // the dgrad continues the chain; the weight gradient is a leaf of the backward graph (side lane)
    const bool wfirst = hgb::g_debug[10] != 0;
    if (wfirst) emit_wgrad(conv, dp, x_in);
    if (dgrad_out >= 0) {
      o = Op(); o.type = B_DGRAD; o.conv = conv; o.a0 = dp; o.a1 = dgrad_out; o.a2 = res1; o.a3 = res2; emit_b(o);
    }
    if (!wfirst) emit_wgrad(conv, dp, x_in);
  }
  // weight gradient: a virtual (deferred-BN) input is read from its source tensor and normalised inside the kernel
  void emit_wgrad(int conv, int dp, int x_in) {
    Op o;
    const Act& x = acts[x_in];
    o.type = B_WGRAD; o.conv = conv; o.a0 = dp; o.a1 = x.bn_of >= 0 ? x.src : x_in; o.bn = x.bn_of; o.lane = leaf_lane();
    emit_b(o);
  }
  // g_out: gradient wrt the block output (read; masked in place when the skip is a conv);
  // g_x: gradient wrt the block input (written); extra: one more tensor summed into g_x.
  void bottleneck_bwd(const BneckRec& r, int g_out, int g_x, int extra) {
    const Act o = acts[r.out];
    const int cmid = convs[r.c1].cout;
    // the main chain alternates between two scratch sets block by block, so the weight gradients of one block
    // (side lane, reading its dp tensors) never hold up the next block; each dp tensor has its own buffer
    if (cur_lane == kLaneMain) { cur_set = main_set; main_set ^= 1; }
    const int dp3 = scratch_act(0, o.n, o.h, o.w, o.c);
    const int dzm = scratch_act(1, o.n, o.h, o.w, cmid);
    const int dpm = scratch_act(2, o.n, o.h, o.w, cmid);
    const int dpm1 = scratch_act(3, o.n, o.h, o.w, cmid);
    bn_conv_bwd(r.bn3, r.c3, g_out, r.y3, dp3, r.z2, dzm, -1, -1, 4);
    bn_conv_bwd(r.bn2, r.c2, dzm, r.y2, dpm, r.z1, dzm, -1, -1, 5);
    if (r.skip_conv >= 0) {
      const bool sep = convs[r.c1].dw >= 0;
      if (sep) {
        // mobile: both branches end in a depthwise stage; the skip's input gradient is written first, conv_1x1_1's is added to it
        Op m;
        m = Op(); m.type = B_BN_REDUCE; m.bn = r.bn1; m.a0 = dzm; m.a1 = r.y1; emit_b(m);
        m = Op(); m.type = B_BN_APPLY; m.bn = r.bn1; m.conv = r.c1; m.a0 = dzm; m.a1 = r.y1; m.a2 = dpm1; emit_b(m);
        m = Op(); m.type = B_RELU_MASK; m.conv = r.skip_conv; m.a0 = g_out; m.a1 = r.s_act; m.flag = 1; emit_b(m);
        separable_bwd(r.skip_conv, g_out, g_x, extra, -1, 7);
        separable_bwd(r.c1, dpm1, g_x, g_x >= 0 ? g_x : -1, -1, 6);
      } else {
        bn_conv_bwd(r.bn1, r.c1, dzm, r.y1, dpm1, r.x, -1, -1, -1);
        Op m;
        m.type = B_RELU_MASK; m.conv = r.skip_conv; m.a0 = g_out; m.a1 = r.s_act; m.flag = 1; emit_b(m);
        if (g_x >= 0) {
          m = Op(); m.type = B_DGRAD; m.conv = r.skip_conv; m.a0 = g_out; m.a1 = g_x; m.a2 = extra; emit_b(m);
          m = Op(); m.type = B_DGRAD; m.conv = r.c1; m.a0 = dpm1; m.a1 = g_x; m.a2 = g_x; emit_b(m);
        }
        emit_wgrad(r.skip_conv, g_out, r.x);
      }
    } else {
      bn_conv_bwd(r.bn1, r.c1, dzm, r.y1, dpm1, r.x, g_x, g_out, extra, 6);
    }
  }
  void linear_conv_bwd(int conv, int dp, int x_in, int dgrad_out, int res1) {
    Op o;
    if (dgrad_out >= 0) {
      o.type = B_DGRAD; o.conv = conv; o.a0 = dp; o.a1 = dgrad_out; o.a2 = res1; emit_b(o);
    }
    // bias gradient = column sums of dp.  conv_1x1_2 and conv_1x1_3 of a stack see the SAME gradient tensor (the three-way
    // Add of hourglass.py:91): one pass writes both bias gradients (a1 = the second convolution)
    bool merged = false;
    for (auto it = bwd_ops[cur_seg].rbegin(); it != bwd_ops[cur_seg].rend() && !merged; ++it)
      if (it->type == B_COLSUM && it->a0 == dp && it->a1 < 0 && convs[it->conv].cout == convs[conv].cout && !hgb::g_debug[29]) {
        it->a1 = conv;
        merged = true;
      }
    if (!merged) { o = Op(); o.type = B_COLSUM; o.conv = conv; o.a0 = dp; o.lane = leaf_lane(); emit_b(o); }
    emit_wgrad(conv, dp, x_in);
  }
};

namespace {

int build(hgb_model* m) {
  const hgb_model_config& cfg = m->cfg;
  const int B = cfg.batch, C = cfg.num_channels, S = cfg.num_stacks;
  m->fwd_ops.assign(S + 1, {});
  m->bwd_ops.assign(S + 1, {});
  // weight-gradient lanes: hgb_debug_set(35, n) before the plan is created, n = 1..3 (0 = the default for this batch)
  m->wgrad_lanes = hgb::g_debug[35] > 0 ? std::min(hgb::g_debug[35], kMaxWgradLanes) : 1;
  m->seg_begin.assign(S + 1, 0);
  m->seg_end.assign(S + 1, 0);

  // ---------------- front module (hourglass.py:54-68)
  m->cur_seg = 0;
  m->col_act = m->new_act(B, cfg.in_h / 2, cfg.in_w / 2, 192);
  { Op o; o.type = F_IM2COL; o.a0 = m->col_act; m->emit_f(o); }
  int conv0, bn0, y0;
  const int z0 = m->conv_unit("front_conv_1x1_1", m->col_act, 7, 3, 64, 1, true, false, -1, &conv0, &bn0, &y0);
  BneckRec fb1 = m->bottleneck(z0, C / 2, "front_bottleneck_1");
  const int fpool = m->pool(fb1.out);
  BneckRec fb2 = m->bottleneck(fpool, C / 2, "front_bottleneck_2");
  BneckRec fb3 = m->bottleneck(fb2.out, C, "front_bottleneck_3");
  m->seg_end[0] = m->train_floats;

  // ---------------- stacks (hourglass.py:35-52, 71-181)
  std::vector<StackRec> stacks(S);
  int x = fb3.out;
  static const char* fnames[4] = {"f1", "f2", "f4", "f8"};
  for (int s = 0; s < S; ++s) {
    m->cur_seg = 1 + s;
    m->seg_begin[1 + s] = m->train_floats;
    StackRec& r = stacks[s];
    const std::string hg = "hg" + std::to_string(s);
    r.x_in = x;
    int cur = x;
    for (int l = 0; l < 4; ++l) {  // create_downsample_blocks
      r.down[l] = m->bottleneck(cur, C, hg + "_downsample_" + fnames[l]);
      r.pool_in[l] = r.down[l].out;
      if (l < 3) { r.pool_out[l] = m->pool(r.down[l].out); cur = r.pool_out[l]; }
    }
    r.pool_out[3] = m->pool(r.down[3].out);  // bottom_block
    cur = r.pool_out[3];
    for (int i = 0; i < 3; ++i) {
      r.bottom[i] = m->bottleneck(cur, C, hg + "_downsample_f8_" + std::to_string(i + 1));
      cur = r.bottom[i].out;
    }
    for (int u = 0; u < 4; ++u) {  // connect_downsample_upsample for f8, f4, f2, f1
      const int l = 3 - u;
      const std::string nm = hg + "_upsample_" + fnames[l];
      m->cur_lane = kLaneShort0 + l;   // independent of everything below this level
      r.shortb[u] = m->bottleneck(r.down[l].out, C, nm + "_short");
      m->cur_lane = kLaneMain;
      const Act sa = m->acts[r.shortb[u].out];
      r.up_low[u] = cur;
      Op* last = m->fwd_ops[m->cur_seg].empty() ? nullptr : &m->fwd_ops[m->cur_seg].back();
      // (batch 256: 147.9 -> 146.5 ms per step.  At batch 32 the step is latency-bound and the folded kernel sits on the critical
      //  path behind the deep levels, where the stand-alone merge was shorter: 24.97 vs 25.13 ms -- so only above batch 48;
      //  hgb_debug_set(42, 2) forces it)
      // Inference keeps the stand-alone merge: there the closing BatchNorm runs inside conv_1x1_3's epilogue (fuse_inference_bn),
      // which a BatchNorm op with a folded merge would block (batch 128, 8 stacks: 21.1 vs 21.5 ms)
      const bool fold_upadd = hgb::g_debug[42] == 2 || (hgb::g_debug[42] == 0 && cfg.batch > 48 && cfg.training);
      if (fold_upadd && last && last->type == F_BN && last->a2 == r.shortb[u].out && last->a1 >= 0 && last->a3 < 0) {
        // UpSampling2D + Add (hourglass.py:152-154) folded into the skip bottleneck's closing BatchNorm: that kernel already
        // reads y3 and the skip and writes the block output -- it now adds the upsampled lower level too and writes the MERGE
        // input; the block's own output is never stored (nothing reads it: its backward pass needs y3 only).  One write and one
        // read of a 256-channel tensor less per level and stack than the stand-alone merge kernel (hgb_debug_set(42, 1)).
        last->a3 = cur;
        r.up_a[u] = r.shortb[u].out;
      } else {
        r.up_a[u] = m->new_act(sa.n, sa.h, sa.w, sa.c);
        Op o; o.type = F_UPADD; o.a0 = r.shortb[u].out; o.a1 = cur; o.a2 = r.up_a[u]; m->emit_f(o);
      }
      r.merged[u] = m->bottleneck(r.up_a[u], C, nm + "_merged");
      cur = r.merged[u].out;
    }
    // create_heads
    r.z_h = m->conv_unit(hg + "_conv_1x1_1", cur, 1, C, C, 1, true, true, -1, &r.conv_h, &r.bn_h, &r.y_h, -1, -1,
                         /*defer_bn=*/true);   // consumers: conv_1x1_predict and conv_1x1_2, both 1x1
    r.logits = m->conv_unit(hg + "_conv_1x1_predict", r.z_h, 1, C, cfg.num_classes, 0, false, true, -1, &r.conv_p, nullptr, nullptr);
    const Act la = m->acts[r.logits];
    r.pbf = m->new_act(la.n, la.h, la.w, 64);
    { Op o; o.type = F_HEAD; o.flag = s; o.a0 = r.logits; o.a1 = r.pbf; m->emit_f(o); }
    r.conv2 = r.conv3 = r.t1 = r.next = -1;
    if (s + 1 < S) {  // the last stack's re-injection branch is pruned by Keras (not on a path to an output)
      r.t1 = m->conv_unit(hg + "_conv_1x1_2", r.z_h, 1, C, C, 0, false, true, -1, &r.conv2, nullptr, nullptr, r.x_in);
      r.next = m->conv_unit(hg + "_conv_1x1_3", r.pbf, 1, cfg.num_classes, C, 0, false, true, -1, &r.conv3, nullptr, nullptr, r.t1);
      x = r.next;
    }
    m->seg_end[1 + s] = m->train_floats;
  }
  m->seg_begin[0] = 0;

  // ---------------- per-step float state (zeroed each training step) + saved BN statistics
  size_t zf = 0;
  for (auto& b : m->bns) { zf += 4 * (size_t)b.c; }
  m->zero_bytes = align_up(zf * sizeof(float), kAlign);
  m->zero_off = m->arena_alloc(m->zero_bytes);
  size_t saved_off = m->arena_alloc(align_up(zf / 2 * sizeof(float), kAlign));
  size_t zcur = m->zero_off;
  for (auto& b : m->bns) {
    b.sums_off = zcur; zcur += 2 * (size_t)b.c * sizeof(float);
    b.bsums_off = zcur; zcur += 2 * (size_t)b.c * sizeof(float);
    b.saved_off = saved_off; saved_off += 2 * (size_t)b.c * sizeof(float);
  }
  // heat maps (f32) and loss gradients per stack
  const int hh = cfg.in_h / 4, hw = cfg.in_w / 4;
  m->hm_h = hh; m->hm_w = hw;
  m->heat_off.resize(S);
  m->dldp_off.resize(S);
  for (int s = 0; s < S; ++s) m->heat_off[s] = m->arena_alloc((size_t)B * hh * hw * cfg.num_classes * 4);
  m->sync_off = m->arena_alloc(m->convs.size() * sizeof(WeightSyncEntry));

  if (cfg.training) {
    for (int s = 0; s < S; ++s) m->dldp_off[s] = m->arena_alloc((size_t)B * hh * hw * cfg.num_classes * 4);
    m->lossws_off = m->arena_alloc((size_t)hgb_loss_workspace_bytes(B, cfg.num_classes));
    // stack-input gradients ping-pong; everything else of a stack's backward lives in a region reused by all stacks
    const Act xa = m->acts[fb3.out];
    int gx[2];
    gx[0] = m->new_act(xa.n, xa.h, xa.w, xa.c);
    gx[1] = m->new_act(xa.n, xa.h, xa.w, xa.c);
    // Two passes over the backward emission: the first only sizes the scratch / gradient regions.
    const size_t acts_mark = m->acts.size();
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) {
        m->acts.resize(acts_mark);
        for (auto& v : m->bwd_ops) v.clear();
        for (int i = 0; i < kScratchSets * kScratchKinds; ++i)
          if (m->scratch_bytes[i]) m->scratch_off[i] = m->arena_alloc(m->scratch_bytes[i]);
        m->grad_base = m->arena_alloc(m->grad_max);
      }
      m->cur_lane = kLaneMain; m->cur_set = 0; m->main_set = 0; m->leaf_rr = 0;
      for (int s = S - 1; s >= 0; --s) {
        m->cur_seg = 1 + s;
        m->grad_cur = 0;
        const StackRec& r = stacks[s];
        const bool last = (s + 1 == S);
        const int g_next = last ? -1 : gx[(s + 1) & 1];
        const Act ha = m->acts[r.z_h];
        const int g_zh = m->grad_act(ha.n, ha.h, ha.w, ha.c);
        const int g_logits = m->grad_act(ha.n, ha.h, ha.w, 64);
        int g_p = -1;
        if (!last) {
          g_p = m->grad_act(ha.n, ha.h, ha.w, 64);
          m->linear_conv_bwd(r.conv3, g_next, r.pbf, g_p, -1);
          m->linear_conv_bwd(r.conv2, g_next, r.z_h, g_zh, -1);
        }
        { Op o; o.type = B_HEAD; o.flag = s; o.a0 = g_p; o.a1 = g_logits; m->emit_b(o); }
        m->linear_conv_bwd(r.conv_p, g_logits, r.z_h, g_zh, last ? -1 : g_zh);
        int g_cur = m->grad_act(ha.n, ha.h, ha.w, ha.c);  // gradient wrt the last merged block's output
        {
          m->cur_set = m->main_set; m->main_set ^= 1;
          const int dp_h = m->scratch_act(0, ha.n, ha.h, ha.w, ha.c);
          m->bn_conv_bwd(r.bn_h, r.conv_h, g_zh, r.y_h, dp_h, r.merged[3].out, g_cur, -1, -1);
        }
        int g_f[4];  // gradients wrt the down-path features f1,f2,f4,f8
        for (int u = 3; u >= 0; --u) {
          const int l = 3 - u;
          const Act aa = m->acts[r.up_a[u]];
          const int g_a = m->grad_act(aa.n, aa.h, aa.w, aa.c);
          m->bottleneck_bwd(r.merged[u], g_cur, g_a, -1);
          const Act lo = m->acts[r.up_low[u]];
          const int g_low = m->grad_act(lo.n, lo.h, lo.w, lo.c);
          { Op o; o.type = B_UPADD; o.a0 = g_a; o.a1 = g_low; m->emit_b(o); }
          g_f[l] = m->grad_act(aa.n, aa.h, aa.w, aa.c);
          m->cur_lane = kLaneShort0 + l; m->cur_set = 2 + l;
          m->bottleneck_bwd(r.shortb[u], g_a, g_f[l], -1);
          m->cur_lane = kLaneMain;
          g_cur = g_low;
        }
        for (int i = 2; i >= 0; --i) {
          const Act ba = m->acts[r.bottom[i].x];
          const int g_in = m->grad_act(ba.n, ba.h, ba.w, ba.c);
          m->bottleneck_bwd(r.bottom[i], g_cur, g_in, -1);
          g_cur = g_in;
        }
        // g_cur = gradient wrt the bottom pool's output
        for (int l = 3; l >= 0; --l) {
          { Op o; o.type = B_POOL; o.a0 = r.pool_in[l]; o.a1 = g_cur; o.a2 = g_f[l]; o.flag = 1; m->emit_b(o); }
          if (l > 0) {
            const Act pa = m->acts[r.down[l].x];
            const int g_in = m->grad_act(pa.n, pa.h, pa.w, pa.c);
            m->bottleneck_bwd(r.down[l], g_f[l], g_in, -1);
            g_cur = g_in;
          } else {
            m->bottleneck_bwd(r.down[0], g_f[0], gx[s & 1], g_next);
          }
        }
      }
      // front module
      m->cur_seg = 0;
      m->grad_cur = 0;
      {
        const Act a2 = m->acts[fb3.x];
        const int g_b2 = m->grad_act(a2.n, a2.h, a2.w, a2.c);
        m->bottleneck_bwd(fb3, gx[0], g_b2, -1);
        const Act ap = m->acts[fb2.x];
        const int g_pool = m->grad_act(ap.n, ap.h, ap.w, ap.c);
        m->bottleneck_bwd(fb2, g_b2, g_pool, -1);
        const Act a1 = m->acts[fb1.out];
        const int g_b1 = m->grad_act(a1.n, a1.h, a1.w, a1.c);
        { Op o; o.type = B_POOL; o.a0 = fb1.out; o.a1 = g_pool; o.a2 = g_b1; o.flag = 0; m->emit_b(o); }
        const Act az = m->acts[z0];
        const int g_z0 = m->grad_act(az.n, az.h, az.w, az.c);
        m->bottleneck_bwd(fb1, g_b1, g_z0, -1);
        m->cur_set = m->main_set; m->main_set ^= 1;
        const int dp0 = m->scratch_act(0, az.n, az.h, az.w, az.c);
        m->bn_conv_bwd(bn0, conv0, g_z0, y0, dp0, m->col_act, -1, -1, -1);
      }
    }
  }
  // fuse each BatchNorm-backward reduction into the dgrad GEMM that produced its dz, when that GEMM is the
  // most recent writer of the tensor (pool / upsample gradients keep the stand-alone reduction)
  if (cfg.training && !hgb::g_debug[4]) {
    for (auto& ops : m->bwd_ops) {
      std::vector<Op> out;
      for (const Op& o : ops) {
        if (o.type == B_BN_REDUCE) {
          int w = -1;
          for (int i = (int)out.size() - 1; i >= 0 && w < 0; --i) {
            const Op& q = out[i];
            const int written = q.type == B_DGRAD ? q.a1 : q.type == B_POOL ? q.a2 : q.type == B_UPADD ? q.a1
                              : q.type == B_BN_APPLY ? q.a2 : q.type == B_RELU_MASK ? q.a0 : q.type == B_HEAD ? q.a1
                              : q.type == B_DW_DGRAD ? q.a1 : -1;
            if (written >= 0 && (written == o.a0 || m->acts[written].off == m->acts[o.a0].off)) w = i;
          }
          if (w >= 0 && out[w].type == B_DGRAD && out[w].a1 == o.a0 && out[w].bn < 0) {
            out[w].bn = o.bn;
            out[w].flag = o.a1 + 1;   // y act (+1 so that 0 means none)
            continue;                 // the stand-alone reduction disappears
          }
          // the gradient of a max-pool / of an upsample-add merge: those kernels accumulate the statistics of what they
          // store (y act in a3 / a2), and the reduction's read of dz goes away with its launch.  hgb_debug_set(44, 1) keeps
          // the stand-alone reduction.
          const int cbn = m->bns[o.bn].c;
          if (w >= 0 && !hgb::g_debug[44] && out[w].bn < 0 && out[w].lane == o.lane && 256 % (cbn / 8) == 0 && cbn <= 2048) {
            if (out[w].type == B_POOL && out[w].a2 == o.a0) { out[w].bn = o.bn; out[w].a3 = o.a1; continue; }
            if (out[w].type == B_UPADD && out[w].a1 == o.a0) { out[w].bn = o.bn; out[w].a2 = o.a1; continue; }
          }
        }
        out.push_back(o);
      }
      ops.swap(out);
    }
  }
  // fuse each BatchNorm-backward apply into the 1x1 dgrad GEMM that consumes its dp (APPLY, DGRAD adjacent in the list;
  // the weight gradient, which also reads dp, comes after the dgrad): the GEMM's transform warps compute dp from dz and y in
  // shared memory and store it on the side, so the stand-alone pass (r dz, r y, w dp) and the GEMM's own read of dp
  // collapse into r dz, r y, w dp.  3x3 dgrads (strip reuse) and the front module's projected-skip blocks keep the pass.
  if (cfg.training && !hgb::g_debug[26]) {
    for (auto& ops : m->bwd_ops) {
      std::vector<Op> out;
      for (size_t i = 0; i < ops.size(); ++i) {
        const Op& o = ops[i];
        if (o.type == B_BN_APPLY && i + 1 < ops.size()) {
          const Op& d = ops[i + 1];
          const ConvL& c = m->convs[o.conv];
          if (d.type == B_DGRAD && d.conv == o.conv && d.a0 == o.a2 && d.lane == o.lane && d.fbn < 0 &&
              conv_gemm_supports_bn_bwd(c.ksize, c.cout_pad, c.cin_pad) && c.cout_pad == m->bns[o.bn].c &&
              m->acts[o.a0].off != m->acts[o.a2].off) {
            Op f = d;
            f.fbn = o.bn; f.fz = o.a0; f.fy = o.a1;
            out.push_back(f);
            ++i;
            continue;
          }
        }
        out.push_back(o);
      }
      ops.swap(out);
    }
  }
  // relocate the non-trainable parameters behind the trainable region
  for (auto& p : m->params)
    if (!p.trainable) p.off += m->train_floats;
  for (auto& b : m->bns) { b.mm_off += m->train_floats; b.mv_off += m->train_floats; }
  m->arena_bytes = m->arena_cur + kAlign;
  return HGB_OK;
}



// ---------------------------------------------------------------------------------------- lanes
inline void add_act(const hgb_model* m, std::vector<Range>& v, int a) {
  if (a < 0) return;
  const Act& t = m->acts[a];
  v.push_back({0, t.off, t.off + (size_t)t.n * t.h * t.w * t.c * 2});
}
inline void add_arena(std::vector<Range>& v, size_t off, size_t bytes) { v.push_back({0, off, off + bytes}); }
inline void add_grad(std::vector<Range>& v, int64_t off, int64_t n) { v.push_back({1, (size_t)off, (size_t)(off + n)}); }

// Everything an op reads (r) and writes (w) that another op of the same pass may touch: activations, scratch and
// gradient tensors, BN statistics, heat maps, parameter gradients.  (Parameters and bf16 weight operands are
// read-only during a pass; the BN moving statistics have a single writer.)  KEEP IN STEP WITH run_op_impl.
void op_access(const hgb_model* m, const Op& o, std::vector<Range>& r, std::vector<Range>& w) {
  r.clear(); w.clear();
  const size_t hm_bytes = (size_t)m->B * m->hm_h * m->hm_w * m->K * 4;
  switch (o.type) {
    case F_IM2COL: add_act(m, w, o.a0); break;
    case F_CONV:
      add_act(m, r, o.a0); add_act(m, r, o.a2); add_act(m, r, o.a3); add_act(m, w, o.a1);
      if (o.bn >= 0) add_arena(w, m->bns[o.bn].sums_off, 2 * (size_t)m->bns[o.bn].c * 4);
      if (o.flag & 0xffff) {
        const BNL& b = m->bns[(o.flag & 0xffff) - 1];
        add_arena(r, b.sums_off, 2 * (size_t)b.c * 4);
        if (o.flag & 0x10000) {
          add_arena(w, b.saved_off, 2 * (size_t)b.c * 4);
          add_arena(w, b.sums_off, 2 * (size_t)b.c * 4);     // sync-BN: this launch all-reduces the sums in place first
        }
      }
      break;
    case F_BN:
      add_act(m, r, o.a0); add_act(m, r, o.a1); add_act(m, (o.flag & 1) ? w : r, o.a3); add_act(m, w, o.a2);
      add_arena(w, m->bns[o.bn].sums_off, 2 * (size_t)m->bns[o.bn].c * 4);   // (sync-BN all-reduces them in place)
      add_arena(r, m->bns[o.bn].sums_off, 2 * (size_t)m->bns[o.bn].c * 4);
      add_arena(w, m->bns[o.bn].saved_off, 2 * (size_t)m->bns[o.bn].c * 4);
      break;
    case F_POOL: add_act(m, r, o.a0); add_act(m, w, o.a1); break;
    case F_UPADD: add_act(m, r, o.a0); add_act(m, r, o.a1); add_act(m, w, o.a2); break;
    case F_HEAD: add_act(m, r, o.a0); add_act(m, w, o.a1); add_arena(w, m->heat_off[o.flag], hm_bytes); break;
    case B_BN_REDUCE:
      add_act(m, r, o.a0); add_act(m, r, o.a1);
      add_arena(w, m->bns[o.bn].bsums_off, 2 * (size_t)m->bns[o.bn].c * 4);
      break;
    case B_BN_APPLY: {
      const BNL& b = m->bns[o.bn];
      add_act(m, r, o.a0); add_act(m, r, o.a1); add_act(m, w, o.a2);
      add_arena(r, b.bsums_off, 2 * (size_t)b.c * 4);
      add_arena(w, b.bsums_off, 2 * (size_t)b.c * 4);        // (sync-BN all-reduces them in place)
      add_arena(r, b.saved_off, 2 * (size_t)b.c * 4);
      add_grad(w, b.gamma_off, b.c); add_grad(w, b.beta_off, b.c);
      add_grad(w, m->convs[o.conv].b_off, m->convs[o.conv].cout);
      break;
    }
    case B_WGRAD: {
      const ConvL& c = m->convs[o.conv];
      add_act(m, r, o.a0); add_act(m, r, o.a1);
      if (o.bn >= 0) add_arena(r, m->bns[o.bn].saved_off, 2 * (size_t)m->bns[o.bn].c * 4);
      add_grad(w, c.w_off, (int64_t)c.cout * c.taps * c.cin);
      break;
    }
    case B_DGRAD:
      add_act(m, r, o.a0); add_act(m, r, o.a2); add_act(m, r, o.a3); add_act(m, w, o.a1);
      if (o.fbn >= 0) {   // fused BatchNorm backward: dz and y in, dp (a0) out, plus everything B_BN_APPLY touches
        const BNL& b = m->bns[o.fbn];
        add_act(m, r, o.fz); add_act(m, r, o.fy); add_act(m, w, o.a0);
        add_arena(r, b.bsums_off, 2 * (size_t)b.c * 4);
        add_arena(w, b.bsums_off, 2 * (size_t)b.c * 4);        // (sync-BN all-reduces them in place)
        add_arena(r, b.saved_off, 2 * (size_t)b.c * 4);
        add_grad(w, b.gamma_off, b.c); add_grad(w, b.beta_off, b.c);
        add_grad(w, m->convs[o.conv].b_off, m->convs[o.conv].cout);
      }
      if (o.bn >= 0) {
        add_act(m, r, o.flag - 1);
        add_arena(w, m->bns[o.bn].bsums_off, 2 * (size_t)m->bns[o.bn].c * 4);
      }
      break;
    case B_RELU_MASK:
      add_act(m, r, o.a0); add_act(m, r, o.a1); add_act(m, w, o.a0);
      add_grad(w, m->convs[o.conv].b_off, m->convs[o.conv].cout);
      break;
    case B_COLSUM:
      add_act(m, r, o.a0); add_grad(w, m->convs[o.conv].b_off, m->convs[o.conv].cout);
      if (o.a1 >= 0) add_grad(w, m->convs[o.a1].b_off, m->convs[o.a1].cout);   // a1: a second convolution with the same gradient
      break;
    case B_POOL:
      add_act(m, r, o.a0); add_act(m, r, o.a1); if (o.flag) add_act(m, r, o.a2); add_act(m, w, o.a2);
      if (o.bn >= 0) { add_act(m, r, o.a3); add_arena(w, m->bns[o.bn].bsums_off, 2 * (size_t)m->bns[o.bn].c * 4); }   // fused BatchNorm-backward reduction
      break;
    case B_UPADD:
      add_act(m, r, o.a0); add_act(m, w, o.a1);
      if (o.bn >= 0) { add_act(m, r, o.a2); add_arena(w, m->bns[o.bn].bsums_off, 2 * (size_t)m->bns[o.bn].c * 4); }
      break;
    case B_HEAD:
      add_act(m, r, o.a0); add_act(m, w, o.a1);
      add_arena(r, m->heat_off[o.flag], hm_bytes);
      if (m->cfg.training) add_arena(r, m->dldp_off[o.flag], hm_bytes);
      break;
    case F_DW: add_act(m, r, o.a0); add_act(m, w, o.a1); break;
    case B_DW_DGRAD: add_act(m, r, o.a0); add_act(m, r, o.a2); add_act(m, r, o.a3); add_act(m, w, o.a1); break;
    case B_DW_WGRAD: {
      const DwL& d = m->dws[o.conv];
      add_act(m, r, o.a0); add_act(m, r, o.a1);
      add_grad(w, d.w_off, (int64_t)d.k * d.k * d.c);
      break;
    }
  }
}

inline bool overlaps(const std::vector<Range>& a, const std::vector<Range>& b) {
  for (const Range& x : a)
    for (const Range& y : b)
      if (x.space == y.space && x.lo < y.hi && y.lo < x.hi) return true;
  return false;
}

// For every op of the sequence: the latest earlier op of each OTHER lane it conflicts with (read-after-write,
// write-after-read, write-after-write on overlapping byte ranges), unless its own lane already waited for that op
// or a later one of the same lane.  Same-lane order is stream order.
void build_sequence(const hgb_model* m, const std::vector<std::vector<Op>>& lists, const std::vector<int>& seg_order,
                    std::vector<SchedOp>& seq, std::vector<int>* seg_begin) {
  seq.clear();
  if (seg_begin) seg_begin->assign(lists.size() + 1, 0);
  for (int seg : seg_order) {
    if (seg_begin) (*seg_begin)[seg] = (int)seq.size();
    for (int i = 0; i < (int)lists[seg].size(); ++i) { SchedOp so; so.seg = seg; so.idx = i; seq.push_back(so); }
  }
  const int n = (int)seq.size();
  std::vector<std::vector<Range>> R(n), W(n);
  std::vector<int> lane(n);
  for (int j = 0; j < n; ++j) {
    const Op& o = lists[seq[j].seg][seq[j].idx];
    op_access(m, o, R[j], W[j]);
    lane[j] = o.lane;
  }
  // synced[a][b]: lane a has (transitively through its own order) waited for every op of lane b up to this index
  std::vector<std::vector<int>> synced(kNumLanes, std::vector<int>(kNumLanes, -1));
  for (int j = 0; j < n; ++j) {
    const int lj = lane[j];
    bool found[kNumLanes] = {false};
    int nfound = 0;
    for (int i = j - 1; i >= 0 && nfound < kNumLanes - 1; --i) {
      const int li = lane[i];
      if (li == lj || found[li]) continue;
      if (i <= synced[lj][li]) { found[li] = true; ++nfound; continue; }   // everything older is already ordered
      if (overlaps(W[i], R[j]) || overlaps(W[i], W[j]) || overlaps(R[i], W[j])) {
        seq[j].deps.push_back({li, i});
        seq[i].signal = true;
        synced[lj][li] = i;
        // what lane li had waited for when op i was issued is ordered before j as well
        found[li] = true; ++nfound;
      }
    }
  }
}

void build_schedule(hgb_model* m) {
  std::vector<int> fo, bo;
  for (int s = 0; s <= m->S; ++s) fo.push_back(s);
  for (int s = m->S; s >= 0; --s) bo.push_back(s);
  build_sequence(m, m->fwd_ops, fo, m->fwd_seq, nullptr);
  build_sequence(m, m->bwd_ops, bo, m->bwd_seq, &m->bwd_seq_begin);
}

inline bf16* act_ptr(const hgb_model* m, int a) { return a < 0 ? nullptr : reinterpret_cast<bf16*>(m->p_arena + m->acts[a].off); }
inline float* arena_f(const hgb_model* m, size_t off) { return reinterpret_cast<float*>(m->p_arena + off); }

int build_maps(hgb_model* m) {
  for (auto& a : m->acts) {
    a.has_tmap = false;
    if (a.bn_of >= 0 || a.c % 64 != 0 || a.w > 128) continue;   // virtual (deferred-BN) tensors own no memory
    int rc = make_tmap_act(&a.tmap, m->p_arena + a.off, a.n, a.h, a.w, a.c);
    if (rc) return rc;
    a.has_tmap = true;
  }
  for (auto& c : m->convs) {
    int rc = make_tmap_mat(&c.tm_wf, m->p_arena + c.wf_off, c.cout_pad, c.taps * c.cin_pad, conv_gemm_block_n(c.cout_pad));
    if (rc) return rc;
    if (c.has_wd) {
      rc = make_tmap_mat(&c.tm_wd, m->p_arena + c.wd_off, c.cin_pad, c.taps * c.cout_pad, conv_gemm_block_n(c.cin_pad));
      if (rc) return rc;
    }
    c.has_w64 = c.ksize == 3 && c.cout_pad == 128 && c.cin_pad == 128 && c.has_wd;
    if (c.has_w64) {
      rc = make_tmap_mat(&c.tm_wf64, m->p_arena + c.wf_off, c.cout_pad, c.taps * c.cin_pad, 64);
      if (rc) return rc;
      rc = make_tmap_mat(&c.tm_wd64, m->p_arena + c.wd_off, c.cin_pad, c.taps * c.cout_pad, 64);
      if (rc) return rc;
    }
  }
  // weight-refresh table
  std::vector<WeightSyncEntry> tab(m->convs.size());
  for (size_t i = 0; i < m->convs.size(); ++i) {
    const ConvL& c = m->convs[i];
    tab[i].w = m->p_params + c.w_off;
    tab[i].wf = reinterpret_cast<bf16*>(m->p_arena + c.wf_off);
    tab[i].wd = c.has_wd ? reinterpret_cast<bf16*>(m->p_arena + c.wd_off) : nullptr;
    tab[i].taps = c.taps; tab[i].cin = c.cin; tab[i].cout = c.cout; tab[i].cin_pad = c.cin_pad; tab[i].cout_pad = c.cout_pad;
  }
  HGB_CUDA(cudaMemcpy(m->p_arena + m->sync_off, tab.data(), tab.size() * sizeof(WeightSyncEntry), cudaMemcpyHostToDevice));
  m->maps_ready = true;
  return HGB_OK;
}

int run_op_impl(hgb_model* m, const Op& o, const float* images, int training, cudaStream_t st);

int run_op(hgb_model* m, const Op& o, const float* images, int training, cudaStream_t st) {
  // programmatic dependent launch pays off when kernels are short (measured: -6.5 % step time at batch 32,
  // +2.5 % at batch 256), so it follows the plan's batch unless forced by hgb_debug_set(7, 1 = on / 2 = off)
  // With the lanes active, an early-launched dependent CTA parks on an SM (holding ~200 KB of shared memory) until its
  // predecessor finishes and keeps the OTHER lanes' kernels off that SM: measured -1.7 % in the backward pass, where the
  // side lanes carry a third of the work, so there it stays off.
  // Small ops are the exception to both rules: a kernel of at most one wave of 128-pixel tiles never fills the chip, so an
  // early-launched dependent parks on an SM nobody else wants, and what it hides (launch latency, mbarrier / TMEM / descriptor
  // set-up: about half of a 7-10 us kernel at the 16x16, 8x8 and 4x4 levels) is what those levels are made of.
  // hgb_debug_set(27, 1) restores the batch / pass rule for every op.
  bool small_op = false;
  if (!hgb::g_debug[27] && o.a0 >= 0) {
    const Act& t0 = m->acts[o.a0];
    small_op = (int64_t)t0.n * t0.h * t0.w <= (int64_t)128 * 148;
  }
  hgb::g_debug[6] = hgb::g_debug[7] == 1 ? 0 : hgb::g_debug[7] == 2 ? 1
                  : small_op ? 0 : (m->B > 64 || (m->pdl_suppressed && !hgb::g_debug[19]));
  bool timed = false;
  if (m->prof_all) {
    timed = m->prof_used + 2 <= m->prof_ev.size();
    if (timed) { cudaEventRecord(m->prof_ev[m->prof_used], st); m->prof_ops.push_back(o); }
  } else if (m->prof_on && (int)o.type == m->prof_type && m->prof_k > 0 && o.conv >= 0) {
    const ConvL& c = m->convs[o.conv];
    timed = c.real_k == m->prof_k && c.real_cin == m->prof_cin && c.cout == m->prof_cout && c.h == m->prof_h &&
            m->prof_used + 2 <= m->prof_ev.size();
    if (timed) { cudaEventRecord(m->prof_ev[m->prof_used], st); m->prof_flops += c.flops; }
  } else if (m->prof_on && (int)o.type == m->prof_type && m->prof_k == 0 && o.a0 >= 0) {
    // non-convolution classes (BatchNorm, pool, merge ...): selected by the channel count and height of their first tensor
    const Act& t = m->acts[o.a0];
    timed = t.c == m->prof_cout && t.h == m->prof_h && m->prof_used + 2 <= m->prof_ev.size();
    if (timed) cudaEventRecord(m->prof_ev[m->prof_used], st);
  }
  const int rc = run_op_impl(m, o, images, training, st);
  if (timed) { cudaEventRecord(m->prof_ev[m->prof_used + 1], st); m->prof_used += 2; }
  return rc;
}

// persistent GEMMs of the skip lanes leave a few SMs free so the main chain's small kernels never queue behind them
inline int side_lane_ctas(const hgb_model* m, const Op& o) {
  if (!is_skip_lane(o.lane) || !m->lanes_ready || hgb::g_debug[8] || m->prof_all) return 0;
  const int reserve = hgb::g_debug[9] > 0 ? hgb::g_debug[9] : 20;
  return m->num_sms > 2 * reserve ? m->num_sms - reserve : 0;
}

int run_op_impl(hgb_model* m, const Op& o, const float* images, int training, cudaStream_t st) {
  int rc = HGB_OK;
  switch (o.type) {
    case F_IM2COL: {
      rc = im2col_7x7s2(images, act_ptr(m, o.a0), m->B, m->cfg.in_h, m->cfg.in_w, st);
      break;
    }
    case F_CONV: {
      const ConvL& c = m->convs[o.conv];
      const Act& in = m->acts[o.a0];
      const Act& out = m->acts[o.a1];
      ConvGemmArgs a;
      a.N = in.n; a.H = in.h; a.W = in.w; a.Cin = c.cin_pad; a.Cout = c.cout_pad; a.ksize = c.ksize; a.tap_sign = 1;
      a.relu = c.relu; a.ldc = out.c;
      a.bias = m->p_params + c.b_off;
      a.res1 = act_ptr(m, o.a2); a.res2 = act_ptr(m, o.a3); a.out = act_ptr(m, o.a1);
      a.stats = (o.bn >= 0 && training) ? arena_f(m, m->bns[o.bn].sums_off) : nullptr;
      a.bn_y = nullptr;
      a.max_ctas = side_lane_ctas(m, o);
      if (o.flag & 0xffff) {   // deferred BatchNorm of the input, applied to the operand tile inside the kernel
        const BNL& b = m->bns[(o.flag & 0xffff) - 1];
        a.bn_in.sums = arena_f(m, b.sums_off); a.bn_in.saved = arena_f(m, b.saved_off);
        a.bn_in.gamma = m->p_params + b.gamma_off; a.bn_in.beta = m->p_params + b.beta_off;
        a.bn_in.moving_mean = m->p_params + b.mm_off; a.bn_in.moving_var = m->p_params + b.mv_off;
        a.bn_in.mode = training ? 0 : 1; a.bn_in.write = (o.flag & 0x10000) ? 1 : 0;
        a.bn_in.M = in.n * in.h * in.w * m->stat_ranks(); a.bn_in.C = b.c;
        if (training && (o.flag & 0x10000) && m->stat_ranks() > 1) {   // sync-BN: global batch statistics
          rc = comm_allreduce_sum_f32(m->comm, arena_f(m, b.sums_off), 2 * b.c, st);
          if (rc) break;
        }
      }
      if (o.flag & 0x20000) {   // inference: the BatchNorm that follows this conv is applied in the epilogue (fuse_inference_bn)
        const BNL& b = m->bns[o.bn];
        a.bn_out.gamma = m->p_params + b.gamma_off; a.bn_out.beta = m->p_params + b.beta_off;
        a.bn_out.moving_mean = m->p_params + b.mm_off; a.bn_out.moving_var = m->p_params + b.mv_off;
        a.bn_out.mode = 1; a.bn_out.M = in.n * in.h * in.w; a.bn_out.C = b.c;
      }
      rc = launch_conv_gemm(in.tmap, c.tm_wf, out.tmap, o.a2 >= 0 ? &m->acts[o.a2].tmap : nullptr, nullptr, a, st, nullptr, nullptr,
                            c.has_w64 ? &c.tm_wf64 : nullptr);
      break;
    }
    case F_BN: {
      const BNL& b = m->bns[o.bn];
      const Act& y = m->acts[o.a0];
      if (training && m->stat_ranks() > 1) {   // sync-BN: global batch statistics
        rc = comm_allreduce_sum_f32(m->comm, arena_f(m, b.sums_off), 2 * b.c, st);
        if (rc) break;
      }
      if (o.flag & 1) {     // + the MaxPool2D of the output
        rc = bn_apply_pool_fwd(act_ptr(m, o.a0), act_ptr(m, o.a1), act_ptr(m, o.a2), act_ptr(m, o.a3), arena_f(m, b.sums_off),
                               arena_f(m, b.saved_off), m->p_params + b.gamma_off, m->p_params + b.beta_off, m->p_params + b.mm_off,
                               m->p_params + b.mv_off, y.n * y.h * y.w, y.n * y.h * y.w * m->stat_ranks(), b.c, y.h, y.w, training, st);
        break;
      }
      rc = bn_apply_fwd(act_ptr(m, o.a0), act_ptr(m, o.a1), act_ptr(m, o.a2), arena_f(m, b.sums_off), arena_f(m, b.saved_off),
                        m->p_params + b.gamma_off, m->p_params + b.beta_off, m->p_params + b.mm_off, m->p_params + b.mv_off,
                        y.n * y.h * y.w, y.n * y.h * y.w * m->stat_ranks(), b.c, training, st, act_ptr(m, o.a3), y.h, y.w);
      break;
    }
    case F_POOL: {
      const Act& out = m->acts[o.a1];
      rc = maxpool_fwd(act_ptr(m, o.a0), act_ptr(m, o.a1), out.n, out.h, out.w, out.c, st);
      break;
    }
    case F_UPADD: {
      const Act& lo = m->acts[o.a1];
      rc = upsample_add_fwd(act_ptr(m, o.a0), act_ptr(m, o.a1), act_ptr(m, o.a2), lo.n, lo.h, lo.w, lo.c, st);
      break;
    }
    case F_HEAD: {
      const Act& l = m->acts[o.a0];
      rc = head_act_fwd(act_ptr(m, o.a0), l.c, arena_f(m, m->heat_off[o.flag]), act_ptr(m, o.a1), l.n * l.h * l.w, m->K,
                        m->cfg.activation, st);
      break;
    }
    case B_BN_REDUCE: {
      const BNL& b = m->bns[o.bn];
      const Act& y = m->acts[o.a1];
      rc = bn_bwd_reduce(act_ptr(m, o.a0), act_ptr(m, o.a1), arena_f(m, b.bsums_off), y.n * y.h * y.w, b.c, st);
      break;
    }
    case B_BN_APPLY: {
      const BNL& b = m->bns[o.bn];
      const ConvL& c = m->convs[o.conv];
      const Act& y = m->acts[o.a1];
      if (m->stat_ranks() > 1) {   // sync-BN: global sums of dz and dz*y; dgamma / dbeta then come out identical on every
        rc = comm_allreduce_sum_f32(m->comm, arena_f(m, b.bsums_off), 2 * b.c, st);   // rank and are pre-divided by the
        if (rc) break;                                                                // world size (the bucket all-reduce sums them)
      }
      rc = bn_bwd_apply(act_ptr(m, o.a0), act_ptr(m, o.a1), act_ptr(m, o.a2), arena_f(m, b.bsums_off), arena_f(m, b.saved_off),
                        m->p_params + b.gamma_off, m->p_grads + b.gamma_off, m->p_grads + b.beta_off, m->p_grads + c.b_off,
                        y.n * y.h * y.w, y.n * y.h * y.w * m->stat_ranks(), 1.f / (float)m->stat_ranks(), b.c, st);
      break;
    }
    case B_WGRAD: {
      const ConvL& c = m->convs[o.conv];
      const Act& dp = m->acts[o.a0];
      const Act& x = m->acts[o.a1];
      WgradArgs a;
      a.N = x.n; a.H = x.h; a.W = x.w; a.Cin = c.cin_pad; a.Cout = c.cout_pad; a.ksize = c.ksize;
      a.Cin_valid = c.cin; a.Cout_valid = c.cout;
      a.dw = m->p_grads + c.w_off;
      // On its side lane a weight gradient is never urgent, but its CTAs are long-lived and cannot be preempted: filling the
      // chip with them makes every main-chain kernel wait for an SM.  64 CTAs (of 148 SMs) measured best at batch 64 and above
      // (batch 256: -3.4 % step time); at batch 32 the weight-gradient lane itself is close to the critical path (cap 16: 36 ms,
      // 64: 25.5 ms, 96: 24.9 ms, none: 25.1 ms; profiles/r02_lane_caps.txt) and gets 96.  Dealing the leaf ops over two or
      // three streams (hgb_debug_set(35, n)) measured no better.  hgb_debug_set(20, n) overrides, -1 = no cap.
      if (m->lanes_ready && !hgb::g_debug[8] && !m->prof_all && o.lane != kLaneMain)
        a.max_ctas = hgb::g_debug[20] > 0 ? hgb::g_debug[20] : (hgb::g_debug[20] < 0 ? 0 : (m->B <= 48 ? 96 : 64));
      if (o.bn >= 0) {   // x = BatchNorm(o.bn)(source tensor), rebuilt from the saved statistics
        const BNL& b = m->bns[o.bn];
        a.bn_in.saved = arena_f(m, b.saved_off);
        a.bn_in.gamma = m->p_params + b.gamma_off; a.bn_in.beta = m->p_params + b.beta_off;
        a.bn_in.mode = 2; a.bn_in.M = x.n * x.h * x.w; a.bn_in.C = b.c;
      }
      rc = launch_conv_wgrad(dp.tmap, x.tmap, a, st);
      break;
    }
    case B_DGRAD: {
      const ConvL& c = m->convs[o.conv];
      const Act& dp = m->acts[o.a0];
      const Act& out = m->acts[o.a1];
      ConvGemmArgs a;
      a.N = dp.n; a.H = dp.h; a.W = dp.w; a.Cin = c.cout_pad; a.Cout = c.cin_pad; a.ksize = c.ksize; a.tap_sign = -1;
      a.relu = 0; a.ldc = out.c; a.bias = nullptr;
      a.res1 = act_ptr(m, o.a2); a.res2 = act_ptr(m, o.a3); a.out = act_ptr(m, o.a1);
      // fused BatchNorm-backward reduction of the BN that consumes this gradient (o.bn, y = act o.flag - 1)
      a.stats = o.bn >= 0 ? arena_f(m, m->bns[o.bn].bsums_off) : nullptr;
      a.bn_y = o.bn >= 0 ? act_ptr(m, o.flag - 1) : nullptr;
      a.max_ctas = side_lane_ctas(m, o);
      if (o.fbn >= 0) {   // BatchNorm backward fused into the operand prologue: the input map is dz, dp is stored on the side
        const BNL& b = m->bns[o.fbn];
        if (m->stat_ranks() > 1) {   // sync-BN: global sums of dz and dz*y (see B_BN_APPLY)
          rc = comm_allreduce_sum_f32(m->comm, arena_f(m, b.bsums_off), 2 * b.c, st);
          if (rc) break;
        }
        a.bn_bwd.bsums = arena_f(m, b.bsums_off); a.bn_bwd.saved = arena_f(m, b.saved_off);
        a.bn_bwd.gamma = m->p_params + b.gamma_off;
        a.bn_bwd.dgamma = m->p_grads + b.gamma_off; a.bn_bwd.dbeta = m->p_grads + b.beta_off; a.bn_bwd.dbias = m->p_grads + c.b_off;
        a.bn_bwd.M_stat = dp.n * dp.h * dp.w * m->stat_ranks(); a.bn_bwd.pscale = 1.f / (float)m->stat_ranks(); a.bn_bwd.C = b.c;
        rc = launch_conv_gemm(m->acts[o.fz].tmap, c.tm_wd, out.tmap, o.a2 >= 0 ? &m->acts[o.a2].tmap : nullptr,
                              o.bn >= 0 ? &m->acts[o.flag - 1].tmap : nullptr, a, st, &m->acts[o.fy].tmap, &dp.tmap);
        break;
      }
      rc = launch_conv_gemm(dp.tmap, c.tm_wd, out.tmap, o.a2 >= 0 ? &m->acts[o.a2].tmap : nullptr,
                            o.bn >= 0 ? &m->acts[o.flag - 1].tmap : nullptr, a, st, nullptr, nullptr, c.has_w64 ? &c.tm_wd64 : nullptr);
      break;
    }
    case B_RELU_MASK: {
      const ConvL& c = m->convs[o.conv];
      const Act& g = m->acts[o.a0];
      rc = relu_mask_colsum(act_ptr(m, o.a0), act_ptr(m, o.a1), act_ptr(m, o.a0), m->p_grads + c.b_off, g.n * g.h * g.w, g.c,
                            c.cout, 1, st);
      break;
    }
    case B_COLSUM: {
      const ConvL& c = m->convs[o.conv];
      const Act& g = m->acts[o.a0];
      rc = relu_mask_colsum(act_ptr(m, o.a0), nullptr, nullptr, m->p_grads + c.b_off, g.n * g.h * g.w, g.c, c.cout, 0, st,
                            o.a1 >= 0 ? m->p_grads + m->convs[o.a1].b_off : nullptr);
      break;
    }
    case B_POOL: {
      const Act& gy = m->acts[o.a1];
      rc = maxpool_bwd(act_ptr(m, o.a0), act_ptr(m, o.a1), act_ptr(m, o.a2), gy.n, gy.h, gy.w, gy.c, o.flag, st,
                       o.bn >= 0 ? act_ptr(m, o.a3) : nullptr, o.bn >= 0 ? arena_f(m, m->bns[o.bn].bsums_off) : nullptr);
      break;
    }
    case B_UPADD: {
      const Act& lo = m->acts[o.a1];
      rc = upsample_add_bwd(act_ptr(m, o.a0), act_ptr(m, o.a1), lo.n, lo.h, lo.w, lo.c, st,
                            o.bn >= 0 ? act_ptr(m, o.a2) : nullptr, o.bn >= 0 ? arena_f(m, m->bns[o.bn].bsums_off) : nullptr);
      break;
    }
    case B_HEAD: {
      const Act& gl = m->acts[o.a1];
      rc = head_act_bwd(arena_f(m, m->dldp_off[o.flag]), act_ptr(m, o.a0), arena_f(m, m->heat_off[o.flag]), act_ptr(m, o.a1),
                        gl.n * gl.h * gl.w, m->K, m->cfg.activation, st);
      break;
    }
    case F_DW: {
      const DwL& d = m->dws[o.conv];
      const Act& x = m->acts[o.a0];
      rc = dwconv(act_ptr(m, o.a0), m->p_params + d.w_off, nullptr, nullptr, act_ptr(m, o.a1), x.n, x.h, x.w, x.c, d.k, 0, st);
      break;
    }
    case B_DW_DGRAD: {
      const DwL& d = m->dws[o.conv];
      const Act& g = m->acts[o.a0];
      rc = dwconv(act_ptr(m, o.a0), m->p_params + d.w_off, act_ptr(m, o.a2), act_ptr(m, o.a3), act_ptr(m, o.a1), g.n, g.h, g.w, g.c, d.k, 1, st);
      break;
    }
    case B_DW_WGRAD: {
      const DwL& d = m->dws[o.conv];
      const Act& g = m->acts[o.a0];
      rc = dwconv_wgrad(act_ptr(m, o.a1), act_ptr(m, o.a0), m->p_grads + d.w_off, g.n, g.h, g.w, g.c, d.k, st);
      break;
    }
  }
  if (rc == HGB_OK) ++m->launches;
  return rc;
}


// ---- lane runtime.  The caller's stream only forks and joins; every op runs on a library-owned stream (the
// main chain at the highest priority so that its small kernels are dispatched ahead of queued side-lane blocks).
int ensure_lanes(hgb_model* m) {
  if (m->lanes_ready) return HGB_OK;
  int lo = 0, hi = 0;
  HGB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = numerically smallest = highest priority
  for (int l = 0; l < kNumLanes; ++l) {
    int pr = l == kLaneMain ? hi : hi + 1;
    if (is_wgrad_lane(l) && hgb::g_debug[21]) pr = hi + 1 + hgb::g_debug[21];   // experiment: weight gradients below the skip lanes
    if (pr > lo) pr = lo;
    HGB_CUDA(cudaStreamCreateWithPriority(&m->lane_stream[l], cudaStreamNonBlocking, pr));
    HGB_CUDA(cudaEventCreateWithFlags(&m->join_ev[l], cudaEventDisableTiming));
  }
  HGB_CUDA(cudaEventCreateWithFlags(&m->fork_ev, cudaEventDisableTiming));
  m->fwd_ev.assign(m->fwd_seq.size(), nullptr);
  m->bwd_ev.assign(m->bwd_seq.size(), nullptr);
  for (size_t i = 0; i < m->fwd_seq.size(); ++i)
    if (m->fwd_seq[i].signal) HGB_CUDA(cudaEventCreateWithFlags(&m->fwd_ev[i], cudaEventDisableTiming));
  for (size_t i = 0; i < m->bwd_seq.size(); ++i)
    if (m->bwd_seq[i].signal) HGB_CUDA(cudaEventCreateWithFlags(&m->bwd_ev[i], cudaEventDisableTiming));
  int dev = 0;
  HGB_CUDA(cudaGetDevice(&dev));
  HGB_CUDA(cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, dev));
  m->lanes_ready = true;
  return HGB_OK;
}

// Inference: BatchNorm is a per-channel affine map with constant statistics, so a 1x1 convolution followed by its
// (stored) BatchNorm runs as ONE launch: the epilogue computes BN(relu(conv + bias)) (+ the BN's residual) from the fp32
// accumulator and writes the BN's output tensor; the pre-BN tensor is never stored and the BN pass disappears.
bool fuse_inference_bn(const hgb_model* m, const Op& conv, const Op* next, int training, Op* fused) {
  if (training || hgb::g_debug[17] || !next || conv.type != F_CONV || next->type != F_BN || conv.bn < 0 || next->bn != conv.bn) return false;
  if (m->convs[conv.conv].ksize != 1 || conv.a2 >= 0 || conv.a3 >= 0 || next->a0 != conv.a1 || conv.lane != next->lane) return false;
  if (next->a3 >= 0) return false;     // the BatchNorm also adds the upsampled lower level: a job for its own kernel
  *fused = conv;
  fused->a1 = next->a2;      // write the BN's output tensor
  fused->a2 = next->a1;      // + the BN's residual (identity / projected skip)
  fused->flag |= 0x20000;
  return true;
}

// Ops [begin, end) of a sequence.  Single-lane mode (hgb_debug_set(8, 1), or per-op profiling) replays them in
// order on the caller's stream -- the reference behaviour the lanes must reproduce.
int join_lanes(hgb_model* m, cudaStream_t st, bool caller) {
  for (int l = 0; l < kNumLanes; ++l)
    if (m->lane_dirty[l]) {
      HGB_CUDA(cudaEventRecord(m->join_ev[l], m->lane_stream[l]));
      HGB_CUDA(cudaStreamWaitEvent(st, m->join_ev[l], 0));
      if (caller) m->lane_dirty[l] = false;
    }
  if (caller) m->bwd_unjoined_from = 1 << 30;
  return HGB_OK;
}

// join = false (backward only): the caller's stream is NOT ordered after the lanes when the call returns; a later
// hgb_model_lanes_join does that.  Data parallelism issues one segment per call and must not stall the main chain at
// every segment boundary until that segment's (capped, low-priority) weight gradients have drained.
int run_sequence(hgb_model* m, bool backward, int begin, int end, const float* images, int training, cudaStream_t st,
                 bool join = true) {
  const std::vector<SchedOp>& seq = backward ? m->bwd_seq : m->fwd_seq;
  const std::vector<std::vector<Op>>& lists = backward ? m->bwd_ops : m->fwd_ops;
  if (begin >= end) return HGB_OK;
  auto op_at = [&](int k) -> const Op& { return lists[seq[k].seg][seq[k].idx]; };
  if (hgb::g_debug[8] || m->prof_all || m->stat_ranks() > 1) {   // (sync-BN interleaves collectives: one stream, plan order)
    for (int k = begin; k < end; ++k) {
      Op fused;
      const bool fz = !backward && fuse_inference_bn(m, op_at(k), k + 1 < end ? &op_at(k + 1) : nullptr, training, &fused);
      int rc = run_op(m, fz ? fused : op_at(k), images, training, st);
      if (rc) return rc;
      if (fz) ++k;
    }
    return HGB_OK;
  }
  int rc = ensure_lanes(m);
  if (rc) return rc;
  std::vector<cudaEvent_t>& ev = backward ? m->bwd_ev : m->fwd_ev;
  bool used[kNumLanes] = {false};
  m->pdl_suppressed = backward;
  HGB_CUDA(cudaEventRecord(m->fork_ev, st));
  for (int k = begin; k < end; ++k) {
    Op fused;
    const bool fz = !backward && fuse_inference_bn(m, op_at(k), k + 1 < end ? &op_at(k + 1) : nullptr, training, &fused);
    const Op& o = fz ? fused : op_at(k);
    cudaStream_t ls = m->lane_stream[o.lane];
    if (!used[o.lane]) { HGB_CUDA(cudaStreamWaitEvent(ls, m->fork_ev, 0)); used[o.lane] = true; }
    for (int q = k; q <= k + (fz ? 1 : 0); ++q)    // a fused launch inherits the dependencies of both ops
      for (const Dep& d : seq[q].deps)   // ops of earlier calls need no wait once the caller's stream has joined them
        if (d.idx >= begin || (backward && d.idx >= m->bwd_unjoined_from)) HGB_CUDA(cudaStreamWaitEvent(ls, ev[d.idx], 0));
    rc = run_op(m, o, images, training, ls);
    if (rc) return rc;
    if (seq[k].signal) HGB_CUDA(cudaEventRecord(ev[k], ls));
    if (fz) {
      ++k;
      if (seq[k].signal) HGB_CUDA(cudaEventRecord(ev[k], ls));
    }
  }
  m->pdl_suppressed = false;
  for (int l = 0; l < kNumLanes; ++l) m->lane_dirty[l] = m->lane_dirty[l] || used[l];
  if (backward && begin < m->bwd_unjoined_from) m->bwd_unjoined_from = begin;
  if (join) return join_lanes(m, st, true);
  return HGB_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------- C ABI
extern "C" int hgb_model_create(const hgb_model_config* cfg, int device, hgb_model** out) {
  HGB_CHECK_ARG(cfg && out, "hgb_model_create: null pointer");
  HGB_CHECK_ARG(cfg->num_stacks >= 1 && cfg->num_stacks <= 64, "hgb_model_create: num_stacks out of range");
  HGB_CHECK_ARG(cfg->num_channels >= 128 && cfg->num_channels % 128 == 0 && cfg->num_channels <= 256,
                "hgb_model_create: num_channels must be 128 or 256 (got %d)", cfg->num_channels);
  HGB_CHECK_ARG(cfg->num_classes >= 1 && cfg->num_classes <= 64, "hgb_model_create: num_classes must be in [1,64]");
  HGB_CHECK_ARG(cfg->in_h == cfg->in_w && cfg->in_h >= 64 && cfg->in_h <= 256 && (cfg->in_h & (cfg->in_h - 1)) == 0,
                "hgb_model_create: input must be square, a power of two in [64,256] (got %dx%d)", cfg->in_h, cfg->in_w);
  HGB_CHECK_ARG(cfg->activation == 0 || cfg->activation == 1, "hgb_model_create: activation must be 0 (linear) or 1 (sigmoid)");
  HGB_CHECK_ARG(cfg->batch >= 1, "hgb_model_create: batch must be >= 1");
  HGB_CHECK_ARG(cfg->mobile == 0 || cfg->mobile == 1, "hgb_model_create: mobile must be 0 or 1 (got %d)", cfg->mobile);
  HGB_CHECK_ARG(((int64_t)cfg->in_h / 4) * (cfg->in_w / 4) * cfg->num_classes % 4 == 0, "hgb_model_create: heat map size");
  hgb_model* m = new hgb_model();
  m->cfg = *cfg;
  m->device = device;
  m->S = cfg->num_stacks; m->C = cfg->num_channels; m->K = cfg->num_classes; m->B = cfg->batch;
  int rc = build(m);
  if (rc) { delete m; return rc; }
  build_schedule(m);
  *out = m;
  return HGB_OK;
}

extern "C" int hgb_model_destroy(hgb_model* m) {
  if (m) {
    for (cudaEvent_t e : m->prof_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : m->fwd_ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : m->bwd_ev) if (e) cudaEventDestroy(e);
    if (m->fork_ev) cudaEventDestroy(m->fork_ev);
    for (int l = 0; l < kNumLanes; ++l) {
      if (m->join_ev[l]) cudaEventDestroy(m->join_ev[l]);
      if (m->lane_stream[l]) cudaStreamDestroy(m->lane_stream[l]);
    }
  }
  delete m;
  return HGB_OK;
}

extern "C" int64_t hgb_model_param_count(const hgb_model* m, int trainable_only) {
  return trainable_only ? m->n_train : m->n_train + m->n_nontrain;
}

extern "C" int64_t hgb_model_buffer_bytes(const hgb_model* m, int which) {
  switch (which) {
    case HGB_BUF_PARAMS: return (m->train_floats + m->nontrain_floats + 256) * 4;
    case HGB_BUF_GRADS:
    case HGB_BUF_ADAM_M:
    case HGB_BUF_ADAM_V: return (m->train_floats + 256) * 4;
    case HGB_BUF_ARENA: return (int64_t)m->arena_bytes;
  }
  set_error("hgb_model_buffer_bytes: unknown buffer %d", which);
  return HGB_ERR_INVALID;
}

extern "C" int hgb_model_bind(hgb_model* m, int which, void* ptr, int64_t bytes) {
  HGB_CHECK_ARG(m && ptr, "hgb_model_bind: null pointer");
  HGB_CHECK_ARG(bytes >= hgb_model_buffer_bytes(m, which), "hgb_model_bind: buffer %d too small (%lld < %lld)", which,
                (long long)bytes, (long long)hgb_model_buffer_bytes(m, which));
  HGB_CHECK_ARG(((uintptr_t)ptr & 255) == 0, "hgb_model_bind: buffers must be 256-byte aligned");
  switch (which) {
    case HGB_BUF_PARAMS: m->p_params = (float*)ptr; break;
    case HGB_BUF_GRADS: m->p_grads = (float*)ptr; break;
    case HGB_BUF_ADAM_M: m->p_m = (float*)ptr; break;
    case HGB_BUF_ADAM_V: m->p_v = (float*)ptr; break;
    case HGB_BUF_ARENA: m->p_arena = (uint8_t*)ptr; break;
    default: set_error("hgb_model_bind: unknown buffer %d", which); return HGB_ERR_INVALID;
  }
  m->maps_ready = false;
  if (m->p_params && m->p_arena) return build_maps(m);
  return HGB_OK;
}

extern "C" int hgb_model_num_tensors(const hgb_model* m) { return (int)m->params.size(); }

extern "C" int hgb_model_tensor_info(const hgb_model* m, int index, const char** name, int* rank, int64_t dims[4], int64_t* offset,
                                     int* trainable) {
  HGB_CHECK_ARG(index >= 0 && index < (int)m->params.size(), "hgb_model_tensor_info: index out of range");
  const ParamT& p = m->params[index];
  if (name) *name = p.name.c_str();
  if (rank) *rank = p.rank;
  if (dims) for (int i = 0; i < 4; ++i) dims[i] = p.dims[i];
  if (offset) *offset = p.off;
  if (trainable) *trainable = p.trainable;
  return HGB_OK;
}

extern "C" int hgb_model_num_convs(const hgb_model* m) { return (int)m->convs.size(); }

extern "C" int hgb_model_conv_info(const hgb_model* m, int index, const char** name, int* k, int* cin, int* cout, int* h, int* w,
                                   double* flops) {
  HGB_CHECK_ARG(index >= 0 && index < (int)m->convs.size(), "hgb_model_conv_info: index out of range");
  const ConvL& c = m->convs[index];
  if (name) *name = c.name.c_str();
  if (k) *k = c.real_k;
  if (cin) *cin = c.real_cin;
  if (cout) *cout = c.cout;
  if (h) *h = c.h;
  if (w) *w = c.w;
  if (flops) *flops = c.flops;
  return HGB_OK;
}

#define HGB_REQUIRE_READY(m)                                                                   \
  do {                                                                                         \
    if (!(m) || !(m)->maps_ready) {                                                            \
      set_error("model buffers are not bound (bind HGB_BUF_PARAMS and HGB_BUF_ARENA first)");  \
      return HGB_ERR_STATE;                                                                    \
    }                                                                                          \
  } while (0)

extern "C" int hgb_model_sync_weights(hgb_model* m, void* stream) {
  HGB_REQUIRE_READY(m);
  int rc = weight_sync(reinterpret_cast<const WeightSyncEntry*>(m->p_arena + m->sync_off), (int)m->convs.size(),
                       m->sync_max_elems, (cudaStream_t)stream);
  if (rc == HGB_OK) ++m->launches;
  return rc;
}

extern "C" int hgb_model_forward(hgb_model* m, const float* images, int training, float* const* heatmaps_out, void* stream) {
  HGB_REQUIRE_READY(m);
  HGB_CHECK_ARG(images, "hgb_model_forward: null images");
  cudaStream_t st = (cudaStream_t)stream;
  if (training) {
    if (!m->cfg.training || !m->p_grads) { set_error("hgb_model_forward: model was not created/bound for training"); return HGB_ERR_STATE; }
    HGB_CUDA(cudaMemsetAsync(m->p_arena + m->zero_off, 0, m->zero_bytes, st));
    HGB_CUDA(cudaMemsetAsync(m->p_grads, 0, (size_t)m->train_floats * 4, st));
  }
  {
    int rc = run_sequence(m, false, 0, (int)m->fwd_seq.size(), images, training, st);
    if (rc) return rc;
  }
  if (heatmaps_out) {
    const size_t bytes = (size_t)m->B * m->hm_h * m->hm_w * m->K * 4;
    for (int s = 0; s < m->S; ++s)
      if (heatmaps_out[s]) HGB_CUDA(cudaMemcpyAsync(heatmaps_out[s], m->p_arena + m->heat_off[s], bytes, cudaMemcpyDeviceToDevice, st));
  }
  m->fwd_training_done = training != 0;
  return HGB_OK;
}

extern "C" int hgb_model_loss(hgb_model* m, int kind, const float* y_true, double inv_count, double* loss_acc, void* stream) {
  HGB_REQUIRE_READY(m);
  HGB_CHECK_ARG(y_true && loss_acc, "hgb_model_loss: null pointer");
  if (!m->cfg.training) { set_error("hgb_model_loss: model was not created for training"); return HGB_ERR_STATE; }
  for (int s = 0; s < m->S; ++s) {
    int rc = hgb_loss_fwd_bwd(kind, y_true, arena_f(m, m->heat_off[s]), HGB_F32, m->B, m->hm_h, m->hm_w, m->K, inv_count,
                              loss_acc + s, arena_f(m, m->dldp_off[s]), HGB_F32, m->p_arena + m->lossws_off, stream);
    if (rc) return rc;
    m->launches += kind >= 2 ? 2 : 1;
  }
  return HGB_OK;
}

extern "C" int hgb_model_num_segments(const hgb_model* m) { return m->S + 1; }

// zero the per-step accumulators (BN sums, gradients); hgb_model_forward(training=1) does this itself
extern "C" int hgb_model_begin_step(hgb_model* m, void* stream) {
  HGB_REQUIRE_READY(m);
  if (!m->cfg.training || !m->p_grads) { set_error("hgb_model_begin_step: model was not created/bound for training"); return HGB_ERR_STATE; }
  HGB_CUDA(cudaMemsetAsync(m->p_arena + m->zero_off, 0, m->zero_bytes, (cudaStream_t)stream));
  HGB_CUDA(cudaMemsetAsync(m->p_grads, 0, (size_t)m->train_floats * 4, (cudaStream_t)stream));
  m->fwd_training_done = true;
  return HGB_OK;
}

extern "C" int hgb_model_backward(hgb_model* m, int seg_lo, int seg_hi, void* stream) {
  HGB_REQUIRE_READY(m);
  HGB_CHECK_ARG(seg_lo >= 0 && seg_hi <= m->S + 1 && seg_lo < seg_hi, "hgb_model_backward: bad segment range");
  if (!m->cfg.training || !m->fwd_training_done) { set_error("hgb_model_backward: needs a training forward pass first"); return HGB_ERR_STATE; }
  // the backward sequence runs segments S..0: [seg_lo, seg_hi) is the contiguous range that starts with seg_hi - 1
  const int begin = m->bwd_seq_begin[seg_hi - 1];
  const int end = seg_lo == 0 ? (int)m->bwd_seq.size() : m->bwd_seq_begin[seg_lo - 1];
  return run_sequence(m, true, begin, end, nullptr, 1, (cudaStream_t)stream);
}

extern "C" int hgb_model_backward_nojoin(hgb_model* m, int seg_lo, int seg_hi, void* stream) {
  HGB_REQUIRE_READY(m);
  HGB_CHECK_ARG(seg_lo >= 0 && seg_hi <= m->S + 1 && seg_lo < seg_hi, "hgb_model_backward_nojoin: bad segment range");
  if (!m->cfg.training || !m->fwd_training_done) { set_error("hgb_model_backward_nojoin: needs a training forward pass first"); return HGB_ERR_STATE; }
  const int begin = m->bwd_seq_begin[seg_hi - 1];
  const int end = seg_lo == 0 ? (int)m->bwd_seq.size() : m->bwd_seq_begin[seg_lo - 1];
  return run_sequence(m, true, begin, end, nullptr, 1, (cudaStream_t)stream, /*join=*/hgb::g_debug[8] || m->prof_all);
}

// ---- data parallelism behind the ABI
extern "C" int hgb_model_set_comm(hgb_model* m, hgb_comm* comm, int sync_bn) {
  HGB_CHECK_ARG(m, "hgb_model_set_comm: null model");
  m->comm = comm;
  m->sync_bn = (comm && sync_bn) ? 1 : 0;
  return HGB_OK;
}

extern "C" int hgb_grad_allreduce_bucket(hgb_model* m, int seg_lo, int seg_hi, void* stream) {
  HGB_REQUIRE_READY(m);
  HGB_CHECK_ARG(seg_lo >= 0 && seg_hi <= m->S + 1 && seg_lo < seg_hi, "hgb_grad_allreduce_bucket: bad segment range");
  if (!m->comm || !m->p_grads) { set_error("hgb_grad_allreduce_bucket: no communicator (hgb_model_set_comm) or gradients bound"); return HGB_ERR_STATE; }
  // segments own consecutive ranges of the flat gradient buffer: [seg_lo, seg_hi) is ONE contiguous bucket
  const int64_t lo = m->seg_begin[seg_lo], hi = m->seg_end[seg_hi - 1];
  return comm_allreduce_sum_f32(m->comm, m->p_grads + lo, hi - lo, (cudaStream_t)stream);
}

// order `stream` after everything issued on the lanes so far; is_caller != 0 marks the lanes as joined (the stream the
// next forward / backward / optimizer call is issued on), 0 is for a side stream (e.g. the gradient all-reduce)
extern "C" int hgb_model_lanes_join(hgb_model* m, void* stream, int is_caller) {
  HGB_REQUIRE_READY(m);
  if (!m->lanes_ready) return HGB_OK;
  return join_lanes(m, (cudaStream_t)stream, is_caller != 0);
}

extern "C" int hgb_model_segment_grads(const hgb_model* m, int seg, int64_t* offset, int64_t* count) {
  HGB_CHECK_ARG(seg >= 0 && seg <= m->S, "hgb_model_segment_grads: bad segment");
  if (offset) *offset = m->seg_begin[seg];
  if (count) *count = m->seg_end[seg] - m->seg_begin[seg];
  return HGB_OK;
}

extern "C" int hgb_model_adam_step(hgb_model* m, double lr, double beta1, double beta2, double eps, int64_t t, double grad_scale,
                                   void* stream) {
  HGB_REQUIRE_READY(m);
  HGB_CHECK_ARG(t >= 1, "hgb_model_adam_step: t starts at 1");
  if (!m->p_grads || !m->p_m || !m->p_v) { set_error("hgb_model_adam_step: optimizer buffers are not bound"); return HGB_ERR_STATE; }
  const double lr_t = lr * std::sqrt(1.0 - std::pow(beta2, (double)t)) / (1.0 - std::pow(beta1, (double)t));
  hgb::g_debug[6] = hgb::g_debug[7] == 1 ? 0 : hgb::g_debug[7] == 2 ? 1 : (m->B > 64);
  int rc = adam_step(m->p_params, m->p_grads, m->p_m, m->p_v, m->train_floats, (float)lr_t, (float)beta1, (float)beta2, (float)eps,
                     (float)grad_scale, (cudaStream_t)stream);
  if (rc) return rc;
  ++m->launches;
  return hgb_model_sync_weights(m, stream);
}

extern "C" int hgb_model_conv_output(const hgb_model* m, int index, int64_t* arena_offset, int dims[4]) {
  HGB_CHECK_ARG(index >= 0 && index < (int)m->convs.size(), "hgb_model_conv_output: index out of range");
  const int which = hgb::g_debug[2];  // 0: conv output, 1: conv input, 2: BN output (if any)
  const ConvL& c = m->convs[index];
  const int ai = which == 1 ? c.in_act : (which == 2 ? c.z_act : c.y_act);
  HGB_CHECK_ARG(ai >= 0, "hgb_model_conv_output: conv %d has no such tensor", index);
  const Act& a = m->acts[ai];
  if (arena_offset) *arena_offset = (int64_t)a.off;
  if (dims) { dims[0] = a.n; dims[1] = a.h; dims[2] = a.w; dims[3] = a.c; }
  return HGB_OK;
}

// deferred BatchNorm: index of the BN the conv applies to its INPUT tile inside the kernel (its input tensor, as
// reported by hgb_model_conv_output / op info, is then the PRE-BN tensor), or -1
extern "C" int hgb_model_conv_input_bn(const hgb_model* m, int conv) {
  if (conv < 0 || conv >= (int)m->convs.size()) return -1;
  return m->convs[conv].in_bn;
}

// ---- plan introspection / single-op stepping (tests replay every op against fp32 torch)
extern "C" int hgb_model_num_ops(const hgb_model* m, int seg, int backward) {
  if (seg < 0 || seg > m->S) return 0;
  return (int)(backward ? m->bwd_ops[seg].size() : m->fwd_ops[seg].size());
}
extern "C" int hgb_model_op_info(const hgb_model* m, int seg, int backward, int index, int info[8]) {
  HGB_CHECK_ARG(seg >= 0 && seg <= m->S, "hgb_model_op_info: bad segment");
  const std::vector<Op>& v = backward ? m->bwd_ops[seg] : m->fwd_ops[seg];
  HGB_CHECK_ARG(index >= 0 && index < (int)v.size(), "hgb_model_op_info: index out of range");
  const Op& o = v[index];
  info[0] = (int)o.type; info[1] = o.conv; info[2] = o.bn; info[3] = o.a0; info[4] = o.a1; info[5] = o.a2; info[6] = o.a3;
  info[7] = o.flag;
  return HGB_OK;
}
// B_DGRAD ops with the BatchNorm backward fused in: out = {BatchNorm index, dz act, y act} (the dp tensor is the op's a0), or
// {-1, -1, -1} for every other op
extern "C" int hgb_model_op_fused_bn(const hgb_model* m, int seg, int backward, int index, int out[3]) {
  HGB_CHECK_ARG(seg >= 0 && seg <= m->S, "hgb_model_op_fused_bn: bad segment");
  const std::vector<Op>& v = backward ? m->bwd_ops[seg] : m->fwd_ops[seg];
  HGB_CHECK_ARG(index >= 0 && index < (int)v.size(), "hgb_model_op_fused_bn: index out of range");
  out[0] = v[index].fbn; out[1] = v[index].fz; out[2] = v[index].fy;
  return HGB_OK;
}
extern "C" int hgb_model_act_info(const hgb_model* m, int act, int64_t* arena_offset, int dims[4]) {
  HGB_CHECK_ARG(act >= 0 && act < (int)m->acts.size(), "hgb_model_act_info: index out of range");
  const Act& a = m->acts[act];
  if (arena_offset) *arena_offset = (int64_t)a.off;
  if (dims) { dims[0] = a.n; dims[1] = a.h; dims[2] = a.w; dims[3] = a.c; }
  return HGB_OK;
}
extern "C" int hgb_model_run_op(hgb_model* m, int seg, int backward, int index, const float* images, int training, void* stream) {
  HGB_REQUIRE_READY(m);
  HGB_CHECK_ARG(seg >= 0 && seg <= m->S, "hgb_model_run_op: bad segment");
  const std::vector<Op>& v = backward ? m->bwd_ops[seg] : m->fwd_ops[seg];
  HGB_CHECK_ARG(index >= 0 && index < (int)v.size(), "hgb_model_run_op: index out of range");
  return run_op(m, v[index], images, training, (cudaStream_t)stream);
}
// info: ksize, taps, cin, cout, cin_pad, cout_pad, relu, has_dgrad; offs: kernel, bias (floats into params/grads)
extern "C" int hgb_model_conv_detail(const hgb_model* m, int conv, int info[8], int64_t offs[2]) {
  HGB_CHECK_ARG(conv >= 0 && conv < (int)m->convs.size(), "hgb_model_conv_detail: index out of range");
  const ConvL& c = m->convs[conv];
  info[0] = c.ksize; info[1] = c.taps; info[2] = c.cin; info[3] = c.cout; info[4] = c.cin_pad; info[5] = c.cout_pad;
  info[6] = c.relu; info[7] = c.has_wd;
  offs[0] = c.w_off; offs[1] = c.b_off;
  return HGB_OK;
}
// depthwise stage `dw` of the mobile variant: info = {k, channels, h, w}; *w_off = float offset of its [k*k][c] weights
extern "C" int hgb_model_dw_detail(const hgb_model* m, int dw, int info[4], int64_t* w_off) {
  HGB_CHECK_ARG(dw >= 0 && dw < (int)m->dws.size(), "hgb_model_dw_detail: index out of range");
  const DwL& d = m->dws[dw];
  info[0] = d.k; info[1] = d.c; info[2] = d.h; info[3] = d.w;
  if (w_off) *w_off = d.w_off;
  return HGB_OK;
}
// offs: channels, gamma, beta, moving_mean, moving_var (floats into params), sums, bsums, saved (bytes into arena)
extern "C" int hgb_model_bn_detail(const hgb_model* m, int bn, int64_t offs[8]) {
  HGB_CHECK_ARG(bn >= 0 && bn < (int)m->bns.size(), "hgb_model_bn_detail: index out of range");
  const BNL& b = m->bns[bn];
  offs[0] = b.c; offs[1] = b.gamma_off; offs[2] = b.beta_off; offs[3] = b.mm_off; offs[4] = b.mv_off;
  offs[5] = (int64_t)b.sums_off; offs[6] = (int64_t)b.bsums_off; offs[7] = (int64_t)b.saved_off;
  return HGB_OK;
}
// byte offsets into the arena of stack s's fp32 heat map and loss gradient
extern "C" int hgb_model_head_buffers(const hgb_model* m, int stack, int64_t offs[2]) {
  HGB_CHECK_ARG(stack >= 0 && stack < m->S, "hgb_model_head_buffers: bad stack");
  offs[0] = (int64_t)m->heat_off[stack];
  offs[1] = m->cfg.training ? (int64_t)m->dldp_off[stack] : -1;
  return HGB_OK;
}

// ---- lane schedule introspection (tests/test_cpu_host.py proves every conflicting pair is ordered)
extern "C" int hgb_model_sched_count(const hgb_model* m, int backward) {
  return (int)(backward ? m->bwd_seq.size() : m->fwd_seq.size());
}
// info: segment, index in the segment's op list, lane, signals; deps: (lane, sequence index) pairs
extern "C" int hgb_model_sched_op(const hgb_model* m, int backward, int k, int info[4], int* ndeps, int deps[16]) {
  const std::vector<SchedOp>& seq = backward ? m->bwd_seq : m->fwd_seq;
  HGB_CHECK_ARG(k >= 0 && k < (int)seq.size(), "hgb_model_sched_op: index out of range");
  const std::vector<std::vector<Op>>& lists = backward ? m->bwd_ops : m->fwd_ops;
  info[0] = seq[k].seg; info[1] = seq[k].idx; info[2] = lists[seq[k].seg][seq[k].idx].lane; info[3] = seq[k].signal;
  const int n = (int)seq[k].deps.size();
  HGB_CHECK_ARG(n <= 8, "hgb_model_sched_op: too many dependencies");
  *ndeps = n;
  for (int i = 0; i < n; ++i) { deps[2 * i] = seq[k].deps[i].lane; deps[2 * i + 1] = seq[k].deps[i].idx; }
  return HGB_OK;
}
// ranges: up to `cap` rows of (space, lo, hi, is_write); returns the number of rows (or < 0)
extern "C" int hgb_model_sched_access(const hgb_model* m, int backward, int k, int cap, int64_t* ranges) {
  const std::vector<SchedOp>& seq = backward ? m->bwd_seq : m->fwd_seq;
  HGB_CHECK_ARG(k >= 0 && k < (int)seq.size(), "hgb_model_sched_access: index out of range");
  const std::vector<std::vector<Op>>& lists = backward ? m->bwd_ops : m->fwd_ops;
  std::vector<Range> r, w;
  op_access(m, lists[seq[k].seg][seq[k].idx], r, w);
  HGB_CHECK_ARG((int)(r.size() + w.size()) <= cap, "hgb_model_sched_access: capacity");
  int n = 0;
  for (const Range& x : r) { ranges[4 * n] = x.space; ranges[4 * n + 1] = (int64_t)x.lo; ranges[4 * n + 2] = (int64_t)x.hi; ranges[4 * n + 3] = 0; ++n; }
  for (const Range& x : w) { ranges[4 * n] = x.space; ranges[4 * n + 1] = (int64_t)x.lo; ranges[4 * n + 2] = (int64_t)x.hi; ranges[4 * n + 3] = 1; ++n; }
  return n;
}

// time every launch of one conv class with CUDA event pairs on the launching stream.
// op_type: 1 forward conv, 8 wgrad, 9 dgrad (OpType).  enable=0 stops; read after a stream sync.
extern "C" int hgb_model_profile_conv(hgb_model* m, int enable, int op_type, int k, int cin, int cout, int h) {
  HGB_CHECK_ARG(m, "hgb_model_profile_conv: null model");
  m->prof_on = enable; m->prof_type = op_type; m->prof_k = k; m->prof_cin = cin; m->prof_cout = cout; m->prof_h = h;
  if (enable) {
    m->prof_used = 0; m->prof_flops = 0;
    while (m->prof_ev.size() < 16384) {
      cudaEvent_t e;
      HGB_CUDA(cudaEventCreate(&e));
      m->prof_ev.push_back(e);
    }
  }
  return HGB_OK;
}
// time EVERY op of the following steps (event pair per op); dump with hgb_model_profile_op after a sync
extern "C" int hgb_model_profile_all(hgb_model* m, int enable) {
  HGB_CHECK_ARG(m, "hgb_model_profile_all: null model");
  m->prof_all = enable; m->prof_on = 0;
  if (enable) {
    m->prof_used = 0; m->prof_ops.clear();
    while (m->prof_ev.size() < 16384) {
      cudaEvent_t e;
      HGB_CUDA(cudaEventCreate(&e));
      m->prof_ev.push_back(e);
    }
  }
  return HGB_OK;
}
extern "C" int hgb_model_profile_count(const hgb_model* m) { return (int)m->prof_ops.size(); }
extern "C" int hgb_model_profile_op(hgb_model* m, int i, int info[8], double* ms) {
  HGB_CHECK_ARG(i >= 0 && i < (int)m->prof_ops.size() && (size_t)(2 * i + 1) < m->prof_used + 0 + 1, "hgb_model_profile_op: index");
  const Op& o = m->prof_ops[i];
  info[0] = (int)o.type; info[1] = o.conv; info[2] = o.bn; info[3] = o.a0; info[4] = o.a1; info[5] = o.a2; info[6] = o.a3;
  info[7] = o.flag;
  float t = 0;
  HGB_CUDA(cudaEventElapsedTime(&t, m->prof_ev[2 * i], m->prof_ev[2 * i + 1]));
  *ms = t;
  return HGB_OK;
}

// {BatchNorm, dz act, y act} of profiled op i when it is a dgrad with the BatchNorm backward fused in, else -1s
extern "C" int hgb_model_profile_op_fused(hgb_model* m, int i, int out[3]) {
  HGB_CHECK_ARG(i >= 0 && i < (int)m->prof_ops.size(), "hgb_model_profile_op_fused: index");
  out[0] = m->prof_ops[i].fbn; out[1] = m->prof_ops[i].fz; out[2] = m->prof_ops[i].fy;
  return HGB_OK;
}

extern "C" int hgb_model_profile_read(hgb_model* m, double* total_ms, int* launches, double* flops) {
  HGB_CHECK_ARG(m && total_ms && launches && flops, "hgb_model_profile_read: null pointer");
  double tot = 0;
  for (size_t i = 0; i + 1 < m->prof_used; i += 2) {
    float ms = 0;
    HGB_CUDA(cudaEventElapsedTime(&ms, m->prof_ev[i], m->prof_ev[i + 1]));
    tot += ms;
  }
  *total_ms = tot; *launches = (int)(m->prof_used / 2); *flops = m->prof_flops;
  return HGB_OK;
}

extern "C" int64_t hgb_model_launch_count(const hgb_model* m) { return m->launches; }
