/*
 * hgb200.h -- C ABI of libhgb200.so, the B200 (sm_100a) implementation of the
 * stacked-hourglass heatmap path of MindlessBoid/single-person-pose-estimation.
 *
 * The reference has no FFI: its boundary is the Python call surface.  Each entry
 * point below names the reference function (file:line, relative to the reference
 * tree) whose arithmetic it replaces; the Python shim in
 * single-person-pose-estimation_b200/ binds these over ctypes and keeps the
 * reference's Python signatures (INTEGRATION.md shows the stubs).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (HGB_ERR_*); the message is
 *     available from hgb_last_error() (thread-local).
 *   - all data pointers are DEVICE pointers owned by the caller; the library never
 *     allocates per call.  `stream` is a cudaStream_t passed as void* (NULL = default).
 *   - calls are asynchronous on `stream`.
 *   - tensors are NHWC, C innermost, densely packed unless a pitch is given.
 */
#ifndef HGB200_H
#define HGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGB_OK               0
#define HGB_ERR_INVALID     -1   /* bad argument / unsupported shape */
#define HGB_ERR_CUDA        -2   /* CUDA runtime / driver error */
#define HGB_ERR_STATE       -3   /* call out of order (e.g. buffers not bound) */

#define HGB_F32   0
#define HGB_BF16  1
#define HGB_U8    2   /* input-path source images only */

/* loss kinds: trainer.py:224-245 string table */
#define HGB_LOSS_WEIGHTED_MSE           0   /* loss.py:2-21 */
#define HGB_LOSS_MSE                    1   /* tf.keras.losses.mean_squared_error, trainer.py:231 */
#define HGB_LOSS_IOU                    2   /* loss.py:23-28 */
#define HGB_LOSS_WEIGHTED_KEYPOINT_MSE  3   /* loss.py:30-36 */

const char* hgb_last_error(void);
int         hgb_version(void);
/* debug/tuning knobs (key, value); key outside [0,32) -> HGB_ERR_INVALID.  Every knob exists so that an optimisation
 * can be A/B-ed inside one process (tools_lanes_ab.py, tools_profile_step.py --debug k=v); 0 is always the default.
 *    1  wgrad: swap LBO/SBO of the MN-major descriptors (bring-up)      2  conv_output selector (tests)
 *    3  residual via per-thread loads instead of the TMA fetch          4  no fusion of BN-backward reductions (plan build)
 *    5  no weight-stationary / strip-reuse 3x3 paths                    7  PDL: 1 force on, 2 force off (default adaptive)
 *    8  replay every op in order on the caller's stream (no lanes)      9  SMs the skip lanes leave free (default 20)
 *   10  emit weight gradients before dgrads (plan build)               11  1: no wide 1x1 wgrad tiles, 2: also 256->256
 *   12  no strip reuse (HALO) in the 3x3 forward/dgrad kernel          13  no rolling-strip 3x3 weight-gradient kernel
 *   14  no deferred BatchNorm (plan build)                             15  HALO tile threshold (default 4 x #SMs)
 *   16  one thread issues all output boxes                             17  no BatchNorm folding in inference
 *   18  PDL trigger at kernel start instead of after the last load     19  PDL also in the multi-lane backward pass
 *   20  CTA cap of side-lane weight gradients (default 64, -1 none)    24  CTAs per sample in the decode kernel
 *   25  N = 256 GEMM tiles always with 8 epilogue warps                26  no BatchNorm-backward fusion into 1x1 dgrads (plan build)
 *   27  no programmatic dependent launch exception for small ops       28  decode ring geometry (10 * stages + vectors / thread)
 *   29  no shared bias-gradient pass of conv_1x1_2 / conv_1x1_3 (plan)  30  1: CTA pairs (cta_group::2) in the 3x3 strip kernel     */
int         hgb_debug_set(int key, int value);

/* ------------------------------------------------------------------------- */
/* Heatmap path                                                              */
/* ------------------------------------------------------------------------- */

/* dataset_builder.py:220-235 (np_gen_heatmaps) + utilities/data_utils.py:187-211 (gaussian).
 * kps_x, kps_y: (B,K) float32 heatmap-pixel coordinates; kps_v: (B,K) int32 visibility.
 * out: (B,H,W,K) float32, fully written (zeros + 7x7 sigma=1 stamps). */
int hgb_render_targets(const float* kps_x, const float* kps_y, const int32_t* kps_v,
                       int B, int H, int W, int K, float* out, void* stream);

/* bytes of scratch hgb_loss_fwd_bwd needs for (B,K) */
int64_t hgb_loss_workspace_bytes(int B, int K);

/* One Keras output's loss and gradient (loss.py:2-36; reduction = mean over everything the
 * loss fn returns).  y_true f32 (B,H,W,K); y_pred / grad in pred_dtype /
 * grad_dtype.  inv_count = 1/(B_global*H*W*K) for the MSE family, 1/(B_global*K) for IoU, so
 * a batch shard produces its share of the global mean.  loss_acc: device double, ADDED to
 * (caller zeroes).  grad may be NULL (loss only). */
int hgb_loss_fwd_bwd(int kind, const float* y_true, const void* y_pred, int pred_dtype,
                     int B, int H, int W, int K, double inv_count,
                     double* loss_acc, void* grad, int grad_dtype,
                     void* workspace, void* stream);

/* The tensor the reference loss functions return: (B,H,W) f32 for the MSE family
 * (loss.py:21,36), (B,) f32 for IoU (loss.py:28). */
int hgb_loss_map(int kind, const float* y_true, const float* y_pred,
                 int B, int H, int W, int K, float* out, void* workspace, void* stream);

/* utilities/data_utils.py:100-132 (version 1) and :135-183 (version 2).
 * heatmaps (B,H,W,K) in `dtype`, H == W required (data_utils.py:122 divides by height).
 * out_idx (B,K,4) int32 = [argmax index, x, y, patch argmax index]; out_kpts (B,K,3) f32 =
 * [x + dx, y + dy, conf] or zeros when conf <= float32(conf_threshold) (numpy >= 2 semantics). The
 * input is NOT modified; the (1,1) element of the clipped 3x3 window is read as 0. */
int hgb_decode(const void* heatmaps, int dtype, int B, int H, int W, int K,
               double conf_threshold, int version, int32_t* out_idx, float* out_kpts, void* stream);

/* eval.py:62-88: per-joint PCK counters.  All (N,K) double except vs (N,K) int32 and
 * bbox_wh (N,2) double (original COCO bbox width,height).  counts: int32[2*K] =
 * correct[K] then visible[K], ADDED to (caller zeroes). */
int hgb_pck_reduce(const double* xs_pred, const double* ys_pred, const double* xs_gt, const double* ys_gt,
                   const int32_t* vs, const double* bbox_wh, int N, int K, double pck_threshold,
                   int32_t* counts, void* stream);

/* COCO object-keypoint-similarity of each prediction with its own annotation (the arithmetic
 * eval.py:39-49 delegates to pycocotools COCOeval.computeOks).  bbox_xywh (N,4), area (N,),
 * oks_out (N,) double. K <= 17. */
int hgb_oks_similarity(const double* xs_pred, const double* ys_pred, const double* xs_gt, const double* ys_gt,
                       const int32_t* vs, const double* area, const double* bbox_xywh, int N, int K,
                       double* oks_out, void* stream);

/* ------------------------------------------------------------------------- */
/* Input path: frame -> network input / keypoints      */
/* ------------------------------------------------------------------------- */

/* tf.image.convert_image_dtype(uint8 -> float32) + crop_and_pad (utilities/data_utils.py:48-98) +
 * tf.image.resize(bilinear, half-pixel centres, no antialias) in one pass: demo.py:44-50 (N person crops of a frame),
 * dataset_builder.py:99 / :133 (whole image, crop_xywh = NULL).
 * src_ptrs: DEVICE array of N device pointers to (h,w,3) images of `src_dtype` (HGB_U8 or HGB_F32; U8 values are
 * multiplied by float32(1/255)); src_hw: device (N,2) int32 [height,width]; crop_xywh: device (N,4) int32
 * [x0, y0, crop_w, crop_h] -- crop pixel (cy,cx) reads source pixel (cy+y0, cx+x0), zero outside the source (the host
 * shim derives these integers with the reference's own Python arithmetic); out: (N,out_h,out_w,3) float32. */
int hgb_crop_resize(const void* const* src_ptrs, const int32_t* src_hw, int src_dtype, const int32_t* crop_xywh, int N,
                    int out_h, int out_w, float* out, void* stream);

/* Image half of DatasetBuilder.np_augment_1 (dataset_builder.py:143-172): optional Fliplr, then imgaug
 * Affine(scale, rotate) = cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) with OpenCV's 10-bit fixed-point source
 * coordinates and 5-bit interpolation fractions.  images/out: (N,H,W,3) float32 (out != images); inv_mats: device (N,6)
 * double, the INVERSE 2x3 map OpenCV derives from the forward matrix (host shim); flip: device (N,) int32. */
int hgb_augment_affine(const float* images, const double* inv_mats, const int32_t* flip, int N, int H, int W, float* out,
                       void* stream);

/* Keypoint half of np_augment_1 + flip_labels (dataset_builder.py:150-185, 270-300): joints with v <= 0 become (0,0);
 * flip maps x -> label_w - x and slot k takes joint flip_partner[k] (x, y and v); forward affine (N,6) double about the
 * label-map centre; coordinates of slots whose (swapped) v <= 0 are zeroed.  kps_* (N,K); out_x/out_y (N,K) float32. */
int hgb_augment_keypoints(const float* kps_x, const float* kps_y, const int32_t* kps_v, const int32_t* flip,
                          const double* fwd_mats, const int32_t* flip_partner, int N, int K, int label_w, float* out_x,
                          float* out_y, void* stream);

/* DatasetBuilder.augment_2 (dataset_builder.py:190-204) with its four random draws supplied: params (N,4) float32 =
 * [brightness delta, contrast factor, saturation factor, hue delta] -> tf.image.adjust_brightness, adjust_contrast
 * (per-channel mean), adjust_saturation (HSV), adjust_hue, then (x - min)/(max - min) over the whole image.
 * images (N,H,W,3) float32, IN PLACE.  workspace: hgb_color_workspace_bytes(N) bytes. */
int64_t hgb_color_workspace_bytes(int N);
int hgb_color_augment(float* images, const float* params, int N, int H, int W, void* workspace, void* stream);

/* CRC-32C (Castagnoli) of a HOST buffer: the checksum of the TFRecord framing (length and payload, masked as
 * ((crc >> 15 | crc << 17) + 0xa282ead8) by the caller) that tf.data.TFRecordDataset verifies (dataset_builder.py:39,48,63). */
uint32_t hgb_crc32c(const void* host_data, int64_t len);            /* SSE4.2 CRC32 instruction when the CPU has it */
uint32_t hgb_crc32c_portable(const void* host_data, int64_t len);   /* slicing-by-8 tables; the definition the fast path is tested against */

/* tf.data.TFRecordDataset (dataset_builder.py:39,48,63): walk the framing of a TFRecord file held in HOST memory (e.g. mmap)
 * from byte `start`: for up to `cap` records writes the payload offset and length, verifies the masked CRC-32C of the length
 * and of the payload when `verify`, and stores the offset of the first unread byte in *next.  Returns the number of records
 * (0 at the end of the file) or HGB_ERR_INVALID with "truncated ..." / "corrupted ..." in hgb_last_error(). */
int64_t hgb_tfrecord_scan(const uint8_t* data, int64_t len, int64_t start, int verify, int64_t* offsets, int64_t* lengths,
                          int64_t cap, int64_t* next);

/* tf.io.parse_single_example (dataset_builder.py:262): one serialized tf.train.Example in a HOST buffer.  Returns the number of
 * features (>= 0) or an error.  table: max_features rows of 6 int64 = [name offset, name length, kind (1 bytes, 2 float,
 * 3 int64, 0 unset), start, count, extra]: for float / int64 lists `start` indexes fvals / ivals and `count` values follow
 * (packed or unpacked encodings); for bytes lists `start` / `extra` are the offset / length of the first value in `data`
 * and `count` the number of values.  HGB_ERR_STATE when a capacity (max_features, fcap, icap) is too small. */
int hgb_example_parse(const uint8_t* data, int64_t len, int max_features, int64_t* table, float* fvals, int64_t fcap,
                      int64_t* ivals, int64_t icap);

/* tf.image.decode_image (dataset_builder.py:263) / tf.io.decode_jpeg (gen_tfrecords.py:112), JPEG only.
 * hgb_jpeg_info parses the frame header of a HOST buffer; hgb_jpeg_decode decodes N HOST streams into N caller-owned
 * DEVICE buffers of (h,w,3) interleaved RGB uint8 (grey streams are replicated to 3 channels) through nvJPEG, which is
 * opened on first use (HGB_ERR_STATE if the toolkit library is missing).  datas / lens / outs / hw are HOST arrays;
 * hw (N,2) = [height,width] the outputs were sized for (checked against the streams). */
int hgb_jpeg_info(const uint8_t* host_data, int64_t len, int32_t* height, int32_t* width, int32_t* components);
int hgb_jpeg_decode(const uint8_t* const* datas, const int64_t* lens, int N, uint8_t* const* outs, const int32_t* hw, void* stream);

/* ------------------------------------------------------------------------- */
/* Hourglass network  (model/hourglass.py:5-206)                             */
/* ------------------------------------------------------------------------- */

typedef struct hgb_model hgb_model;

typedef struct {
  int num_classes;    /* 17 */
  int num_stacks;     /* model/hourglass.py:21 */
  int num_channels;   /* 256; must be a multiple of 128 */
  int in_h, in_w;     /* 256, 256 (multiples of 64) */
  int activation;     /* predict_activation: 0 linear, 1 sigmoid */
  int batch;          /* per-device batch the execution plan is built for */
  int training;       /* 1: plan keeps activations and allocates the backward pass */
  int mobile;         /* 1: bottleneck_block_mobile (model/hourglass.py:9-11,209-231): every convolution of a bottleneck is a
                         SeparableConv2D = depthwise k x k (per-channel stencil kernel) + pointwise 1x1 (the tcgen05 GEMM) */
} hgb_model_config;

/* buffers the caller allocates and binds */
#define HGB_BUF_PARAMS   0   /* fp32: trainable scalars first, then BN moving statistics */
#define HGB_BUF_GRADS    1   /* fp32, trainable count */
#define HGB_BUF_ADAM_M   2
#define HGB_BUF_ADAM_V   3
#define HGB_BUF_ARENA    4   /* activations, bf16 weight shadows, scratch */

int     hgb_model_create(const hgb_model_config* cfg, int device, hgb_model** out);
int     hgb_model_destroy(hgb_model* m);
int64_t hgb_model_param_count(const hgb_model* m, int trainable_only);
int64_t hgb_model_buffer_bytes(const hgb_model* m, int which);
int     hgb_model_bind(hgb_model* m, int which, void* ptr, int64_t bytes);

/* parameter table: names are the Keras layer names of model/hourglass.py plus
 * /kernel /bias /gamma /beta /moving_mean /moving_variance; kernel dims are HWIO. */
int     hgb_model_num_tensors(const hgb_model* m);
int     hgb_model_tensor_info(const hgb_model* m, int index, const char** name, int* rank,
                              int64_t dims[4], int64_t* offset, int* trainable);
int     hgb_model_num_convs(const hgb_model* m);
/* per-conv FLOPs (2*MAC) of one forward pass at the plan's batch, and its class */
int     hgb_model_conv_info(const hgb_model* m, int index, const char** name, int* k, int* cin, int* cout,
                            int* h, int* w, double* flops);

/* re-derive the bf16 GEMM-layout weights from the fp32 parameters (after load / Adam) */
int hgb_model_sync_weights(hgb_model* m, void* stream);

/* forward.  images: (B,in_h,in_w,3) f32.  training=1 uses batch statistics and updates
 * the moving averages (Keras fit); 0 uses moving statistics (predict / evaluate).
 * heatmaps_out: array of num_stacks device pointers (B,H/4,W/4,num_classes) f32, entries may be
 * NULL to skip that stack's output copy. */
int hgb_model_forward(hgb_model* m, const float* images, int training, float* const* heatmaps_out, void* stream);

/* losses of all stacks + gradient wrt each stack's pre-activation output; y_true f32
 * (B,h,w,classes) shared by every stack (Keras compile with one loss fn, trainer.py:35).
 * loss_acc: device double[num_stacks], added to.  Needs a preceding training forward. */
int hgb_model_loss(hgb_model* m, int kind, const float* y_true, double inv_count, double* loss_acc, void* stream);

/* backward over segments [seg_lo, seg_hi): segment 0 = front module, 1+i = stack i.
 * Must be called from the highest segment down.  Gradients are written (not accumulated). */
int hgb_model_num_segments(const hgb_model* m);
int hgb_model_backward(hgb_model* m, int seg_lo, int seg_hi, void* stream);
/* Same, but `stream` is NOT ordered after the lanes on return: the main chain of the next segment starts while this
 * segment's weight gradients are still draining.  hgb_model_lanes_join(m, s, 0) orders a side stream `s` (the one the
 * gradient all-reduce of the segment is issued on) after everything issued so far; hgb_model_lanes_join(m, stream, 1)
 * must be called on the caller's stream before the optimizer step. */
int hgb_model_backward_nojoin(hgb_model* m, int seg_lo, int seg_hi, void* stream);
int hgb_model_lanes_join(hgb_model* m, void* stream, int is_caller);
/* contiguous range of the gradient buffer that segment `seg` owns (for bucketed allreduce) */
int hgb_model_segment_grads(const hgb_model* m, int seg, int64_t* offset, int64_t* count);

/* Keras legacy Adam (trainer.py:31): w -= lr_t*m/(sqrt(v)+eps) with
 * lr_t = lr*sqrt(1-b2^t)/(1-b1^t); grads are multiplied by grad_scale first (1 when hgb_model_loss was given
 * inv_count of the GLOBAL batch -- the summed buckets are then the global-mean gradient -- 1/world_size when the
 * ranks normalised by their local batch).  Also refreshes the bf16 GEMM weights. */
int hgb_model_adam_step(hgb_model* m, double lr, double beta1, double beta2, double eps, int64_t t,
                        double grad_scale, void* stream);

/* ------------------------------------------------------------------------- */
/* Data parallelism: the one collective of the path      */
/* ------------------------------------------------------------------------- */
/* The reference has no distributed code (trainer.py:49-56 is a single-device fit); training shards the batch over one
 * process per GPU and exchanges gradients once per step.  The communicator is NCCL (NVLink / NVSwitch), bound at run time
 * (dlopen of libnccl.so.2: inside a torch process that is the library torch already loaded).
 *   hgb_comm_unique_id   rank 0 creates the 128-byte id; the caller ships it to the other ranks (any transport)
 *   hgb_comm_init        collective over the nranks processes; one communicator per process / device
 *   hgb_comm_allreduce_f32   in-place sum of a device buffer, asynchronous on `stream` */
#define HGB_COMM_UNIQUE_ID_BYTES 128
typedef struct hgb_comm hgb_comm;
int hgb_comm_unique_id(void* id_out, int bytes);
int hgb_comm_init(int nranks, int rank, const void* unique_id, hgb_comm** out);
int hgb_comm_destroy(hgb_comm* c);
int hgb_comm_info(const hgb_comm* c, int* nranks, int* rank, int* nccl_version);
int hgb_comm_allreduce_f32(hgb_comm* c, float* buf, int64_t count, void* stream);
/* Attach a communicator to a plan (NULL detaches).  sync_bn != 0: the BatchNorm batch statistics (forward sums, backward
 * sums) of every layer are all-reduced as well -- one 2*C-float all-reduce per BatchNorm and pass, ops replayed in plan order
 * on the caller's stream -- so N ranks compute exactly the single-device step of the concatenated batch (parity tests; the
 * default, like Keras layers under mirrored replicas, is per-replica statistics). */
int hgb_model_set_comm(hgb_model* m, hgb_comm* c, int sync_bn);
/* Sum over all ranks of the gradient range owned by segments [seg_lo, seg_hi) -- consecutive segments are one contiguous
 * bucket of the flat gradient buffer -- asynchronous on `stream` (order it after the segments' backward with
 * hgb_model_lanes_join(m, stream, 0)). */
int hgb_grad_allreduce_bucket(hgb_model* m, int seg_lo, int seg_hi, void* stream);

/* debugging / layer-wise parity: where conv `index` wrote its (bias+activation) output, bf16 NHWC,
 * dims = {N,H,W,C_padded}, as a byte offset into the bound arena */
int hgb_model_conv_output(const hgb_model* m, int index, int64_t* arena_offset, int dims[4]);

/* Deferred BatchNorm: a BN whose output feeds only 1x1 convolutions is never written to memory; those convolutions
 * (forward GEMM and weight gradient) normalise their operand tile in shared memory, bit-identically to the stand-alone
 * pass.  Returns the index of the BN applied to the INPUT of conv `conv` inside the kernel (the conv's input tensor is
 * then the pre-BN tensor), or -1.  Forward op info: flag & 0xffff = 1 + that BN index, flag & 0x10000 = this launch
 * stores the saved statistics / moving averages; weight-gradient op info: bn = that BN index. */
int hgb_model_conv_input_bn(const hgb_model* m, int conv);

/* plan introspection and single-op stepping: lets a test replay EVERY op of the real execution plan
 * (forward and backward) against an fp32 reference computed from the device's own input tensors.
 * op info = {type, conv, bn, a0, a1, a2, a3, flag}; see csrc/model.cu (OpType) for the roles.  Types 15-17 (mobile variant):
 * depthwise forward / input gradient / weight gradient, `conv` = index for hgb_model_dw_detail. */
int hgb_model_num_ops(const hgb_model* m, int seg, int backward);
int hgb_model_op_info(const hgb_model* m, int seg, int backward, int index, int info[8]);
int hgb_model_act_info(const hgb_model* m, int act, int64_t* arena_offset, int dims[4]);
/* 1x1 B_DGRAD ops that carry the BatchNorm backward of their input gradient (the stand-alone bn_bwd_apply pass is fused into
 * the GEMM's operand prologue: dz and y in, dp computed in shared memory, fed to the MMAs and stored for the weight gradient):
 * out = {BatchNorm index, dz tensor, y tensor}; the dp tensor is the op's a0.  {-1,-1,-1} for every other op. */
int hgb_model_op_fused_bn(const hgb_model* m, int seg, int backward, int index, int out[3]);
int hgb_model_run_op(hgb_model* m, int seg, int backward, int index, const float* images, int training, void* stream);
int hgb_model_conv_detail(const hgb_model* m, int conv, int info[8], int64_t offs[2]);
int hgb_model_bn_detail(const hgb_model* m, int bn, int64_t offs[8]);
int hgb_model_dw_detail(const hgb_model* m, int dw, int info[4], int64_t* w_off);   /* mobile: depthwise stage {k, c, h, w} */
int hgb_model_head_buffers(const hgb_model* m, int stack, int64_t offs[2]);
int hgb_model_begin_step(hgb_model* m, void* stream);

/* live timing of one convolution class inside a running step (bench.py roofline): CUDA event pairs
 * are recorded on the launching stream around every matching launch.  op_type: the plan's op type (1 forward conv,
 * 8 wgrad, 9 dgrad, 2 BatchNorm forward, 7 BatchNorm backward, 12 pool backward ...); (k, cin, cout, h) select a
 * convolution class by its Keras shape, k = 0 selects a non-convolution class by (channels = cout, height = h) of its
 * first tensor.  Read after a stream sync: total milliseconds, number of launches, and (convolutions) their
 * algorithmic FLOPs (2*MAC, forward count). */
int hgb_model_profile_conv(hgb_model* m, int enable, int op_type, int k, int cin, int cout, int h);
int hgb_model_profile_read(hgb_model* m, double* total_ms, int* launches, double* flops);
/* time EVERY op (event pair per op) of the steps that follow; read back per op after a stream sync */
int hgb_model_profile_all(hgb_model* m, int enable);
int hgb_model_profile_count(const hgb_model* m);
int hgb_model_profile_op(hgb_model* m, int i, int info[8], double* ms);
int hgb_model_profile_op_fused(hgb_model* m, int i, int out[3]);   /* see hgb_model_op_fused_bn */

/* Execution lanes.  Forward and backward run as a static schedule over several CUDA streams owned by the
 * handle (the skip bottleneck of every hourglass level beside the deeper sub-hourglass, weight gradients
 * beside the dgrad chain); `stream` only forks and joins.  Cross-lane dependencies are derived from every
 * op's read/write byte ranges when the plan is created.  hgb_debug_set(8, 1) replays everything in order
 * on `stream` instead.  These three calls expose the schedule so a test can prove that every pair of
 * conflicting ops is ordered:
 *   sched_op: info = {segment, op index, lane, signals another lane}; deps = (lane, sequence index) pairs
 *   sched_access: rows of (space 0 = arena bytes | 1 = gradient floats, lo, hi, is_write); returns #rows */
int hgb_model_sched_count(const hgb_model* m, int backward);
int hgb_model_sched_op(const hgb_model* m, int backward, int k, int info[4], int* ndeps, int deps[16]);
int hgb_model_sched_access(const hgb_model* m, int backward, int k, int cap, int64_t* ranges);

/* number of kernels this handle has launched since creation (bench "gpu_launches") */
int64_t hgb_model_launch_count(const hgb_model* m);

/* ------------------------------------------------------------------------- */
/* Stand-alone convolution GEMM entry points (unit tests / kernel benchmarks) */
/* ------------------------------------------------------------------------- */

/* out[M,Cout] = act(conv_kxk(in[N,H,W,Cin], w) + bias) (+res1 +res2), bf16 NHWC, fp32 accumulate.
 * w: bf16 [Cout][k*k*Cin] (tap-major, channel-minor).  k in {1,3}; Cin % 64 == 0; Cout % 16 == 0, <= 256.
 * ldc: row pitch of out/res in elements.  stats: optional float[2*Cout] (sum, sum of squares of the
 * stored values), ADDED to. tap_sign = +1 forward, -1 for dgrad (taps mirrored). */
int hgb_conv_gemm(const void* in, const void* w, const float* bias, const void* res1, const void* res2,
                  void* out, float* stats, int N, int H, int W, int Cin, int Cout, int ksize,
                  int relu, int ldc, int tap_sign, void* stream);

/* dw[Cout][k*k*Cin] (fp32, ADDED to) = sum_pixels dy[p][Cout]^T x_tap[p][Cin]  (weight gradient) */
int hgb_conv_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                   int ksize, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HGB200_H */
