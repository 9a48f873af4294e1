"""Batched device entry points over libhgb200 (torch tensors are only buffer holders).

Every function takes/returns CUDA torch tensors and launches on the current torch stream.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


def _torch():
    return _lib.require_cuda()


def _dev(x, dtype):
    """Move an array-like to the current CUDA device with the given torch dtype (contiguous)."""
    torch = _torch()
    if isinstance(x, torch.Tensor):
        return x.to(device="cuda", dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda").to(dtype).contiguous()


# ------------------------------------------------------------------ target rendering
def render_targets(kps_x, kps_y, kps_v, height: int = 64, width: int = 64, out=None):
    """Batched DatasetBuilder.np_gen_heatmaps (dataset_builder.py:220-235) -> (B,H,W,K) f32."""
    torch = _torch()
    kx = _dev(kps_x, torch.float32)
    ky = _dev(kps_y, torch.float32)
    kv = _dev(kps_v, torch.int32)
    if kx.dim() != 2 or kx.shape != ky.shape or kx.shape != kv.shape:
        raise ValueError("kps_x, kps_y, kps_v must all be (B, K)")
    B, K = kx.shape
    if out is None:
        out = torch.empty((B, height, width, K), dtype=torch.float32, device="cuda")
    check(lib.hgb_render_targets(ptr(kx), ptr(ky), ptr(kv), B, height, width, K, ptr(out), stream_ptr()))
    return out


# ------------------------------------------------------------------ losses
def _tdtype(t):
    torch = _torch()
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise ValueError(f"unsupported dtype {t.dtype}")


def loss_fwd_bwd(kind, y_true, y_pred, want_grad=True, grad_dtype=None, global_batch=None):
    """Keras-reduced scalar loss of one output and d(loss)/d(y_pred) in one pass.

    Returns (loss: 0-dim float64 CUDA tensor, grad or None)."""
    torch = _torch()
    k = _lib.LOSS_KINDS[kind] if isinstance(kind, str) else int(kind)
    yt = _dev(y_true, torch.float32)
    yp = y_pred if isinstance(y_pred, torch.Tensor) and y_pred.is_cuda and y_pred.is_contiguous() else _dev(y_pred, torch.float32)
    if yt.shape != yp.shape or yt.dim() != 4:
        raise ValueError("y_true and y_pred must both be (B,H,W,K)")
    B, H, W, K = yt.shape
    gb = B if global_batch is None else int(global_batch)
    inv = 1.0 / (gb * K) if k == 2 else 1.0 / (gb * H * W * K)
    loss = torch.zeros((), dtype=torch.float64, device="cuda")
    grad = None
    if want_grad:
        grad = torch.empty_like(yp, dtype=grad_dtype or yp.dtype)
    ws = torch.empty(int(lib.hgb_loss_workspace_bytes(B, K)), dtype=torch.uint8, device="cuda")
    check(lib.hgb_loss_fwd_bwd(k, ptr(yt), ptr(yp), _tdtype(yp), B, H, W, K, inv, ptr(loss), ptr(grad),
                               _tdtype(grad) if grad is not None else _lib.F32, ptr(ws), stream_ptr()))
    return loss, grad


def loss_map(kind, y_true, y_pred):
    """The tensor the reference loss function returns: (B,H,W) or, for IoU, (B,)."""
    torch = _torch()
    k = _lib.LOSS_KINDS[kind] if isinstance(kind, str) else int(kind)
    yt = _dev(y_true, torch.float32)
    yp = _dev(y_pred, torch.float32)
    if yt.shape != yp.shape or yt.dim() != 4:
        raise ValueError("y_true and y_pred must both be (B,H,W,K)")
    B, H, W, K = yt.shape
    out = torch.empty((B,) if k == 2 else (B, H, W), dtype=torch.float32, device="cuda")
    ws = torch.empty(int(lib.hgb_loss_workspace_bytes(B, K)), dtype=torch.uint8, device="cuda")
    check(lib.hgb_loss_map(k, ptr(yt), ptr(yp), B, H, W, K, ptr(out), ptr(ws), stream_ptr()))
    return out


# ------------------------------------------------------------------ decode
def decode_batch(heatmaps, conf_threshold: float = 1e-6, version: int = 2):
    """(B,H,W,K) f32/bf16 CUDA tensor -> (idx int32 (B,K,4), kpts f32 (B,K,3)) on device."""
    torch = _torch()
    hm = heatmaps
    if not (isinstance(hm, torch.Tensor) and hm.is_cuda):
        hm = _dev(hm, torch.float32)
    hm = hm.contiguous()
    if hm.dim() != 4:
        raise ValueError("heatmaps must be (B,H,W,K)")
    B, H, W, K = hm.shape
    idx = torch.empty((B, K, 4), dtype=torch.int32, device="cuda")
    kp = torch.empty((B, K, 3), dtype=torch.float32, device="cuda")
    if version not in (1, 2) or H != W:
        raise ValueError("decode: version must be 1 or 2 and maps must be square (data_utils.py:122)")
    if B == 0:
        return idx, kp
    check(lib.hgb_decode(ptr(hm), _tdtype(hm), B, H, W, K, float(conf_threshold), int(version), ptr(idx), ptr(kp),
                         stream_ptr()))
    return idx, kp


# ------------------------------------------------------------------ scoring
def pck_counts(xs_pred, ys_pred, xs_gt, ys_gt, vs, bbox_wh, pck_threshold: float = 0.05):
    """eval.py:62-88 counters -> (correct[K], visible[K]) int64 numpy."""
    torch = _torch()
    xp, yp, xg, yg = (_dev(a, torch.float64) for a in (xs_pred, ys_pred, xs_gt, ys_gt))
    v = _dev(vs, torch.int32)
    bb = _dev(bbox_wh, torch.float64)
    N, K = xp.shape
    counts = torch.zeros(2 * K, dtype=torch.int32, device="cuda")
    check(lib.hgb_pck_reduce(ptr(xp), ptr(yp), ptr(xg), ptr(yg), ptr(v), ptr(bb), N, K, float(pck_threshold),
                             ptr(counts), stream_ptr()))
    c = counts.cpu().numpy().astype(np.int64)
    return c[:K], c[K:]


def oks_similarity(xs_pred, ys_pred, xs_gt, ys_gt, vs, area, bbox_xywh):
    torch = _torch()
    xp, yp, xg, yg = (_dev(a, torch.float64) for a in (xs_pred, ys_pred, xs_gt, ys_gt))
    v = _dev(vs, torch.int32)
    ar = _dev(area, torch.float64)
    bb = _dev(bbox_xywh, torch.float64)
    N, K = xp.shape
    out = torch.empty(N, dtype=torch.float64, device="cuda")
    check(lib.hgb_oks_similarity(ptr(xp), ptr(yp), ptr(xg), ptr(yg), ptr(v), ptr(ar), ptr(bb), N, K, ptr(out),
                                 stream_ptr()))
    return out


# ------------------------------------------------------------------ input path
def crop_resize(sources, crops=None, out_h: int = 256, out_w: int = 256, source_index=None):
    """Fused uint8->float32 conversion, crop_and_pad and bilinear resize (demo.py:44-50, dataset_builder.py:99,133).

    sources: list of (h,w,3) uint8 or float32 arrays / CUDA tensors (one dtype per call); crops: None (whole images) or
    (N,4) integers [x0, y0, crop_w, crop_h] from `data_utils.crop_and_pad_params`; source_index: which source each of the
    N outputs reads (default: output n reads source n).  Returns (N,out_h,out_w,3) float32 on device and keeps nothing."""
    torch = _torch()
    srcs = []
    for im in sources:
        if not isinstance(im, torch.Tensor):
            im = torch.as_tensor(np.ascontiguousarray(im))
        if im.dtype not in (torch.uint8, torch.float32):
            im = im.to(torch.float32)
        if im.dim() != 3 or im.shape[2] != 3:
            raise ValueError("source images must be (h, w, 3)")
        srcs.append(im.to(device="cuda").contiguous())
    if not srcs:
        return torch.empty((0, out_h, out_w, 3), dtype=torch.float32, device="cuda")
    if any(t.dtype != srcs[0].dtype for t in srcs):
        raise ValueError("all sources of one call must share a dtype")
    index = list(range(len(srcs))) if source_index is None else [int(i) for i in source_index]
    N = len(index)
    if N == 0:
        return torch.empty((0, out_h, out_w, 3), dtype=torch.float32, device="cuda")
    table = torch.tensor([srcs[i].data_ptr() for i in index], dtype=torch.int64).cuda()
    hw = torch.tensor([[srcs[i].shape[0], srcs[i].shape[1]] for i in index], dtype=torch.int32).cuda()
    crop_t = None
    if crops is not None:
        crop_t = torch.as_tensor(np.asarray(crops, dtype=np.int32).reshape(N, 4)).cuda()
    out = torch.empty((N, out_h, out_w, 3), dtype=torch.float32, device="cuda")
    check(lib.hgb_crop_resize(ptr(table), ptr(hw), _lib.U8 if srcs[0].dtype == torch.uint8 else _lib.F32, ptr(crop_t), N,
                              out_h, out_w, ptr(out), stream_ptr()))
    for t in srcs:                                        # the launch is asynchronous: keep sources alive on this stream
        t.record_stream(torch.cuda.current_stream())
    return out


def augment_affine(images, inv_mats, flip, out=None):
    """Fliplr + cv2.warpAffine(INTER_LINEAR, constant 0) of (N,H,W,3) float32 images; inv_mats (N,2,3) float64 inverse maps."""
    torch = _torch()
    x = _dev(images, torch.float32)
    N, H, W, _ = x.shape
    m = _dev(np.asarray(inv_mats, np.float64).reshape(N, 6), torch.float64)
    f = _dev(np.asarray(flip).astype(np.int32).reshape(N), torch.int32)
    if out is None:
        out = torch.empty_like(x)
    check(lib.hgb_augment_affine(ptr(x), ptr(m), ptr(f), N, H, W, ptr(out), stream_ptr()))
    return out


def augment_keypoints(kps_x, kps_y, kps_v, flip, fwd_mats, flip_partner, label_w: int = 64):
    torch = _torch()
    kx, ky, kv = _dev(kps_x, torch.float32), _dev(kps_y, torch.float32), _dev(kps_v, torch.int32)
    N, K = kx.shape
    m = _dev(np.asarray(fwd_mats, np.float64).reshape(N, 6), torch.float64)
    f = _dev(np.asarray(flip).astype(np.int32).reshape(N), torch.int32)
    pt = _dev(np.asarray(flip_partner, np.int32).reshape(K), torch.int32)
    ox, oy = torch.empty_like(kx), torch.empty_like(ky)
    check(lib.hgb_augment_keypoints(ptr(kx), ptr(ky), ptr(kv), ptr(f), ptr(m), ptr(pt), N, K, int(label_w), ptr(ox), ptr(oy),
                                    stream_ptr()))
    return ox, oy


def color_augment(images, params):
    """IN PLACE on a CUDA float32 (N,H,W,3) tensor; params (N,4) = brightness delta, contrast, saturation, hue delta."""
    torch = _torch()
    if not (isinstance(images, torch.Tensor) and images.is_cuda and images.dtype == torch.float32 and images.is_contiguous()):
        raise ValueError("color_augment works in place on a contiguous CUDA float32 tensor")
    N, H, W, _ = images.shape
    p = _dev(np.asarray(params, np.float32).reshape(N, 4) if not isinstance(params, torch.Tensor) else params, torch.float32)
    ws = torch.empty(int(lib.hgb_color_workspace_bytes(N)), dtype=torch.uint8, device="cuda")
    check(lib.hgb_color_augment(ptr(images), ptr(p), N, H, W, ptr(ws), stream_ptr()))
    return images


# ------------------------------------------------------------------ convolution kernels (tests / microbench)
def conv_gemm(x, w, bias=None, res1=None, res2=None, ksize=1, relu=False, tap_sign=1, stats=None, out=None):
    """x: (N,H,W,Cin) bf16; w: (Cout, k*k*Cin) bf16 -> (N,H,W,Cout) bf16."""
    torch = _torch()
    N, H, W, Cin = x.shape
    Cout = w.shape[0]
    if out is None:
        out = torch.empty((N, H, W, Cout), dtype=torch.bfloat16, device="cuda")
    check(lib.hgb_conv_gemm(ptr(x), ptr(w), ptr(bias), ptr(res1), ptr(res2), ptr(out), ptr(stats), N, H, W, Cin, Cout,
                            ksize, int(relu), out.shape[-1], tap_sign, stream_ptr()))
    return out


def conv_wgrad(x, dy, ksize=1, dw=None):
    """x: (N,H,W,Cin) bf16; dy: (N,H,W,Cout) bf16 -> dw (Cout, k*k*Cin) f32 (accumulated into dw if given)."""
    torch = _torch()
    N, H, W, Cin = x.shape
    Cout = dy.shape[-1]
    if dw is None:
        dw = torch.zeros((Cout, ksize * ksize * Cin), dtype=torch.float32, device="cuda")
    check(lib.hgb_conv_wgrad(ptr(x), ptr(dy), ptr(dw), N, H, W, Cin, Cout, ksize, stream_ptr()))
    return dw
