"""ctypes binding of libhgb200.so (include/hgb200.h).  No compute happens here."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhgb200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C single-person-pose-estimation_b200/csrc`. There is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

F32, BF16, U8 = 0, 1, 2
LOSS_KINDS = {"weighted_mse": 0, "mse": 1, "iou": 2, "weighted_keypoint_mse": 3}
BUF_PARAMS, BUF_GRADS, BUF_ADAM_M, BUF_ADAM_V, BUF_ARENA = range(5)

vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double


class ModelConfig(C.Structure):
    _fields_ = [("num_classes", i32), ("num_stacks", i32), ("num_channels", i32), ("in_h", i32), ("in_w", i32),
                ("activation", i32), ("batch", i32), ("training", i32), ("mobile", i32)]


# name -> (restype, argtypes); kept in one table so tests can check every symbol of hgb200.h is exported
PROTOTYPES = {
    "hgb_last_error": (C.c_char_p, []),
    "hgb_version": (i32, []),
    "hgb_debug_set": (i32, [i32, i32]),
    "hgb_render_targets": (i32, [vp, vp, vp, i32, i32, i32, i32, vp, vp]),
    "hgb_loss_workspace_bytes": (i64, [i32, i32]),
    "hgb_loss_fwd_bwd": (i32, [i32, vp, vp, i32, i32, i32, i32, i32, f64, vp, vp, i32, vp, vp]),
    "hgb_loss_map": (i32, [i32, vp, vp, i32, i32, i32, i32, vp, vp, vp]),
    "hgb_decode": (i32, [vp, i32, i32, i32, i32, i32, f64, i32, vp, vp, vp]),
    "hgb_pck_reduce": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, f64, vp, vp]),
    "hgb_oks_similarity": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp]),
    "hgb_crop_resize": (i32, [vp, vp, i32, vp, i32, i32, i32, vp, vp]),
    "hgb_augment_affine": (i32, [vp, vp, vp, i32, i32, i32, vp, vp]),
    "hgb_augment_keypoints": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp]),
    "hgb_color_workspace_bytes": (i64, [i32]),
    "hgb_color_augment": (i32, [vp, vp, i32, i32, i32, vp, vp]),
    "hgb_crc32c": (C.c_uint32, [vp, i64]),
    "hgb_crc32c_portable": (C.c_uint32, [vp, i64]),
    "hgb_tfrecord_scan": (i64, [vp, i64, i64, i32, vp, vp, i64, vp]),
    "hgb_example_parse": (i32, [vp, i64, i32, vp, vp, i64, vp, i64]),
    "hgb_jpeg_info": (i32, [vp, i64, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "hgb_jpeg_decode": (i32, [vp, vp, i32, vp, vp, vp]),
    "hgb_conv_gemm": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
    "hgb_conv_wgrad": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "hgb_model_create": (i32, [C.POINTER(ModelConfig), i32, C.POINTER(vp)]),
    "hgb_model_destroy": (i32, [vp]),
    "hgb_model_param_count": (i64, [vp, i32]),
    "hgb_model_buffer_bytes": (i64, [vp, i32]),
    "hgb_model_bind": (i32, [vp, i32, vp, i64]),
    "hgb_model_num_tensors": (i32, [vp]),
    "hgb_model_tensor_info": (i32, [vp, i32, C.POINTER(C.c_char_p), C.POINTER(i32), C.POINTER(i64 * 4), C.POINTER(i64),
                                    C.POINTER(i32)]),
    "hgb_model_num_convs": (i32, [vp]),
    "hgb_model_conv_info": (i32, [vp, i32, C.POINTER(C.c_char_p), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                                  C.POINTER(i32), C.POINTER(i32), C.POINTER(f64)]),
    "hgb_model_sync_weights": (i32, [vp, vp]),
    "hgb_model_forward": (i32, [vp, vp, i32, C.POINTER(vp), vp]),
    "hgb_model_loss": (i32, [vp, i32, vp, f64, vp, vp]),
    "hgb_model_num_segments": (i32, [vp]),
    "hgb_model_backward": (i32, [vp, i32, i32, vp]),
    "hgb_model_backward_nojoin": (i32, [vp, i32, i32, vp]),
    "hgb_model_lanes_join": (i32, [vp, vp, i32]),
    "hgb_model_segment_grads": (i32, [vp, i32, C.POINTER(i64), C.POINTER(i64)]),
    "hgb_model_adam_step": (i32, [vp, f64, f64, f64, f64, i64, f64, vp]),
    "hgb_comm_unique_id": (i32, [vp, i32]),
    "hgb_comm_init": (i32, [i32, i32, vp, C.POINTER(vp)]),
    "hgb_comm_destroy": (i32, [vp]),
    "hgb_comm_info": (i32, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
    "hgb_comm_allreduce_f32": (i32, [vp, vp, i64, vp]),
    "hgb_model_set_comm": (i32, [vp, vp, i32]),
    "hgb_grad_allreduce_bucket": (i32, [vp, i32, i32, vp]),
    "hgb_model_conv_output": (i32, [vp, i32, C.POINTER(i64), C.POINTER(i32 * 4)]),
    "hgb_model_conv_input_bn": (i32, [vp, i32]),
    "hgb_model_num_ops": (i32, [vp, i32, i32]),
    "hgb_model_op_info": (i32, [vp, i32, i32, i32, C.POINTER(i32 * 8)]),
    "hgb_model_act_info": (i32, [vp, i32, C.POINTER(i64), C.POINTER(i32 * 4)]),
    "hgb_model_op_fused_bn": (i32, [vp, i32, i32, i32, C.POINTER(i32 * 3)]),
    "hgb_model_run_op": (i32, [vp, i32, i32, i32, vp, i32, vp]),
    "hgb_model_conv_detail": (i32, [vp, i32, C.POINTER(i32 * 8), C.POINTER(i64 * 2)]),
    "hgb_model_dw_detail": (i32, [vp, i32, C.POINTER(i32 * 4), C.POINTER(i64)]),
    "hgb_model_bn_detail": (i32, [vp, i32, C.POINTER(i64 * 8)]),
    "hgb_model_head_buffers": (i32, [vp, i32, C.POINTER(i64 * 2)]),
    "hgb_model_begin_step": (i32, [vp, vp]),
    "hgb_model_profile_conv": (i32, [vp, i32, i32, i32, i32, i32, i32]),
    "hgb_model_profile_read": (i32, [vp, C.POINTER(f64), C.POINTER(i32), C.POINTER(f64)]),
    "hgb_model_profile_all": (i32, [vp, i32]),
    "hgb_model_profile_count": (i32, [vp]),
    "hgb_model_profile_op": (i32, [vp, i32, C.POINTER(i32 * 8), C.POINTER(f64)]),
    "hgb_model_profile_op_fused": (i32, [vp, i32, C.POINTER(i32 * 3)]),
    "hgb_model_sched_count": (i32, [vp, i32]),
    "hgb_model_sched_op": (i32, [vp, i32, i32, C.POINTER(i32 * 4), C.POINTER(i32), C.POINTER(i32 * 16)]),
    "hgb_model_sched_access": (i32, [vp, i32, i32, i32, C.POINTER(i64)]),
    "hgb_model_launch_count": (i64, [vp]),
}

MISSING = []
for _name, (_res, _args) in PROTOTYPES.items():
    try:
        _fn = getattr(lib, _name)
    except AttributeError:
        MISSING.append(_name)
        continue
    _fn.restype = _res
    _fn.argtypes = _args


class HgbError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = lib.hgb_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise HgbError(f"libhgb200 error {rc}: {msg}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise HgbError("libhgb200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
