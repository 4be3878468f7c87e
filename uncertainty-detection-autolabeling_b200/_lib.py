"""ctypes binding of libudal.so (C ABI: include/udal.h).

The shared library is hand-written CUDA for sm_100a.  There is NO fallback: if it is missing
or cannot be loaded, importing the product fails with an explicit error.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libudal.so")

MAX_LEVELS = 8
ABI_VERSION = 1

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NOMEM = 0, -1, -2, -3, -4
DECODE_METHODS = {"l-norm": 0, "n-flow": 1, "falsedec": 2}
NMS_HARD, NMS_GAUSSIAN = 0, 1
HEADS_FP32, HEADS_BF16_TC, HEADS_FP16_TC, HEADS_FP32X3_TC = 0, 1, 2, 3
FEAT_F32, FEAT_F16 = 0, 1
STAGE_SLOTS = 4
FUSE_SUM, FUSE_FASTATTN, FUSE_ATTN = 0, 1, 2
CLASSCAL = {"ts_all": 0, "ts_percls": 1, "iso_all": 2, "iso_percls": 3}
ACT_NONE, ACT_BN_SWISH, ACT_BN = 0, 1, 2
HEAD_CLASS, HEAD_BOX = 0, 1


class Config(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32), ("device", ctypes.c_int32),
        ("image_h", ctypes.c_int32), ("image_w", ctypes.c_int32),
        ("num_levels", ctypes.c_int32),
        ("level_h", ctypes.c_int32 * MAX_LEVELS), ("level_w", ctypes.c_int32 * MAX_LEVELS),
        ("anchors_per_loc", ctypes.c_int32), ("num_classes", ctypes.c_int32),
        ("num_filters", ctypes.c_int32), ("repeats", ctypes.c_int32),
        ("mc_samples", ctypes.c_int32), ("loss_attenuation", ctypes.c_int32),
        ("cls_mc", ctypes.c_int32), ("box_mc", ctypes.c_int32),
        ("rate_class", ctypes.c_float), ("rate_box", ctypes.c_float),
        ("decode_method", ctypes.c_int32), ("nms_method", ctypes.c_int32),
        ("nms_iou_thresh", ctypes.c_float), ("nms_score_thresh", ctypes.c_float),
        ("nms_sigma_tf", ctypes.c_float), ("nms_variant_old", ctypes.c_int32),
        ("max_nms_inputs", ctypes.c_int32), ("max_output_size", ctypes.c_int32),
        ("heads_mode", ctypes.c_int32), ("prefilter_k", ctypes.c_int32),
        ("inv_keep_class", ctypes.c_float), ("inv_keep_box", ctypes.c_float),
        ("decode_precision", ctypes.c_int32), ("reserved", ctypes.c_int32 * 4),
    ]


class PreNmsOut(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("mean_logits", "std_logits", "boxes", "albox", "mcbox", "scores", "classes")]


class PreNmsTopkOut(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("mean_logits", "topk_idx", "boxes", "albox", "mcbox", "mcclass", "scores", "classes")]


class Detections(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("boxes", "scores", "classes", "valid", "logits")]


class UdalError(RuntimeError):
    pass


_VP = ctypes.c_void_p
_PP = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); mirrors include/udal.h one to one
SIGNATURES = {
    "udal_last_error": (ctypes.c_char_p, []),
    "udal_abi_version": (ctypes.c_int, []),
    "udal_create": (ctypes.c_int, [ctypes.POINTER(Config), _PP]),
    "udal_destroy": (ctypes.c_int, [_VP]),
    "udal_set_stream": (ctypes.c_int, [_VP, _VP]),
    "udal_sync": (ctypes.c_int, [_VP]),
    "udal_set_feature_format": (ctypes.c_int, [_VP, ctypes.c_int]),
    "udal_conv1x1_bn": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP, _VP, _VP, _VP,
                                       ctypes.c_int, _VP]),
    "udal_bifpn_fuse": (ctypes.c_int, [_VP, ctypes.c_int, _VP, _VP, _VP, _VP, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP]),
    "udal_sepconv_bn": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP, _VP,
                                       _VP, _VP, _VP, ctypes.c_int, _VP]),
    "udal_sepconv_tc_prepare": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, ctypes.POINTER(ctypes.c_void_p)]),
    "udal_sepconv_tc": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, ctypes.c_int, _VP, _VP, ctypes.c_int, _VP]),
    "udal_calibrate_class": (ctypes.c_int, [_VP, _VP, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, _VP, _VP, _VP, _VP, _VP, _VP]),
    "udal_stage_begin": (ctypes.c_int, [_VP, ctypes.c_int]),
    "udal_stage_h2d": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_size_t]),
    "udal_stage_end": (ctypes.c_int, [_VP, ctypes.c_int]),
    "udal_stage_acquire": (ctypes.c_int, [_VP, ctypes.c_int]),
    "udal_stage_release": (ctypes.c_int, [_VP, ctypes.c_int]),
    "udal_fetch_d2h": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_size_t]),
    "udal_fetch_mark": (ctypes.c_int, [_VP, ctypes.c_int]),
    "udal_fetch_wait": (ctypes.c_int, [_VP, ctypes.c_int]),
    "udal_get_stream": (ctypes.c_int, [_VP, _PP]),
    "udal_wait_stream": (ctypes.c_int, [_VP, _VP]),
    "udal_wait_context": (ctypes.c_int, [_VP, _VP]),
    "udal_malloc": (ctypes.c_int, [_VP, ctypes.c_size_t, _PP]),
    "udal_free": (ctypes.c_int, [_VP, _VP]),
    "udal_host_alloc": (ctypes.c_int, [ctypes.c_size_t, _PP]),
    "udal_host_free": (ctypes.c_int, [_VP]),
    "udal_memcpy_h2d": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_size_t]),
    "udal_memcpy_d2h": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_size_t]),
    "udal_memcpy_d2d": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_size_t]),
    "udal_memset": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_size_t]),
    "udal_timer_start": (ctypes.c_int, [_VP]),
    "udal_timer_stop": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_float)]),
    "udal_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "udal_num_anchors": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_int64)]),
    "udal_launch_count": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_int64)]),
    "udal_profile_layers": (ctypes.c_int, [_VP, ctypes.c_int]),
    "udal_get_layer_times": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_float), ctypes.c_int,
                                            ctypes.POINTER(ctypes.c_int)]),
    "udal_scratch_bytes": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_size_t)]),
    "udal_set_anchors": (ctypes.c_int, [_VP, _VP, ctypes.c_int64]),
    "udal_set_head_weights": (ctypes.c_int, [_VP, ctypes.c_int] + [_VP] * 10),
    "udal_heads_sample": (ctypes.c_int, [_VP, _PP, ctypes.c_int, _VP, ctypes.c_uint64, _PP, _PP]),
    "udal_decode_moments": (ctypes.c_int, [_VP, _PP, _PP, ctypes.c_int, ctypes.POINTER(PreNmsOut)]),
    "udal_topk": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int64, ctypes.c_int, _VP, _VP]),
    "udal_prenms_topk": (ctypes.c_int, [_VP, _PP, _PP, ctypes.c_int, ctypes.POINTER(PreNmsTopkOut)]),
    "udal_nms_v5": (ctypes.c_int, [_VP, _VP, _VP, ctypes.c_int, ctypes.c_int, _VP, _VP, _VP]),
    "udal_postprocess_global": (ctypes.c_int, [_VP, _PP, _PP, ctypes.c_int, _VP,
                                               ctypes.POINTER(Detections)]),
    "udal_postprocess_per_class": (ctypes.c_int, [_VP, _PP, _PP, ctypes.c_int, _VP, ctypes.c_int,
                                                  ctypes.POINTER(Detections)]),
    "udal_per_class_nms": (ctypes.c_int, [_VP, _VP, _VP, _VP, ctypes.c_int, ctypes.c_int, _VP, _VP,
                                          ctypes.c_int64, ctypes.c_int, ctypes.POINTER(Detections)]),
    "udal_format_detections": (ctypes.c_int, [_VP, _VP, ctypes.c_int, _VP, _VP, ctypes.c_int, _VP, _VP,
                                              ctypes.c_int, _VP, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, _VP]),
    "udal_transform_detections": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, ctypes.c_int, _VP]),
    "udal_concat_channels": (ctypes.c_int, [_VP, _VP, ctypes.c_int, _VP, ctypes.c_int,
                                            ctypes.c_int64, _VP]),
    "udal_gather_rows": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int64, ctypes.c_int, _VP,
                                        ctypes.c_int, ctypes.c_int, _VP]),
    "udal_run": (ctypes.c_int, [_VP, _PP, ctypes.c_int, _VP, ctypes.c_uint64, _VP,
                                ctypes.POINTER(Detections)]),
    "udal_clip_boxes": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, ctypes.c_float, ctypes.c_float, _VP]),
    "udal_max_reduce": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, ctypes.c_int, _VP, _VP]),
    "udal_divmod_i32": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, ctypes.c_int, _VP, _VP]),
    "udal_sigmoid": (ctypes.c_int, [_VP, _VP, ctypes.c_int64, _VP]),
    "udal_decode_sample": (ctypes.c_int, [_VP, _VP, _VP, _VP, ctypes.c_int64, ctypes.c_int, _VP, ctypes.c_uint64, _VP, _VP]),
    "udal_run_prenms": (ctypes.c_int, [_VP, _PP, ctypes.c_int, _VP, ctypes.c_uint64, ctypes.POINTER(PreNmsOut)]),
    "udal_nms_np": (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                   ctypes.c_float, ctypes.c_float, _VP,
                                   ctypes.POINTER(ctypes.c_int32)]),
}



class AutolabelParams(ctypes.Structure):
    """include/udal.h: udal_autolabel_params"""
    _fields_ = [("calib_method_box", ctypes.c_int32), ("num_tables", ctypes.c_int32),
                ("table_x", ctypes.c_void_p), ("table_y", ctypes.c_void_p), ("table_off", ctypes.c_void_p),
                ("temps", ctypes.c_float * 4), ("class_temp", ctypes.c_float), ("w_entropy", ctypes.c_float),
                ("w_albox", ctypes.c_float), ("threshold", ctypes.c_float), ("min_score", ctypes.c_float),
                ("strict_reference", ctypes.c_int32)]


CALIB_METHODS = {None: 0, "none": 0, "ts_all": 1, "ts_percoo": 2, "iso_all": 3, "iso_percoo": 4,
                 "iso_perclscoo": 5, "rel_iso_perclscoo": 6}
SIGNATURES["udal_autolabel"] = (ctypes.c_int, [_VP, _VP, ctypes.c_int, ctypes.c_int, _VP, _VP, ctypes.c_int, _VP,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.POINTER(AutolabelParams), _VP, _VP, _VP, _VP, _VP])

_lib = None


def load():
    """Loads libudal.so (binding every symbol of include/udal.h) or raises - never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libudal.so is missing (%s). Build it with `python "
            "uncertainty-detection-autolabeling_b200/build.py`; this package has no CPU fallback."
            % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if lib.udal_abi_version() != ABI_VERSION:
        raise ImportError("libudal.so ABI %d != binding ABI %d" % (lib.udal_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(status):
    """Maps a udal_status to the exception the reference's Python API would raise."""
    if status == OK:
        return
    msg = load().udal_last_error().decode("utf-8", "replace")
    if status == ERR_INVALID:
        raise ValueError(msg)
    if status == ERR_NOMEM:
        raise MemoryError(msg)
    raise UdalError(msg)


def device_count():
    n = ctypes.c_int(0)
    lib = load()
    st = lib.udal_device_count(ctypes.byref(n))
    if st != OK:
        return 0
    return n.value
