"""ctypes binding of libesm_b200.so (the C ABI declared in include/esm_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ESM_LIB") or os.path.join(_HERE, "csrc", "libesm_b200_prof.so" if os.environ.get("ESM_TC_PROFILE") == "1" else "libesm_b200.so")  # ESM_LIB: an A/B build of the same ABI (diagnostics)

ACT = {None: 0, "none": 0, "gelu": 1, "relu": 2, "silu": 3, "sigmoid": 4, "2sigmoid": 5, "relu6": 6}
SRC_TENSORS, SRC_GWC = 0, 1

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int)
vp = C.c_void_p


class EsmSrc(C.Structure):
    _fields_ = [("ptr", vp), ("C", C.c_int), ("sB", C.c_longlong), ("sC", C.c_longlong),
                ("sD", C.c_longlong), ("sH", C.c_longlong)]


class EsmConv(C.Structure):
    _fields_ = [
        ("src", EsmSrc * 3), ("nsrc", C.c_int), ("src_mode", C.c_int), ("gwc_groups", C.c_int),
        ("in_mul", vp),
        ("B", C.c_int), ("Cin", C.c_int), ("Din", C.c_int), ("Hin", C.c_int), ("Win", C.c_int),
        ("Cout", C.c_int), ("Dout", C.c_int), ("Hout", C.c_int), ("Wout", C.c_int),
        ("kd", C.c_int), ("kh", C.c_int), ("kw", C.c_int), ("stride", C.c_int),
        ("pd", C.c_int), ("ph", C.c_int), ("pw", C.c_int), ("transposed", C.c_int),
        ("weight", vp), ("scale", vp), ("shift", vp), ("act", C.c_int),
        ("out_mul", vp), ("residual", vp), ("act2", C.c_int), ("out_scale", C.c_float),
        ("pixel_shuffle", C.c_int), ("out", vp),
        ("oB", C.c_longlong), ("oC", C.c_longlong), ("oD", C.c_longlong), ("oH", C.c_longlong),
        ("engine", C.c_int),
    ]


class EsmMixerMlp(C.Structure):
    _fields_ = [("ln_w", vp), ("fc0_w", vp), ("fc0_b", vp), ("fc2_w", vp), ("fc2_b", vp), ("hidden", C.c_int)]


class EsmPf(C.Structure):
    _fields_ = [("data", vp), ("B", C.c_int), ("C", C.c_int), ("Dp", C.c_int), ("Hp", C.c_int), ("P", C.c_int),
                ("d0", C.c_int), ("d1", C.c_int), ("y0", C.c_int), ("y1", C.c_int), ("x0", C.c_int), ("x1", C.c_int)]


class EsmConvPf(C.Structure):
    _fields_ = [
        ("src", EsmPf * 3), ("nsrc", C.c_int),
        ("Cout", C.c_int), ("kd", C.c_int), ("kh", C.c_int), ("kw", C.c_int), ("stride", C.c_int), ("transposed", C.c_int),
        ("d0", C.c_int), ("d1", C.c_int), ("y0", C.c_int), ("y1", C.c_int), ("x0", C.c_int), ("x1", C.c_int),
        ("oD", C.c_int), ("oH", C.c_int), ("oW", C.c_int),
        ("weight", vp), ("scale", vp), ("shift", vp), ("act", C.c_int), ("act2", C.c_int), ("out_scale", C.c_float),
        ("pixel_shuffle", C.c_int), ("out_pf", EsmPf), ("res_pf", vp), ("out", vp),
        ("oB", C.c_longlong), ("oC", C.c_longlong), ("oDs", C.c_longlong), ("oHs", C.c_longlong), ("residual", vp),
    ]


# name -> (restype, argtypes); every symbol include/esm_b200.h declares
SIGNATURES = {
    "esm_last_error": (C.c_char_p, []),
    "esm_version": (C.c_int, []),
    "esm_set_pdl": (C.c_int, [C.c_int]),
    "esm_device_info": (C.c_int, [i32p, i32p, i32p]),
    "esm_packed_weight_elems": (C.c_longlong, [C.c_int] * 6),
    "esm_pack_conv_weight_f32": (C.c_int, [vp, vp] + [C.c_int] * 6 + [vp]),
    "esm_fold_bn_f32": (C.c_int, [vp, vp, vp, vp, vp, C.c_float, C.c_int, vp, vp, vp]),
    "esm_conv_f32": (C.c_int, [C.POINTER(EsmConv), vp]),
    "esm_conv_plans_export": (C.c_longlong, [C.c_char_p, C.c_longlong]),
    "esm_conv_plans_import": (C.c_int, [C.c_char_p]),
    "esm_conv_tuned_calls": (C.c_longlong, []),
    "esm_tc_conv_launches": (C.c_longlong, []),
    "esm_tcg_conv_launches": (C.c_longlong, []),
    "esm_pw_conv_launches": (C.c_longlong, []),
    "esm_gwc_volume_f32": (C.c_int, [vp, vp, vp] + [C.c_int] * 6 + [vp]),
    "esm_norm_corr_volume_f32": (C.c_int, [vp, vp, vp, vp] + [C.c_int] * 5 + [vp]),
    "esm_concat_volume_f32": (C.c_int, [vp, vp, vp] + [C.c_int] * 5 + [vp]),
    "esm_substract_volume_f32": (C.c_int, [vp, vp, vp] + [C.c_int] * 6 + [vp]),
    "esm_gwc_volume_norm_f32": (C.c_int, [vp, vp, vp] + [C.c_int] * 6 + [vp]),
    "esm_preprocess_u8_f32": (C.c_int, [vp, vp] + [C.c_int] * 8 + [C.POINTER(C.c_float), C.POINTER(C.c_float), vp]),
    "esm_postprocess_disp_u16": (C.c_int, [vp, vp] + [C.c_int] * 7 + [C.c_float, vp]),
    "esm_regression_top2_f32": (C.c_int, [vp, vp, vp] + [C.c_int] * 4 + [vp]),
    "esm_regression_top2_subpixel_f32": (C.c_int, [vp] + [C.c_longlong] * 4 + [vp, vp] + [C.c_int] * 4 + [vp]),
    "esm_pixel_shuffle3d_f32": (C.c_int, [vp] + [C.c_longlong] * 4 + [vp] + [C.c_int] * 4 + [vp]),
    "esm_copy_f32": (C.c_int, [vp, vp, C.c_longlong, vp]),
    "esm_disparity_publish_u16": (C.c_int, [vp, vp] + [C.c_int] * 4 + [C.c_float, C.c_float, vp]),
    "esm_disparity_regression_f32": (C.c_int, [vp, vp] + [C.c_int] * 4 + [vp]),
    "esm_bilinear_add_f32": (C.c_int, [vp, vp, vp] + [C.c_int] * 4 + [C.c_float, vp]),
    "esm_sm_pointwise_f32": (C.c_int, [vp, vp] + [C.c_int] * 4 + [C.POINTER(EsmMixerMlp), vp, vp]),
    "esm_sm_spatial_f32": (C.c_int, [vp, vp] + [C.c_int] * 4 + [vp, vp, C.c_int, C.POINTER(EsmMixerMlp), vp, vp]),
    "esm_sm_layer_f32": (C.c_int, [vp, vp] + [C.c_int] * 4 + [C.POINTER(EsmMixerMlp), vp, vp, C.c_int, C.POINTER(EsmMixerMlp), vp, vp]),
    "esm_laf_cost_top7_f32": (C.c_int, [vp, vp] + [C.c_int] * 4 + [vp]),
    "esm_laf_attention_f32": (C.c_int, [vp] * 7 + [C.c_int] * 4 + [vp]),
    "esm_laf_sample_embed_f32": (C.c_int, [vp] * 8 + [C.c_int] * 4 + [vp]),
    "esm_conf_convex_up4_f32": (C.c_int, [vp] * 5 + [C.c_int] * 4 + [vp]),
    "esm_fill_f32": (C.c_int, [vp, C.c_longlong, C.c_float, vp]),
    "esm_download": (C.c_int, [vp, vp, C.c_longlong]),
    "esm_dwconv2d_f32": (C.c_int, [vp, vp, vp, vp, C.c_int, vp] + [C.c_int] * 6 + [vp]),
    "esm_global_avgpool_f32": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "esm_scale_channels_f32": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "esm_umma_tf32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_float), vp]),
    "esm_pf_elems": (C.c_longlong, [C.c_int] * 5),
    "esm_pf_guard_elems": (C.c_longlong, [C.c_int] * 3),
    "esm_pf_from_nchw_f32": (C.c_int, [vp] + [C.c_longlong] * 4 + [C.POINTER(EsmPf), vp]),
    "esm_pf_to_nchw_f32": (C.c_int, [C.POINTER(EsmPf), vp] + [C.c_longlong] * 4 + [vp]),
    "esm_packed_weight_pf_elems": (C.c_longlong, [C.c_int, C.c_int, i32p] + [C.c_int] * 4),
    "esm_pack_conv_weight_pf_f32": (C.c_int, [vp, vp, C.c_int, C.c_int, i32p] + [C.c_int] * 4 + [vp]),
    "esm_conv_pf_f32": (C.c_int, [C.POINTER(EsmConvPf), vp]),
    "esm_tcf_conv_launches": (C.c_longlong, []),
}

_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "esmstereo_b200: %s is missing -- build it with `python -m esmstereo_b200.build` "
                "(there is no CPU or PyTorch fallback for the hot path)" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
        _import_plans(handle)
    return _lib


PLANS_DIR = os.path.join(_HERE, "plans")


def _import_plans(handle) -> int:
    """Pin the engine plans shipped in esmstereo_b200/plans/*.txt (tuned once on a B200 by scripts/tune_plans.py): the
    conv engines then never time candidates for those shapes, so engine choice -- and rounding -- is the same in every
    process.  ESM_PLANS=0 skips this (every shape is then timed on first use)."""
    if os.environ.get("ESM_PLANS", "1") == "0" or not os.path.isdir(PLANS_DIR):
        return 0
    n = 0
    for name in sorted(os.listdir(PLANS_DIR)):
        if name.endswith(".txt"):
            with open(os.path.join(PLANS_DIR, name), "rb") as f:
                n += handle.esm_conv_plans_import(f.read())
    return n


def export_plans() -> str:
    """The plans this process has tuned or imported, in the text form of esmstereo_b200/plans/*.txt."""
    h = lib()
    n = h.esm_conv_plans_export(None, 0)
    buf = C.create_string_buffer(int(n))
    h.esm_conv_plans_export(buf, n)
    return buf.value.decode()


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().esm_last_error()
        raise RuntimeError("esm_b200 %s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))
