"""ctypes binding of the C ABI declared in include/lerf_b200.h.

There is deliberately no fallback: if liblerf_b200.so is missing or a call fails, the caller gets
an exception.  Nothing here imports oracle/.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# LERF_B200_EXPERIMENTS=1 (scripts/kbench.py, A/B runs) loads the build with every tuning variant compiled in
# (`python -m lerf_pytorch_b200.build --experiments`); everything else uses the product library.
EXPERIMENTS = os.environ.get("LERF_B200_EXPERIMENTS", "") not in ("", "0")
LIB_PATH = os.path.join(HERE, "liblerf_b200_exp.so" if EXPERIMENTS else "liblerf_b200.so")

LERF_KIND_GAUSS, LERF_KIND_LINEAR = 0, 1
LERF_OUT_F32, LERF_OUT_U8, LERF_OUT_U8_HWC = 0, 1, 2
LERF_WARP_NEAREST, LERF_WARP_BILINEAR, LERF_WARP_BICUBIC, LERF_WARP_LANCZOS2, LERF_WARP_LANCZOS3 = 0, 1, 2, 3, 4

_c_i, _c_ll, _c_f, _c_p, _c_sz = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol of include/lerf_b200.h (tests check it)
PROTOTYPES = {
    "lerf_abi_version": (_c_i, []),
    "lerf_last_error_string": (ctypes.c_char_p, []),
    "lerf_luts_create": (_c_i, [_c_p, _c_i, _c_i, _c_p]),
    "lerf_luts_destroy": (None, [_c_p]),
    "lerf_luts_oc": (_c_i, [_c_p]),
    "lerf_luts_pin_l2": (_c_i, [_c_p, _c_p]),
    "lerf_lut_pass": (_c_i, [_c_p, _c_p, _c_i, _c_i, _c_i, ctypes.c_char, _c_i, _c_p, _c_p]),
    "lerf_lut_stage1": (_c_i, [_c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_ll, _c_ll, _c_ll, _c_ll, _c_i, _c_i, _c_p, _c_p]),
    "lerf_lut_stage2": (_c_i, [_c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p, _c_p]),
    "lerf_sr_plan_create": (_c_i, [_c_i, _c_i, _c_i, _c_i, _c_p, _c_p, _c_p, _c_p, _c_i, _c_p]),
    "lerf_sr_plan_create_ex": (_c_i, [_c_i, _c_i, _c_i, _c_i, _c_i, _c_p, _c_p, _c_p, _c_p, _c_i, ctypes.c_double, _c_i, _c_p]),
    "lerf_sr_plan_destroy": (None, [_c_p]),
    "lerf_resize_sr": (_c_i, [_c_i, _c_p, _c_p, _c_p, _c_i, _c_i, _c_f, _c_i, _c_i, _c_p, _c_i, _c_p]),
    "lerf_resize_sr_f32": (_c_i, [_c_i, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_f, _c_p, _c_p]),
    "lerf_warp": (_c_i, [_c_i, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p, _c_i, _c_i, _c_f, _c_p, _c_i,
                         _c_p, _c_i, _c_i, _c_i, _c_p]),
    "lerf_warp_f32": (_c_i, [_c_i, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p, _c_i, _c_i, _c_f,
                             _c_p, _c_p]),
    "lerf_warp_ex": (_c_i, [_c_i, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p, _c_i, _c_i, _c_i, _c_i,
                            _c_f, _c_p, _c_i, _c_p]),
    "lerf_warp_fixed_support": (_c_i, [_c_i]),
    "lerf_warp_fixed": (_c_i, [_c_i, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p, _c_i, _c_i, _c_p, _c_p]),
    "lerf_sr_scratch_bytes": (_c_sz, [_c_i, _c_i, _c_i, _c_i]),
    "lerf_sr_fused": (_c_i, [_c_p, _c_i, _c_p, _c_p, _c_i, _c_i, _c_ll, _c_ll, _c_ll, _c_ll, _c_f, _c_i, _c_i, _c_p,
                             _c_p, _c_i, _c_p]),
    "lerf_png_stored_bytes": (_c_ll, [_c_i, _c_i, _c_i]),
    "lerf_png_encode_stored": (_c_i, [_c_p, _c_i, _c_i, _c_i, _c_p, _c_ll, _c_p, _c_p]),
    "lerf_lut_ft_forward": (_c_i, [_c_p, _c_i, _c_p, _c_i, _c_i, _c_i, ctypes.c_char, _c_i, _c_p, _c_p]),
    "lerf_lut_ft_backward": (_c_i, [_c_p, _c_i, _c_p, _c_i, _c_i, _c_i, ctypes.c_char, _c_i, _c_p, _c_p]),
    "lerf_resize_sr_f32_backward": (_c_i, [_c_i, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_f, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "lerf_launch_count": (_c_ll, []),
    "lerf_launch_count_reset": (None, []),
}

# include/lerf_b200_testing.h: test / tuning switches, not part of the drop-in interface
TESTING_PROTOTYPES = {
    "lerf_build_has_experiments": (_c_i, []),
    "lerf_debug_lut_variant": (None, [_c_i, _c_i]),
    "lerf_debug_cell_hash": (None, [_c_i, _c_i, _c_i]),
    "lerf_debug_force_generic": (None, [_c_i]),
    "lerf_debug_warp_records": (None, [_c_i]),
    "lerf_debug_resize_variant": (None, [_c_i]),
    "lerf_debug_pipeline": (None, [_c_i, _c_i, _c_i]),
    "lerf_debug_l2_window": (None, [_c_i]),
    "lerf_debug_carveout": (None, [_c_i]),
}

_lib = None


class LerfError(RuntimeError):
    pass


def lib():
    """Load liblerf_b200.so (once).  Raises if it has not been built -- there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "lerf_pytorch_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in list(PROTOTYPES.items()) + list(TESTING_PROTOTYPES.items()):
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.lerf_abi_version() != 1:
            raise ImportError("lerf_pytorch_b200: ABI version mismatch, rebuild liblerf_b200.so")
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().lerf_last_error_string().decode("utf-8", "replace")
        if rc == 1:
            raise ValueError(msg)  # the reference raises ValueError for bad modes / arguments
        raise LerfError("lerf_b200 error %d: %s" % (rc, msg))
