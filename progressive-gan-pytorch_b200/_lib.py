"""ctypes binding of libprogan_b200.so — the C-ABI declared in include/progan_b200.h.

Plays the role of the reference's plugin loader (ada/torch_utils/custom_ops.py:46-124,
`get_plugin`): load once, cache, raise on failure.  Unlike the reference there is no
fallback implementation: if the library is missing or a call fails, a RuntimeError is
raised (north_star: "no CPU fallback").
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libprogan_b200.so")

c_int, c_ll, c_float, c_void_p = ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_void_p
P = c_void_p

# name -> argtypes, exactly the prototypes of include/progan_b200.h
SIGNATURES = {
    "pg_abi_version": [],
    "pg_device_info": [ctypes.POINTER(c_int)] * 3,
    "pg_pack_conv_weight": [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "pg_pack_conv_weight_multi": [P, c_int, P],
    "pg_wgrad_unpack_multi": [P, c_int, P],
    "pg_conv_fwd_simt": [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                         c_int, c_float, c_int, P],
    "pg_conv_wgrad_simt": [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                           c_int, c_int, c_int, P],
    "pg_conv_tc": [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                   c_int, c_float, P, P],
    "pg_conv_tc_actbwd": [P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P, P, c_float, c_int, P, P],
    "pg_conv_wgrad_tc": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                         c_float, c_int, c_int, c_int, P],
    "pg_pn_lrelu_fwd": [P, P, P, c_ll, c_int, c_float, c_int, c_int, P],
    "pg_pn_lrelu_bwd": [P, P, P, P, c_ll, c_int, c_float, c_int, c_int, c_int, P, P, c_int, P],
    "pg_pn_lrelu_bwd_bwd": [P, P, P, P, P, P, c_ll, c_int, c_float, c_int, c_int, c_int, c_int, P],
    "pg_colsum": [P, P, c_ll, c_int, c_int, P],
    "pg_pw_expand": [P, P, P, P, c_int, c_ll, c_int, c_int, c_int, c_int, c_float, c_int, P],
    "pg_pw_reduce": [P, P, P, P, c_int, c_ll, c_int, c_int, c_int, c_int, c_float, c_int, P],
    "pg_pw_wgrad": [P, P, P, P, c_int, c_ll, c_int, c_int, c_int, c_int, c_float, c_int, P],
    "pg_img_chansum": [P, P, c_int, c_ll, c_int, P],
    "pg_avgpool2": [P, P, c_int, c_int, c_int, c_int, c_int, P],
    "pg_avgpool2_bwd": [P, P, c_int, c_int, c_int, c_int, c_int, P],
    "pg_upsample2": [P, P, c_int, c_int, c_int, c_int, c_int, P],
    "pg_upsample2_bwd": [P, P, c_int, c_int, c_int, c_int, c_int, P],
    "pg_blend": [P, P, P, c_ll, P, c_int, P],
    "pg_scale": [P, P, c_ll, c_float, c_float, P, c_int, P],
    "pg_tanh_fwd": [P, P, c_ll, P],
    "pg_tanh_bwd": [P, P, P, c_ll, P],
    "pg_mbstd_fwd": [P, P, P, c_int, c_int, c_int, c_int, P],
    "pg_mbstd_bwd": [P, P, P, P, c_int, c_int, c_int, c_int, P],
    "pg_mbstd_bwd_bwd": [P, P, P, P, P, P, c_int, c_int, c_int, c_int, P],
    "pg_wgan_loss": [P, P, P, c_int, c_int, c_float, P],
    "pg_interp_xhat": [P, P, P, P, c_int, c_ll, P],
    "pg_gp_fwd": [P, P, P, c_int, c_ll, c_float, P],
    "pg_gp_bwd": [P, P, P, P, c_int, c_ll, c_float, P],
    "pg_adam_step": [P, P, P, P, c_ll, c_float, c_float, c_float, c_float, P, c_float, P],
    "pg_adam_multi": [P, P, P, P, P, c_int, P, c_float, c_float, c_float, c_float, c_float, P],
    "pg_ema": [P, P, c_ll, c_float, P],
}



class PackEntry(ctypes.Structure):
    """PgPackEntry of include/progan_b200.h."""
    _fields_ = [("w", c_void_p), ("out", c_void_p), ("total", c_ll)] + [
        (n, c_int) for n in ("d0", "d1", "taps", "swap_io", "flip", "layout", "ci_pad", "co_pad",
                             "dtype", "reserved")]


class UnpackEntry(ctypes.Structure):
    """PgUnpackEntry of include/progan_b200.h."""
    _fields_ = [("ws", c_void_p), ("dw", c_void_p)] + [
        (n, c_int) for n in ("Cin", "Cout", "Cin_p", "Cout_p", "taps", "swap_io", "flip", "reserved")
    ] + [("scale", c_float), ("reserved2", c_float)]


_lib = None
_lock = threading.Lock()


def load():
    """Load (once) and return the ctypes handle; raise RuntimeError if unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "progan_b200: %s not found. Build it with `python __graft_entry__.py` "
                "(or progressive-gan-pytorch_b200/build.py); there is no fallback path."
                % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        lib.pg_last_error.restype = ctypes.c_char_p
        lib.pg_last_error.argtypes = []
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = c_int
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc, name):
    if rc != 0:
        msg = _lib.pg_last_error().decode("utf-8", "replace") if _lib is not None else ""
        raise RuntimeError("progan_b200: %s failed (%d): %s" % (name, rc, msg))
