"""ctypes binding of libtcs_b200.so — the C-ABI declared in include/tcs_b200.h.

Nothing here computes: it loads the library, declares every entry point's signature, and turns a non-zero
status into a RuntimeError carrying tcs_last_error().  There is no fallback of any kind: if the library is
missing, the import of the package's compute modules fails.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtcs_b200.so")

PREC_BF16, PREC_BF16X3, PREC_FP16, PREC_FP16X3 = 0, 1, 2, 3
PRECISIONS = {"bf16": PREC_BF16, "bf16x3": PREC_BF16X3, "fp16": PREC_FP16, "fp16x3": PREC_FP16X3}
MAX_LEVELS = 4
MAX_RADIUS = 8
ABI_VERSION = 10

_p = ctypes.c_void_p
_i = ctypes.c_int
_ll = ctypes.c_longlong
_f = ctypes.c_float

# name -> (restype, argtypes); one entry per declaration in include/tcs_b200.h
SIGNATURES = {
    "tcs_abi_version": (_i, []),
    "tcs_last_error": (ctypes.c_char_p, []),
    "tcs_corr_prepass": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "tcs_corr_prepass_kblocked": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "tcs_corr_build": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "tcs_corr_build_fused": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "tcs_corr_build_fp32": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "tcs_corr_lookup": (_i, [_p, _p, _p, _p, _p, _ll, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "tcs_corr_lookup_encode": (_i, [_p, _p, _p, _p, _p, _ll, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "tcs_corr_encode_packed_bytes": (_i, []),
    "tcs_corr_encode_pack_weights": (_i, [_p, _p, _p, _p]),
    "tcs_corr_lookup_encode_tc": (_i, [_p, _p, _p, _p, _p, _ll, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "tcs_corr_lookup_backward": (_i, [_p, _p, _ll, _p, _i, _i, _i, _i, _i, _i, _p]),
    "tcs_init_loss_forward": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "tcs_init_loss_backward": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "tcs_fmap_pool_w": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "tcs_corr_lookup_alt": (_i, [_p, _p, _p, _p, _p, _p, _ll, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "tcs_corr_lookup_alt_tc": (_i, [_p, _p, _p, _p, _p, _ll, _p, _i, _i, _i, _i, _i, _i, _p]),
    "tcs_corr_argmax": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "tcs_corr_cost_volume": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "tcs_warp_scratch_bytes": (_ll, [_i, _i, _i, _i]),
    "tcs_warp_forward": (_i, [_p] * 14 + [_i, _i, _i, _i, _i, _p]),
    "tcs_relative_pose": (_i, [_p, _p, _p, _i, _p]),
    "tcs_backward_grid": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "tcs_bilinear_sample": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "tcs_grid_halve": (_i, [_p, _p, _i, _i, _i, _p]),
    "tcs_warp_hidden_states": (_i, [_p] * 7 + [_i] * 6 + [_p]),
    "tcs_disp_gradient_xy": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "tcs_disp_grad_candidates": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "tcs_disp_propagate": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "tcs_completor_stems_weight_floats": (_i, []),
    "tcs_completor_stems": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "tcs_convex_upsample": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
}

_lib = None


class TcsError(RuntimeError):
    """A libtcs_b200 entry point returned a non-zero status."""


def load():
    """Load libtcs_b200.so once and declare its signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libtcs_b200.so is not built (expected at %s). Run `python __graft_entry__.py build` or "
            "`python temporally-consistent-stereo-matching_b200/build.py`; there is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    got = lib.tcs_abi_version()
    if got != ABI_VERSION:
        raise ImportError("libtcs_b200.so has ABI version %d, this package needs %d; rebuild it" % (got, ABI_VERSION))
    _lib = lib
    return lib


def call(name, *args):
    """Call an int-status entry point; raise TcsError with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.tcs_last_error()
        kind = "argument" if rc < 0 else "CUDA"
        raise TcsError("%s failed (%s error %d): %s" % (name, kind, rc, msg.decode() if msg else "?"))


def warp_scratch_bytes(B, C, H, W):
    return int(load().tcs_warp_scratch_bytes(B, C, H, W))
