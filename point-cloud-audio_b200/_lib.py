"""ctypes binding of libpcaudio_b200.so (C ABI declared in include/pcaudio_b200.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PCAUDIO_B200_LIB points experiments at an alternative build of the same library (never at a fallback)
LIB_PATH = os.environ.get("PCAUDIO_B200_LIB") or os.path.join(_HERE, "csrc", "libpcaudio_b200.so")

PREC_FP32 = 0
PREC_BF16 = 2


class StDims(C.Structure):
    _fields_ = [("d_in", C.c_int), ("D", C.c_int), ("H", C.c_int), ("M", C.c_int),
                ("S", C.c_int), ("C", C.c_int), ("ln", C.c_int)]


class PipelineCfg(C.Structure):
    _fields_ = [("n_samples", C.c_int), ("n_fft", C.c_int), ("hop", C.c_int), ("scale", C.c_float),
                ("mode", C.c_int), ("ntemp", C.c_int), ("top_k", C.c_int), ("precision", C.c_int),
                ("st", StDims), ("use_threshold", C.c_int), ("threshold", C.c_float)]


_P = C.c_void_p
_I = C.c_int
_SZ = C.c_size_t
_F = C.c_float

# name -> (restype, argtypes); must list every symbol of include/pcaudio_b200.h
PROTOTYPES = {
    "pca_version": (_I, []),
    "pca_last_error": (C.c_char_p, []),
    "pca_launch_count": (C.c_ulonglong, []),
    "pca_profile_enable": (None, [_I]),
    "pca_profile_report": (_I, [C.c_char_p, _SZ]),
    "pca_stft_logmag_f32": (_I, [_P, _I, _I, _I, _I, _P, _P, _F, _I, _I, _P, _P]),
    "pca_build_clouds_f32": (_I, [_P, _I, _I, _I, _P, _P, _P, _P]),
    "pca_topk_compact_f32": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P]),
    "pca_select_compact_f32": (_I, [_P, _I, _I, _I, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P]),
    "pca_frontend_fused_f32": (_I, [_P, _I, _I, _I, _I, _P, _P, _F, _I, _I, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P]),
    "pca_st_fwd_masked": (_I, [_P, _P, _I, _I, C.POINTER(StDims), _P, _P, _P, _SZ, _I, _P]),
    "pca_deepset_fwd_masked_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "pca_mab_param_count": (C.c_longlong, [_I, _I, _I, _I]),
    "pca_isab_param_count": (C.c_longlong, [_I, _I, _I, _I]),
    "pca_pma_param_count": (C.c_longlong, [_I, _I, _I]),
    "pca_st_param_count": (C.c_longlong, [C.POINTER(StDims)]),
    "pca_mab_workspace_bytes": (_SZ, [_I] * 7),
    "pca_mab_fwd_f32": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "pca_isab_workspace_bytes": (_SZ, [_I] * 6),
    "pca_isab_fwd_f32": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "pca_pma_workspace_bytes": (_SZ, [_I] * 5),
    "pca_pma_fwd_f32": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "pca_st_workspace_bytes": (_SZ, [C.POINTER(StDims), _I, _I, _I]),
    "pca_st_fwd": (_I, [_P, _I, _I, C.POINTER(StDims), _P, _P, _P, _SZ, _I, _P]),
    "pca_deepset_workspace_bytes": (_SZ, [_I] * 5),
    "pca_deepset_fwd_f32": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "pca_st_train_bwd_phase_f32": (_I, [_P, _P, _I, _I, C.POINTER(StDims), _P, _F, C.c_ulonglong, _P, _P, _SZ, _P, _P, _P, _SZ, _I,
                                        C.POINTER(C.c_longlong), _P]),
    "pca_linear_bwd_f32": (_I, [_P, _P, C.c_longlong, _I, _I, _P, _P, _P, _P]),
    "pca_dropout_f32": (_I, [_P, _P, C.c_longlong, _F, C.c_ulonglong, _P]),
    "pca_linear_fwd_f32": (_I, [_P, C.c_longlong, _I, _I, _P, _P, _P]),
    "pca_random_keys_f32": (_I, [_P, C.c_longlong, C.c_ulonglong, _P]),
    "pca_gather_points_f32": (_I, [_P, _I, _I, _I, _P, _P, _P, _I, _P, _P]),
    "pca_importance_map_f32": (_I, [_P, _I, _I, _I, _P, _I, _P, _I, _P, _P, _P]),
    "pca_multinomial_f32": (_I, [_P, _I, _I, _I, C.c_ulonglong, _P, _P, _P]),
    "pca_resample_f32": (_I, [_P, _I, _I, _I, C.c_double, _P, _P, _I, _I, _F, _P, _P]),
    "pca_st_train_saved_bytes": (_SZ, [C.POINTER(StDims), _I, _I, _F]),
    "pca_st_train_workspace_bytes": (_SZ, [C.POINTER(StDims), _I, _I]),
    "pca_st_train_fwd_f32": (_I, [_P, _P, _I, _I, C.POINTER(StDims), _P, _F, C.c_ulonglong, _P, _P, _SZ, _P, _SZ, _P]),
    "pca_st_train_bwd_f32": (_I, [_P, _P, _I, _I, C.POINTER(StDims), _P, _F, C.c_ulonglong, _P, _P, _SZ, _P, _P, _P, _SZ, _P]),
    "pca_mab_train_saved_bytes": (_SZ, [_I] * 7),
    "pca_mab_train_workspace_bytes": (_SZ, [_I] * 7),
    "pca_mab_train_fwd_f32": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P, _SZ, _P]),
    "pca_mab_train_bwd_f32": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P, _P, _P, _P, _SZ, _P]),
    "pca_deepset_train_saved_bytes": (_SZ, [_I, _I, _I]),
    "pca_deepset_train_workspace_bytes": (_SZ, [_I, _I, _I]),
    "pca_deepset_train_fwd_f32": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P, _SZ, _P]),
    "pca_deepset_train_bwd_f32": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P, _P, _P, _SZ, _P]),
    "pca_cross_entropy_f32": (_I, [_P, _P, _I, _I, _P, _P, _P, _P]),
    "pca_adam_step_f32": (_I, [_P, _P, _P, _P, C.c_longlong, _F, _F, _F, _F, _F, _I, _F, _P]),
    "pca_debug_st_stages": (_I, [_P, _I, _I, C.POINTER(StDims), _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "pca_debug_set_timeline": (None, [_P]),
    "pca_debug_set_tail_max": (None, [_I]),
    "pca_debug_set_reduce_variant": (None, [_I]),
    "pca_debug_set_pool_variant": (None, [_I]),
    "pca_debug_set_stft_generic": (None, [_I]),
    "pca_debug_set_gemm_tc": (None, [_I]),
    "pca_debug_linear_tc": (_I, [_P, _P, _I, _P, _P, _P, _P, C.c_longlong, _I, _I, _I, _P, _SZ, _P]),
    "pca_debug_grad_weight_tc": (_I, [_P, _P, _P, C.c_longlong, _I, _I, _P]),
    "pca_debug_set_attn_tc": (None, [_I]),
    "pca_debug_attn_tc_eligible": (_I, [_I, _I, _I, _I, _I]),
    "pca_debug_attn_ws_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "pca_debug_attn_fwd": (_I, [_P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "pca_debug_attn_bwd_tc": (_I, [_P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "pca_debug_umma_probe": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "pca_pipeline_clouds_per_clip": (_I, [C.POINTER(PipelineCfg)]),
    "pca_pipeline_points_per_cloud": (_I, [C.POINTER(PipelineCfg)]),
    "pca_pipeline_workspace_bytes": (_SZ, [C.POINTER(PipelineCfg), _I]),
    "pca_pipeline_run": (_I, [C.POINTER(PipelineCfg), _P, _I, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "pca_pipeline_run_host": (_I, [C.POINTER(PipelineCfg), _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "pca_pipeline_run_host_chunked": (_I, [C.POINTER(PipelineCfg), _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _I, _P, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library once; raise loudly if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"pcaudio_b200: {LIB_PATH} is missing -- the CUDA library must be built "
                "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)      # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().pca_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"pcaudio_b200.{what} failed (code {code}): {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def launch_count() -> int:
    return int(lib().pca_launch_count())


def profile_enable(on: bool) -> None:
    lib().pca_profile_enable(int(on))


def profile_report() -> dict:
    """Per-kernel device time / algorithmic flops / bytes since profile_enable(True)."""
    import json
    buf = C.create_string_buffer(1 << 16)
    check(lib().pca_profile_report(buf, len(buf)), "profile_report")
    return json.loads(buf.value.decode())
