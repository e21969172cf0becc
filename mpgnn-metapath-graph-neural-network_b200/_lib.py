"""ctypes binding of libmpgnn_b200.so (the C ABI in include/mpgnn_b200.h).

There is no CPU fallback: if the library is missing, or a call fails, this raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MPGNN_B200_LIB: an alternative build of the same library (profiling / ablation builds only)
LIB_PATH = os.environ.get("MPGNN_B200_LIB") or os.path.join(HERE, "libmpgnn_b200.so")

OK, EINVAL, ECUDA, ERANGE, ENOTSUP = 0, -1, -2, -3, -4
F_RELU, F_DROPOUT_SEED, F_DROPOUT_MASK, F_NEED_GX, F_TF32X3, F_BF16 = 1, 2, 4, 8, 16, 32
F_COMPACT_H, F_DENSE_H = 64, 128

_c = ctypes
_i64, _i32, _u32, _u64, _dbl, _ptr = _c.c_int64, _c.c_int, _c.c_uint32, _c.c_uint64, _c.c_double, _c.c_void_p

# name -> (restype, argtypes); mirrors include/mpgnn_b200.h one to one
PROTOTYPES = {
    "mpgnn_last_error": (_c.c_char_p, []),
    "mpgnn_abi_version": (_i32, []),
    "mpgnn_launch_count": (_c.c_longlong, []),
    "mpgnn_timing_enable": (None, [_i32]),
    "mpgnn_timing_reset": (None, []),
    "mpgnn_timing_collect": (_i32, [_c.c_char_p, _i64, _ptr, _ptr, _i64]),
    "mpgnn_graph_build": (_i32, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _c.POINTER(_ptr)]),
    "mpgnn_graph_build_host": (_i32, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _c.POINTER(_ptr)]),
    "mpgnn_graph_free": (None, [_ptr]),
    "mpgnn_graph_info": (_i32, [_ptr, _c.POINTER(_i64), _c.POINTER(_i64), _c.POINTER(_i64)]),
    "mpgnn_graph_relation_view": (_i32, [_ptr, _i64, _i32, _c.POINTER(_ptr), _c.POINTER(_ptr), _c.POINTER(_ptr),
                                         _c.POINTER(_i64)]),
    "mpgnn_graph_relation_counts": (_i32, [_ptr, _ptr]),
    "mpgnn_spmm": (_i32, [_ptr, _i64, _i32, _i32, _ptr, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr]),
    "mpgnn_hop_fwd": (_i32, [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _ptr, _i64, _u32, _dbl, _u64, _u64, _ptr, _ptr,
                             _ptr, _ptr, _ptr, _i64, _ptr]),
    "mpgnn_hop_bwd": (_i32, [_ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _u32, _dbl, _ptr,
                             _ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "mpgnn_scale_rows_by_degree": (_i32, [_ptr, _i64, _ptr, _i64, _i64, _ptr, _i64, _ptr]),
    "mpgnn_hop_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "mpgnn_hop_h_rows": (_i64, [_ptr, _i64, _i64, _i64, _u32]),
    "mpgnn_graph_relation_rows": (_i32, [_ptr, _i64, _c.POINTER(_ptr), _c.POINTER(_i64)]),
    "mpgnn_gemm_rows": (_i32, [_ptr, _i64, _i64, _i64, _ptr, _i64, _i64, _i64, _ptr, _i32, _ptr, _i64, _ptr, _i64,
                               _ptr, _i64, _ptr]),
    "mpgnn_gemm_tn": (_i32, [_ptr, _i64, _i64, _i64, _ptr, _i64, _i64, _ptr, _i64, _ptr, _ptr, _i64, _ptr]),
    "mpgnn_gemm_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "mpgnn_logsoftmax_nll": (_i32, [_ptr, _i64, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "mpgnn_macro_f1": (_i32, [_ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _ptr]),
    "mpgnn_adam_step": (_i32, [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _dbl, _dbl, _dbl, _dbl, _dbl, _ptr]),
    "mpgnn_score_bags_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "mpgnn_score_bags": (_i32, [_ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _ptr, _i32, _i64, _dbl,
                                _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "mpgnn_trainer_create": (_i32, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _dbl,
                                    _u64, _u32, _i64, _c.POINTER(_ptr)]),
    "mpgnn_trainer_create_multi": (_i32, [_ptr, _ptr, _i64, _i64, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _ptr, _i64,
                                          _dbl, _u64, _u32, _i64, _c.POINTER(_ptr)]),
    "mpgnn_set_tc_cta_cap": (None, [_i32]),
    "mpgnn_trainer_free": (None, [_ptr]),
    "mpgnn_trainer_num_params": (_i64, [_ptr]),
    "mpgnn_trainer_set_params": (_i32, [_ptr, _ptr, _ptr]),
    "mpgnn_trainer_get_params": (_i32, [_ptr, _ptr, _ptr]),
    "mpgnn_trainer_run": (_i32, [_ptr, _i64, _dbl, _dbl, _dbl, _dbl, _dbl, _i32, _ptr, _ptr, _ptr]),
    "mpgnn_trainer_evaluate": (_i32, [_ptr, _ptr, _ptr, _i64, _ptr, _ptr, _ptr]),
    "mpgnn_score_workspace_bytes": (_i64, [_i64]),
    "mpgnn_score_relation": (_i32, [_ptr, _i64, _ptr, _ptr, _ptr, _i64, _dbl, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
}

_lib = None


def load():
    """Load the shared library (once) and type its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "mpgnn_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    msg = load().mpgnn_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc):
    """Translate a status code into the exception the reference's Python path would raise."""
    if rc == OK:
        return
    msg = last_error()
    if rc in (EINVAL, ERANGE):
        raise ValueError("mpgnn_b200: " + msg)
    if rc == ENOTSUP:
        raise NotImplementedError("mpgnn_b200: " + msg)
    raise RuntimeError("mpgnn_b200: " + msg)


def current_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def timing_collect():
    """{kernel class: (milliseconds, calls)} accumulated since the last timing_reset()."""
    import numpy as np
    lib = load()
    names = ctypes.create_string_buffer(4096)
    ms = np.zeros(64, dtype=np.float64)
    calls = np.zeros(64, dtype=np.int64)
    n = lib.mpgnn_timing_collect(names, 4096, ms.ctypes.data_as(ctypes.c_void_p), calls.ctypes.data_as(ctypes.c_void_p),
                                 64)
    keys = [k for k in names.value.decode().split(";") if k][:n]
    return {k: (float(ms[i]), int(calls[i])) for i, k in enumerate(keys)}
