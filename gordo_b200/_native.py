"""
ctypes binding of libgordo_b200.so (include/gordo_b200.h).  There is NO CPU fallback: if the
library is missing, or a call fails, this module raises -- the product path never routes
around the CUDA kernels.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgordo_b200.so")

MAX_LAYERS = 16
ACT_CODES = {"linear": 0, "tanh": 1, "relu": 2, "sigmoid": 3, "elu": 4, "softplus": 5}
PREC_F32, PREC_BF16_TC, PREC_F16X3_TC = 0, 1, 2
PREC_CODES = {"f32": PREC_F32, "bf16": PREC_BF16_TC, "f16x3": PREC_F16X3_TC}


class FFArch(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("widths", C.c_int32 * (MAX_LAYERS + 1)),
                ("acts", C.c_int32 * MAX_LAYERS), ("l1", C.c_float * MAX_LAYERS)]


class LSTMArch(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("n_features", C.c_int32), ("n_features_out", C.c_int32),
                ("units", C.c_int32 * MAX_LAYERS), ("acts", C.c_int32 * MAX_LAYERS),
                ("out_act", C.c_int32), ("lookback_window", C.c_int32), ("lookahead", C.c_int32)]


class Adam(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta_1", C.c_float), ("beta_2", C.c_float), ("epsilon", C.c_float)]


_P = C.c_void_p
_I32, _I64 = C.c_int32, C.c_int64

# name -> (restype, argtypes); every symbol include/gordo_b200.h declares
SIGNATURES = {
    "gb200_abi_version": (C.c_int, []),
    "gb200_last_error": (C.c_char_p, []),
    "gb200_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3 + [C.c_char_p, C.c_int]),
    "gb200_fleet_create": (C.c_int, [C.POINTER(_P), _I32, C.POINTER(_I64)]),
    "gb200_fleet_create_ranges": (C.c_int, [C.POINTER(_P), _I32, C.POINTER(_I64), C.POINTER(_I64)]),
    "gb200_fleet_destroy": (None, [_P]),
    "gb200_ff_score": (C.c_int, [_P, C.POINTER(FFArch), C.c_int] + [_P] * 18),
    "gb200_ff_packed_bytes": (_I64, [C.POINTER(FFArch)]),
    "gb200_ff_pack_bf16": (C.c_int, [C.POINTER(FFArch), _I32, _P, _P, _P]),
    "gb200_ff_packed_bytes_prec": (_I64, [C.POINTER(FFArch), C.c_int]),
    "gb200_ff_pack": (C.c_int, [C.POINTER(FFArch), C.c_int, _I32, _P, _P, _P]),
    "gb200_ff_param_count": (_I64, [C.POINTER(FFArch)]),
    "gb200_minmax_fit": (C.c_int, [_I32, _P, _P, _P, _I32, _P, _P, _P]),
    "gb200_rolling_min_max": (C.c_int, [_I32, _P, _P, _P, _I32, _I32, _P, _P]),
    "gb200_smooth": (C.c_int, [_I32, _P, _P, _P, _I32, _I32, _I32, _P, _P]),
    "gb200_quantile": (C.c_int, [_I32, _P, _P, _P, _I32, C.c_double, _P, _P]),
    "gb200_cv_sums": (C.c_int, [_I32, _P, _P, _P, _P, _I32, _P, _P]),
    "gb200_ff_fit": (C.c_int, [C.POINTER(FFArch), C.POINTER(Adam), _I32] + [_P] * 9 + [_I32] * 3 + [_P] * 6),
    "gb200_lstm_param_count": (_I64, [C.POINTER(LSTMArch)]),
    "gb200_lstm_out_rows": (_I64, [C.POINTER(LSTMArch), _I64]),
    "gb200_lstm_scratch_bytes": (_I64, [C.POINTER(LSTMArch), _I64, C.c_int]),
    "gb200_lstm_predict": (C.c_int, [_P, C.POINTER(LSTMArch), C.c_int] + [_P] * 6 + [_P, _I64, _P]),
    "gb200_lstm_fit_scratch_bytes": (_I64, [C.POINTER(LSTMArch), _I32, _I32]),
    "gb200_lstm_fit": (C.c_int, [C.POINTER(LSTMArch), C.POINTER(Adam), _I32, C.POINTER(_I64), C.POINTER(_I64)]
                       + [_P] * 4 + [_I32, _I32] + [_P] * 3 + [_P, _I64, _P]),
    "gb200_score_outputs": (C.c_int, [_I32, _P, _P, _I32] + [_P] * 12),
    "gb200_host_expand_columns": (C.c_int, [_I32, _P, _I32] + [_P] * 7 + [_I32]),
    "gb200_host_stream_seconds": (C.c_double, [_P, _P, _I64, _I32, _I32]),
    "gb200_resample": (C.c_int, [_I32] + [_P] * 7 + [_I64, _I32, _I64, _I64, _I64, _P, _P]),
    "gb200_interpolate": (C.c_int, [_I32, _P, _P, _P, _I32, _I64, _P, _P]),
    "gb200_filter_rows": (C.c_int, [_I32, _P, _P, _P, _I32, _P, _I64, C.POINTER(_I32), C.POINTER(_I32), _I32,
                                    C.POINTER(C.c_double), _I32, _I32, _P, _P]),
    "gb200_compact_rows": (C.c_int, [_I32, _P, _P, _P, _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m gordo_b200.build_native` "
                "(gordo_b200 has no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().gb200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libgordo_b200 {what} failed ({rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def make_ff_arch(widths, acts, l1=None):
    n = len(widths) - 1
    if not (1 <= n <= MAX_LAYERS):
        raise ValueError(f"feed-forward stack of {n} Dense layers is outside [1, {MAX_LAYERS}]")
    a = FFArch()
    a.n_layers = n
    for i, w in enumerate(widths):
        a.widths[i] = int(w)
    for i, name in enumerate(acts):
        if name not in ACT_CODES:
            raise ValueError(f"unsupported activation {name!r}; supported: {sorted(ACT_CODES)}")
        a.acts[i] = ACT_CODES[name]
    for i in range(n):
        a.l1[i] = float(l1[i]) if l1 is not None else 0.0
    return a


def make_lstm_arch(n_features, n_features_out, units, acts, out_act, lookback_window, lookahead):
    n = len(units)
    if not (1 <= n <= MAX_LAYERS):
        raise ValueError(f"LSTM stack of {n} layers is outside [1, {MAX_LAYERS}]")
    a = LSTMArch()
    a.n_layers, a.n_features, a.n_features_out = n, int(n_features), int(n_features_out)
    for i, (u, name) in enumerate(zip(units, acts)):
        if name not in ACT_CODES:
            raise ValueError(f"unsupported activation {name!r}")
        a.units[i] = int(u); a.acts[i] = ACT_CODES[name]
    if out_act not in ACT_CODES:
        raise ValueError(f"unsupported activation {out_act!r}")
    a.out_act = ACT_CODES[out_act]
    a.lookback_window, a.lookahead = int(lookback_window), int(lookahead)
    return a
