"""ctypes binding of ``libslode_b200.so`` (the C ABI declared in ``include/slode_b200.h``).

There is no fallback: if the shared library has not been built (``python -c "import
__graft_entry__ as g; g.build()"``) importing the compute entry points raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# SLODE_B200_LIB points at another build of the same library (kernel A/B measurements); there is still no fallback
LIB_PATH = os.environ.get("SLODE_B200_LIB") or os.path.join(_HERE, "csrc", "libslode_b200.so")

# slode_b200.h constants
METHOD_EULER, METHOD_MIDPOINT, METHOD_RK4, METHOD_DOPRI5 = 0, 1, 2, 3
METHODS = {"euler": METHOD_EULER, "midpoint": METHOD_MIDPOINT, "rk4": METHOD_RK4, "dopri5": METHOD_DOPRI5}
BWD_DISCRETE, BWD_TDE_ADJOINT = 0, 1
F32, F64 = 0, 1
Q_VERSION, Q_SM_ARCH, Q_MAX_HIDDEN, Q_MAX_STATE, Q_N_SHAPES = 0, 1, 2, 3, 4
Q_FWD_LAUNCHES, Q_BWD_LAUNCHES, Q_TOTAL_LAUNCHES, Q_SOURCE_HASH, Q_SHAPE_BASE = 10, 11, 12, 13, 100

_i, _i64, _p = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p

# symbol -> (restype, argtypes); every symbol include/slode_b200.h declares
SIGNATURES = {
    "slode_query": (_i, [_i]),
    "slode_last_error": (ctypes.c_char_p, []),
    "slode_mlp_supported": (_i, [_i, _i]),
    "slode_dopri5_supported": (_i, [_i, _i]),
    "slode_fixed_workspace_bytes": (_i64, [_i, _i, _i, _i64, _i, _i, _i, _i, _i, _i]),
    "slode_mlp_fixed_fwd": (_i, [_i, _i64, _i, _i, _i] + [_p] * 8 + [_p, _i64, _i64, _p, _i64, _p]),
    "slode_mlp_fixed_bwd": (_i, [_i, _i, _i64, _i, _i, _i] + [_p] * 7 + [_p, _i64, _i64, _p, _i64, _i64]
                            + [_p, _p, _p, _p, _i64, _p]),
    "slode_latent_fixed_fwd": (_i, [_i, _i64, _i, _i, _i, _i] + [_p] * 13 + [_p, _i64, _i64, _p, _i64, _p]),
    "slode_latent_fixed_heads_fwd": (_i, [_i, _i64, _i, _i, _i, _i] + [_p] * 13 + [_i, _i, _p, _p, _i64]
                                     + [_p, _i64, _i64, _p, _i64, _p]),
    "slode_latent_fixed_bwd": (_i, [_i, _i, _i64, _i, _i, _i, _i] + [_p] * 12 + [_p, _i64, _i64, _p, _i64, _i64]
                               + [_p, _p, _p, _p, _i64, _p]),
    "slode_mlp_dopri5_fwd": (_i, [_i64, _i, _i, _i] + [_p] * 8 + [ctypes.c_double] * 3 + [_i64, _p, _i64]
                             + [_p, _i64, _i64, _p, _i64, _p, _i64, _p, _p]),
    "slode_mlp_dopri5_step_workspace_bytes": (_i64, [_i64, _i]),
    "slode_mlp_dopri5_fwd_step": (_i, [_i64, _i, _i, _i] + [_p] * 8 + [ctypes.c_double] * 3 + [_i64, _i64, _i, _p, _p]
                                  + [_p, _i64, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _p]),
    "slode_mlp_dopri5_bwd": (_i, [_i64, _i, _i, _i] + [_p] * 7 + [_i64, _p, _p, _p, _p, _i64, _i64, _p, _p, _p, _p]),
    "slode_mlp_dopri5_adjoint_workspace_bytes": (_i64, [_i64, _i, _i, _i]),
    "slode_mlp_dopri5_adjoint_bwd": (_i, [_i64, _i, _i, _i, _i] + [_p] * 8 + [_p, _i64, _i64, _p, _i64, _i64]
                                     + [ctypes.c_double, ctypes.c_double, _i64, _p, _i64, _p, _p, _p, _i64, _p, _p, _i64, _p]),
    "slode_heads_fwd": (_i, [_i64, _i, _i, _i, _i, _p, _i64, _i64, _p, _p, _p]),
    "slode_heads_bwd": (_i, [_i64, _i, _i, _i, _i, _p, _i64, _i64, _p, _p, _p, _i64, _i64, _p, _p]),
    "slode_heads_bwd_split": (_i, [_i64, _i, _i, _i, _i, _p, _i64, _i64, _p, _p, _p, _p, _p, _i64, _i64, _p, _p]),
    "slode_cvs_fixed_fwd": (_i, [_i, _i, _i64, _i, _i] + [_p] * 5 + [_p, _i64, _i64, _p]),
    "slode_cvs_fixed_bwd": (_i, [_i, _i, _i, _i64, _i, _i] + [_p] * 4 + [_p, _i64, _i64, _p, _i64, _i64]
                            + [_p] * 5),
}

_lib = None
_lock = threading.Lock()


class SlodeError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it is missing (no CPU / eager fallback exists)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.isfile(LIB_PATH):
                    raise SlodeError(
                        f"{LIB_PATH} is missing: build the CUDA extension first "
                        "(python -c 'import __graft_entry__ as g; g.build()'). There is no fallback path.")
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype, fn.argtypes = res, args
                _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().slode_last_error().decode("utf-8", "replace")
        if rc == 2:
            raise NotImplementedError(f"{what}: {msg}")
        raise SlodeError(f"{what} failed (code {rc}): {msg}")


def supported_shapes():
    L = lib()
    n = L.slode_query(Q_N_SHAPES)
    return [(L.slode_query(Q_SHAPE_BASE + 2 * i), L.slode_query(Q_SHAPE_BASE + 2 * i + 1)) for i in range(n)]
