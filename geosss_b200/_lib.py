"""
ctypes binding of libgeosss_cuda.so (include/geosss_cuda.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is
raised.  The library is built in-tree by ``__graft_entry__.build()`` / ``make -C geosss_b200/csrc``.
"""

import ctypes as C
import os

F32, F64 = 0, 1
REJECT, SHRINK = 0, 1
RWMH, INDEPENDENCE, MIXTURE, HMC = 2, 3, 4, 5
PATH_LITERAL, PATH_FAST = 0, 1
STEP_MAJOR, CHAIN_MAJOR = 0, 1

E_INVALID, E_CUDA, E_UNSUPPORTED, E_NOMEM = -1, -2, -3, -4

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgeosss_cuda.so")

_p, _i, _i64, _u64, _dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
_pd = C.POINTER(C.c_double)

# name -> (restype, argtypes); one entry per function declared in include/geosss_cuda.h
SIGNATURES = {
    "geosss_version": (_i, []),
    "geosss_last_error": (C.c_char_p, []),
    "geosss_target_create_vmf_mixture": (_i, [_i, _i, _pd, _pd, _pd, C.POINTER(_p)]),
    "geosss_target_create_bingham": (_i, [_i, _pd, _pd, C.POINTER(_p)]),
    "geosss_target_create_curved_vmf": (_i, [_i, _i, _pd, _dbl, C.POINTER(_p)]),
    "geosss_target_create_cpd": (_i, [_i, _i, _pd, _pd, _pd, _pd, _dbl, _dbl, _dbl, _i, _i, C.POINTER(_p)]),
    "geosss_target_create_cpd_projected": (_i, [_i, _i, _pd, _pd, _pd, _pd, _dbl, _dbl, _dbl, _i, _i, C.POINTER(_p)]),
    "geosss_target_cpd_set_translation": (_i, [_p, _pd]),
    "geosss_target_cpd_grid_info": (_i, [_p, _pd, _pd, C.POINTER(_i), C.POINTER(_i)]),
    "geosss_target_dim": (_i, [_p]),
    "geosss_target_bingham_eigenrows": (_i, [_p, _pd]),
    "geosss_target_destroy": (_i, [_p]),
    "geosss_log_prob": (_i, [_p, _i, _i, _p, _p, _i64, _p]),
    "geosss_sampler_create": (_i, [_p, _i, _i, _i, _i64, _u64, _u64, C.POINTER(_p)]),
    "geosss_sampler_set_state": (_i, [_p, _p, _p]),
    "geosss_sampler_get_state": (_i, [_p, _p, _p, _p]),
    "geosss_sampler_run": (_i, [_p, _i64, _i64, _p, _p, _p, _p]),
    "geosss_sampler_run_layout": (_i, [_p, _i64, _i64, _i, _i64, _i64, _p, _p, _p, _p]),
    "geosss_sampler_run_block": (_i, [_p, _i64, _i64, _i64, _p, _p]),
    "geosss_sampler_run_tape": (_i, [_p, _i64, _p, _p, _p, _p, _p, _p]),
    "geosss_sampler_steps_done": (_i64, [_p]),
    "geosss_sampler_trace_out": (_i, [_p, _pd, _p]),
    "geosss_sampler_destroy": (_i, [_p]),
    "geosss_release_cached_memory": (_i, []),
    "geosss_sampler_set_state_host": (_i, [_p, _pd]),
    "geosss_sampler_get_state_host": (_i, [_p, _pd, _p]),
    "geosss_sampler_run_host": (_i, [_p, _i64, _i64, _p]),
    "geosss_diag_ess_partials": (_i, [_p, _i, _i, _i64, _i64, _i, _p, _i, _p, _p]),
    "geosss_diag_stream_update": (_i, [_p, _i, _i, _i64, _i64, _i64, _i, _p, _p]),
    "geosss_diag_stream_finish": (_i, [_p, _i64, _i64, _i, _p, _p]),
    "geosss_sampler_mh_configure": (_i, [_p, C.c_double, C.c_double, _i]),
    "geosss_sampler_mh_reset": (_i, [_p, _i64]),
    "geosss_sampler_mh_momenta_out": (_i, [_p, _p]),
    "geosss_sampler_mh_get_host": (_i, [_p, _pd, _p, _pd]),
    "geosss_gradient": (_i, [_p, _i, _p, _p, _i64, _p]),
    "geosss_exact_vmf": (_i, [_i, _pd, _u64, _u64, _i, _p, _i64, C.POINTER(_u64), _p]),
    "geosss_exact_bingham": (_i, [_i, _pd, _pd, _dbl, _dbl, _u64, _u64, _i, _p, _i64, _i, C.POINTER(_u64), _p]),
    "geosss_diag_path_stats": (_i, [_p, _i, _i, _i64, _i64, _i, _p, _i, _i, _p, _p]),
    "geosss_sample_host": (_i, [_p, _i, _i, _i, _i64, _pd, _u64, _u64, _i64, _i64, _p, _p]),
}

_lib = None


class GeosssError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libgeosss_cuda error {code}: {message}")
        self.code = code


def lib():
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C geosss_b200/csrc`. geosss_b200 has no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise GeosssError(rc, lib().geosss_last_error().decode())


def dptr(arr):
    """numpy float64 array -> POINTER(c_double) (None passes NULL)."""
    return None if arr is None else arr.ctypes.data_as(_pd)
