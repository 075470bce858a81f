"""ctypes binding of libmil_b200.so (C ABI: include/mil_b200.h).

No torch types cross the boundary: callers pass `tensor.data_ptr()` integers and the current CUDA stream
handle.  There is no CPU fallback: loading fails loudly when the library has not been built, and every
compute call fails loudly (RuntimeError with the library's own message) without an sm_100a device.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _build

_LOCK = threading.Lock()
_LIB = None

c_void_p, c_int, c_size_t, c_ll, c_char_p, c_float = C.c_void_p, C.c_int, C.c_size_t, C.c_longlong, C.c_char_p, C.c_float

# name -> (restype, argtypes); mirrors include/mil_b200.h one to one
PROTOTYPES = {
    "mil_abi_version": (c_int, []),
    "mil_last_error": (c_char_p, []),
    "mil_kernel_launch_count": (c_ll, []),
    "mil_param_count": (c_int, []),
    "mil_param_name": (c_char_p, [c_int]),
    "mil_param_shape": (c_int, [c_int, C.POINTER(c_int), C.POINTER(c_ll)]),
    "mil_param_offset": (c_ll, [c_int]),
    "mil_param_total": (c_ll, []),
    "mil_set_option": (c_int, [c_char_p, c_int]),
    "mil_get_option": (c_int, [c_char_p, C.POINTER(c_int)]),
    "mil_extractor_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mil_extractor_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t,
                                      c_void_p, c_void_p]),
    "mil_extractor_forward_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t,
                                         c_void_p, c_void_p]),
    "mil_extractor_infer_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mil_extractor_infer": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t,
                                    c_void_p, c_void_p]),
    "mil_extractor_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t,
                                       c_void_p, c_void_p, c_void_p]),
    "mil_extractor_backward_staged": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t,
                                              c_void_p, c_void_p, c_void_p, c_void_p]),
    "mil_extractor_read_activation": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "mil_debug_dump_gradient": (c_int, [c_int, c_int, c_int, c_void_p]),
    "mil_head_workspace_bytes": (c_size_t, [c_int]),
    "mil_head_stats": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "mil_head_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_void_p, c_void_p]),
    "mil_head_finalize": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "mil_head_backward_a": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                    c_void_p, c_void_p]),
    "mil_head_backward_b": (c_int, [c_void_p, c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "mil_pf8_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "mil_to_pf8": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mil_from_pf8": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mil_upsample2_pf8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "mil_conv_workspace_bytes": (c_size_t, [c_int] * 8),
    "mil_conv_pf8": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                             c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                             c_size_t, c_void_p]),
    "mil_conv_wgrad_pf8": (c_int, [c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                   c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mil_stem_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mil_stem_forward": (c_int, [c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                 c_void_p]),
    "mil_stem_backward": (c_int, [c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                  c_void_p]),
    "mil_wide_conv_workspace_bytes": (c_size_t, [c_int] * 5),
    "mil_wide_conv_pf8": (c_int, [c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_int, c_void_p, c_size_t,
                                  c_void_p]),
    "mil_wide_wgrad_workspace_bytes": (c_size_t, [c_int] * 6),
    "mil_wide_wgrad_pf8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                   c_size_t, c_void_p]),
    "mil_merge2_pf8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mil_wide_wgrad_s2_workspace_bytes": (c_size_t, [c_int] * 5),
    "mil_wide_wgrad_s2_pf8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_size_t,
                                      c_void_p]),
    "mil_split2_pf8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mil_wide_param_count": (c_int, [c_void_p]),
    "mil_wide_param_info": (c_int, [c_void_p, c_int, c_char_p, c_int, C.POINTER(c_int), C.POINTER(c_ll), C.POINTER(c_ll)]),
    "mil_wide_param_total": (c_ll, [c_void_p]),
    "mil_wide_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "mil_wide_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p,
                                 c_void_p]),
    "mil_wide_backward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "mil_minmax_normalize": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_void_p]),
    "mil_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_float, c_float, c_float,
                              c_float, c_float, c_void_p]),
    "mil_adam_step_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_void_p, c_void_p]),
    "mil_ingest_tiles_u8": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                    c_void_p, c_void_p, c_void_p]),
}


class MilError(RuntimeError):
    """An entry point of libmil_b200.so returned a non-zero status."""


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = False):
    """Load libmil_b200.so (once).  Raises RuntimeError if it has not been built."""
    global _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = lib_path()
        if not os.path.exists(path):
            if build_if_missing:
                _build.build()
            else:
                raise RuntimeError(
                    f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a).  This package has no CPU / PyTorch fallback path.")
        lib = C.CDLL(path)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if lib.mil_abi_version() != 1:
            raise RuntimeError(f"{path}: ABI version {lib.mil_abi_version()} != 1; rebuild the library")
        _LIB = lib
        return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().mil_last_error().decode(errors="replace")
        raise MilError(f"{what or 'libmil_b200'} failed (status {status}): {msg}")


def param_table():
    """[(name, shape tuple, float offset)] for the 65 state-dict tensors, straight from the library."""
    lib = load()
    out = []
    nd = c_int(0)
    shp = (c_ll * 4)()
    for i in range(lib.mil_param_count()):
        check(lib.mil_param_shape(i, C.byref(nd), shp), "mil_param_shape")
        out.append((lib.mil_param_name(i).decode(), tuple(int(shp[k]) for k in range(nd.value)),
                    int(lib.mil_param_offset(i))))
    return out


def set_option(name: str, value: int) -> None:
    """Runtime switch of the library (include/mil_b200.h: "disable_tc", "stem_unfused", ...)."""
    check(load().mil_set_option(name.encode(), int(value)), "mil_set_option")


def get_option(name: str) -> int:
    v = c_int(0)
    check(load().mil_get_option(name.encode(), C.byref(v)), "mil_get_option")
    return int(v.value)
