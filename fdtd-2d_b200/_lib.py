"""ctypes binding of libfdtd2d.so (include/fdtd2d.h).  No fallback: if the library is missing the
import of any compute entry point raises, loudly."""
from __future__ import annotations

import ctypes
import os

from .build import LIB_PATH

F32, F64 = 0, 1
PHASE_H, PHASE_E, PHASE_SRC = 1, 2, 4
MAX_K = 12
PEER_BLOB_BYTES = 640
PLAN_INFO_WORDS = 13

_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_u64 = ctypes.c_uint64
_u32 = ctypes.c_uint32
_sz = ctypes.c_size_t
_d = ctypes.c_double
_pp = ctypes.POINTER(_vp)
_ip = ctypes.POINTER(_i)

_SIGNATURES = {
    "fdtd2d_abi_version": ([], _i),
    "fdtd2d_last_error": ([], ctypes.c_char_p),
    "fdtd2d_device_count": ([_ip], _i),
    "fdtd2d_create": ([_pp, _i, _i, _i, _i, _i], _i),
    "fdtd2d_create_slab": ([_pp, _i, _i, _i, _i, _i, _i, _i], _i),
    "fdtd2d_destroy": ([_vp], _i),
    "fdtd2d_set_stream": ([_vp, _vp], _i),
    "fdtd2d_reset_stream": ([_vp], _i),
    "fdtd2d_get_stream": ([_vp, _pp], _i),
    "fdtd2d_sync": ([_vp], _i),
    "fdtd2d_set_option": ([_vp, ctypes.c_char_p, _i], _i),
    "fdtd2d_get_option": ([_vp, ctypes.c_char_p, _ip], _i),
    "fdtd2d_geometry": ([_vp, _ip, _ip, _ip, _ip, _ip, _ip, ctypes.POINTER(_sz)], _i),
    "fdtd2d_upload_state": ([_vp, _vp, _vp, _vp], _i),
    "fdtd2d_download_state": ([_vp, _vp, _vp, _vp], _i),
    "fdtd2d_zero_state": ([_vp], _i),
    "fdtd2d_upload_state_async": ([_vp, _vp, _vp, _vp], _i),
    "fdtd2d_download_state_async": ([_vp, _vp, _vp, _vp], _i),
    "fdtd2d_copy_wait": ([_vp], _i),
    "fdtd2d_set_materials_async": ([_vp, _vp, _vp, _d, _d], _i),
    "fdtd2d_set_coeffs": ([_vp, _vp, _vp, _vp], _i),
    "fdtd2d_set_materials": ([_vp, _vp, _vp, _d, _d], _i),
    "fdtd2d_set_mur_coef": ([_vp, _vp], _i),
    "fdtd2d_set_materials_random": ([_vp, _u64, _d, _d, _d], _i),
    "fdtd2d_set_materials_gray": ([_vp, _vp, _d, _d, _d], _i),
    "fdtd2d_canvas_clear": ([_vp, _i], _i),
    "fdtd2d_canvas_rect": ([_vp, _i, _i, _i, _i, _i, _i], _i),
    "fdtd2d_canvas_ellipse": ([_vp, _i, _i, _i, _i, _i, _i, _i], _i),
    "fdtd2d_canvas_segment": ([_vp, _i, _d, _d, _d, _d, _d, _i], _i),
    "fdtd2d_canvas_download": ([_vp, _vp], _i),
    "fdtd2d_canvas_apply": ([_vp, _d, _d, _d], _i),
    "fdtd2d_generate_materials_blobs": ([_vp, _u64, _vp, _d, _d, _d, _d, _d, _vp], _i),
    "fdtd2d_hash_uniform": ([_u64, _u32, _u32, _u32], _d),
    "fdtd2d_download_coeffs": ([_vp, _vp, _vp, _vp], _i),
    "fdtd2d_set_sources": ([_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp], _i),
    "fdtd2d_set_probes": ([_vp, _i, _vp, _vp, _vp, _i], _i),
    "fdtd2d_read_probes": ([_vp, _vp, _i64, _i], _i),
    "fdtd2d_step": ([_vp, _i, _i], _i),
    "fdtd2d_step_phases": ([_vp, _i], _i),
    "fdtd2d_get_step_index": ([_vp, ctypes.POINTER(_i64)], _i),
    "fdtd2d_set_step_index": ([_vp, _i64], _i),
    "fdtd2d_source_steps": ([_vp, _ip, _ip, ctypes.POINTER(_i64)], _i),
    "fdtd2d_set_kernel_variant": ([_vp, _i], _i),
    "fdtd2d_plan_host": ([_vp, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _i], _i),
    "fdtd2d_plan_host_fused": ([_vp, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _vp, _i], _i),
    "fdtd2d_plan_info": ([_vp, _i, _vp, _i], _i),
    "fdtd2d_launch_count": ([_vp, ctypes.POINTER(_i64)], _i),
    "fdtd2d_pass_count": ([_vp, ctypes.POINTER(_i64)], _i),
    "fdtd2d_plan_wave_runs": ([_i, _vp, _vp, _i, _i, _i, _vp, _vp], _i),
    "fdtd2d_plan_resident": ([_i, _i, _i, _i, _vp], _i),
    "fdtd2d_plan_edge_reserve": ([_i64, _i64, _i, _i], _i),
    "fdtd2d_halo_block": ([_vp, _i, _i, _pp, _pp, ctypes.POINTER(_sz)], _i),
    "fdtd2d_halo_block_next": ([_vp, _i, _i, _pp, _pp, ctypes.POINTER(_sz)], _i),
    "fdtd2d_peer_export": ([_vp, _vp], _i),
    "fdtd2d_peer_attach": ([_vp, _i, _vp], _i),
    "fdtd2d_peer_detach": ([_vp], _i),
    "fdtd2d_peer_status": ([_vp, _vp], _i),
    "fdtd2d_pass_begin": ([_vp, _i], _i),
    "fdtd2d_pass_end": ([_vp], _i),
    "fdtd2d_set_snapshot_background": ([_vp, _vp, _vp], _i),
    "fdtd2d_render_snapshot": ([_vp, _i, _d, _d, _vp], _i),
    "fdtd2d_device_field": ([_vp, _i, _pp], _i),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


class Fdtd2dError(RuntimeError):
    """A libfdtd2d call returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libfdtd2d error {code}: {message}")
        self.code = code


def lib():
    """The loaded library (loads on first use).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)"
            )
        L = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here means the .so is stale
            fn.argtypes = argtypes
            fn.restype = restype
        if L.fdtd2d_abi_version() != 1:
            raise RuntimeError("libfdtd2d ABI version mismatch; rebuild")
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise Fdtd2dError(rc, lib().fdtd2d_last_error().decode(errors="replace"))
