"""ctypes binding of libb2n.so (the C ABI declared in include/b2n.h).

There is no CPU fallback: if the library is missing or a call fails this module raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2N_LIB") or os.path.join(_HERE, "lib", "libb2n.so")     # B2N_LIB: a variant build (experiments)

MAX_LEVELS = 32


class GridLayout(C.Structure):
    """b2n_grid_layout (include/b2n.h)."""
    _fields_ = [("n_levels", C.c_int32), ("n_features", C.c_int32),
                ("scale", C.c_float * MAX_LEVELS), ("resolution", C.c_uint32 * MAX_LEVELS),
                ("size", C.c_uint32 * MAX_LEVELS), ("offset", C.c_uint32 * (MAX_LEVELS + 1)),
                ("x_offset", C.c_float), ("x_scale", C.c_float)]

    @property
    def n_entries(self):
        return int(self.offset[self.n_levels])

    @property
    def n_params(self):
        return self.n_entries * self.n_features


_P, _I, _L, _F, _D = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double
_SIGS = {
    "b2n_ray_aabb_intersect": [_P, _P, _P, _P, _L, _L, _I, _P, _P, _P, _P],
    "b2n_ray_sphere_intersect": [_P, _P, _P, _P, _L, _L, _I, _P, _P, _P, _P],
    "b2n_clamp_near": [_P, _L, _F, _P],
    "b2n_render_schedule": [_P, _L, _I, _I, _P],
    "b2n_render_rays": [_P, _P, _P, _L, _P, _I, _F, _F, _I, _I, C.POINTER(GridLayout), _P, _P, _F, _F, _P, _P, _P, _I, _P, _P, _P,
                        _P, _P],
    "b2n_raymarching_test_dev": [_P, _P, _P, _P, _P, _I, _F, _F, _I, _I, _L, _P, _P, _P, _P, _P, _P, _P],
    "b2n_composite_test_fw_dev": [_P, _P, _P, _P, _P, _P, _F, _P, _L, _P, _P, _P, _P, _P],
    "b2n_peer_alloc": [_L, C.POINTER(C.c_void_p), _P],
    "b2n_peer_open": [_P, C.POINTER(C.c_void_p)],
    "b2n_peer_close": [_P],
    "b2n_peer_free": [_P],
    "b2n_peer_barrier": [_P, _I, _I, _P, _D, _P],
    "b2n_adam_step_peer": [_P, _P, _P, _P, _P, _L, _L, _P, _I, _L, _L, _F, _F, _F, _F, _F, _I, _P, _P, _P, _P],
    "b2n_scaler_update": [_P, _P, _I, _P],
    "b2n_field_image_halves": [_I],
    "b2n_grad_pack_half": [_P, _P, _L, _L, _P],
    "b2n_rays_from_indices": [_P, _P, _P, _P, _L, _P, _P, _P],
    "b2n_morton3D": [_P, _L, _P, _P],
    "b2n_morton3D_invert": [_P, _L, _P, _P],
    "b2n_packbits": [_P, _L, _F, _P, _P, _P],
    "b2n_raymarching_train_count": [_P, _P, _P, _P, _I, _F, _F, _P, _I, _I, _L, _L, _P, _P, _P, _P],
    "b2n_raymarching_train_count_serial": [_P, _P, _P, _P, _I, _F, _F, _P, _I, _I, _L, _L, _P, _P, _P, _P],
    "b2n_raymarching_train_write": [_P, _P, _P, _P, _I, _F, _F, _P, _I, _I, _L, _P, _P, _P, _P, _P, _P, _P],
    "b2n_raymarcher_bw": [_P, _P, _P, _P, _L, _P, _P, _P],
    "b2n_raymarching_test": [_P, _P, _P, _P, _P, _I, _F, _F, _I, _I, _I, _L, _P, _P, _P, _P, _P, _P],
    "b2n_composite_train_fw": [_P, _P, _P, _P, _P, _F, _L, _P, _P, _P, _P, _P],
    "b2n_composite_train_bw": [_P] * 13 + [_F, _L, _P, _P, _P, _P, _P],
    "b2n_composite_loss_fwbw": [_P] * 6 + [_F, _L, _F, _F, _F] + [_P] * 10,
    "b2n_composite_test_fw": [_P, _P, _P, _P, _P, _P, _F, _P, _I, _L, _P, _P, _P, _P],
    "b2n_hashgrid_layout": [_I, _I, _I, _I, _D, C.POINTER(GridLayout)],
    "b2n_hashgrid_fw": [_P, _P, C.POINTER(GridLayout), _L, _P, _P, _I, _P],
    "b2n_hashgrid_bw": [_P, _P, _I, C.POINTER(GridLayout), _L, _P, _F, _P, _P, _P],
    "b2n_frequency_fw": [_P, _I, _L, _P, _P, _I, _F, _F, _P],
    "b2n_sh4_fw": [_P, _I, _L, _P, _P, _I, _P],
    "b2n_mlp_fw": [_P, _I, _I, _P, _I, _I, _L, _P, _P, _P, _P],
    "b2n_mlp_bw": [_P, _P, _I, _I, _P, _I, _I, _L, _P, _P, _P, _F, _P, _P, _P],
    "b2n_field_pack_weights": [_P, _P, _P, _I, _P, _P],
    "b2n_field_mlp_fw": [_P, _I, _P, _P, _L, _P, _P, _P, _P, _P],
    "b2n_field_mlp_bw": [_P, _P, _P, _I, _P, _P, _L, _P, _P, _P, _F, _P, _P, _P, _P, _I, _P, _P],
    "b2n_adam_step": [_P, _P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _P, _I, _P],
    "b2n_cast_half": [_P, _P, _L, _P],
    "b2n_grid_cell_positions": [_P, _P, _L, _I, _F, _F, _F, _I, _P, _P],
    "b2n_grid_scatter": [_P, _P, _L, _P, _L, _P],
    "b2n_grid_ema": [_P, _P, _L, _F, _P],
    "b2n_grid_threshold": [_P, _L, _F, _P, _P, _P],
    "b2n_mark_invisible_cells": [_P, _P, _I, _I, _I, _F, _I, _I, _F, _P, _P],
    "b2n_membench_read": [_P, _L, _I, _P, _P],
    "b2n_membench_gather": [_P, _L, _I, _P, C.POINTER(C.c_int64), _P],
    "b2n_ssi_depth_loss_fwbw": [_P, _P, _L, _F, _F, _P, _P, _P, _P, _P],
    "b2n_nerf_loss_fwbw": [_P, _P, _P, _L, _F, _F, _F, _P, _P, _P, _P, _P, _P],
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"libb2n.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.b2n_last_error.restype = C.c_char_p
        _lib.b2n_version.restype = C.c_int
        for name, sig in _SIGS.items():
            fn = getattr(_lib, name)
            fn.argtypes = sig
            fn.restype = C.c_int
    return _lib


def exported_symbols():
    return ["b2n_version", "b2n_last_error", *_SIGS.keys()]


def ptr(t):
    """Device (or host) pointer of a tensor; None -> NULL."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Invoke a b2n entry point on torch's current stream; raise RuntimeError on a non-zero status."""
    l = lib()
    rc = getattr(l, name)(*args, stream())
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {l.b2n_last_error().decode()}")


def call_nostream(name, *args):
    l = lib()
    rc = getattr(l, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed ({rc}): {l.b2n_last_error().decode()}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("google-nerf_b200 kernels need CUDA tensors (there is no CPU fallback)")
