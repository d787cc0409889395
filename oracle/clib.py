"""ctypes loader for oracle/c/vren_oracle.c -- TEST INFRASTRUCTURE (see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "c", "vren_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _c(t, dtype):
    return t.detach().to(dtype).contiguous()


def morton3D(coords):
    c = _c(coords, torch.int32); out = torch.empty(c.shape[0], dtype=torch.int32)
    lib().orc_morton3D(_p(c), C.c_int64(c.shape[0]), _p(out)); return out


def morton3D_invert(idx):
    i = _c(idx, torch.int32); out = torch.empty(i.shape[0], 3, dtype=torch.int32)
    lib().orc_morton3D_invert(_p(i), C.c_int64(i.shape[0]), _p(out)); return out


def packbits(grid, thr, bitfield):
    g = _c(grid, torch.float32)
    lib().orc_packbits(_p(g), C.c_int64(bitfield.numel()), C.c_float(thr), _p(bitfield)); return bitfield


def ray_aabb_intersect(o, d, c, h, max_hits, sphere=False):
    o, d, c, h = (_c(v, torch.float32) for v in (o, d, c, h))
    n = o.shape[0]
    cnt = torch.empty(n, dtype=torch.int32); ht = torch.empty(n, max_hits, 2); hv = torch.empty(n, max_hits, dtype=torch.int64)
    fn = lib().orc_ray_sphere_intersect if sphere else lib().orc_ray_aabb_intersect
    fn(_p(o), _p(d), _p(c), _p(h), C.c_int64(n), C.c_int64(c.shape[0]), C.c_int(max_hits), _p(cnt), _p(ht), _p(hv))
    return cnt, ht, hv


def raymarching_train(o, d, hits_t, bitfield, cascades, scale, esf, noise, grid_size, max_samples):
    o, d, hits_t, noise = (_c(v, torch.float32) for v in (o, d, hits_t, noise))
    n = o.shape[0]
    rays_a = torch.empty(n, 3, dtype=torch.int64); counter = torch.zeros(2, dtype=torch.int32)
    args = lambda xs, ds, dl, ts: (_p(o), _p(d), _p(hits_t), _p(bitfield), C.c_int(cascades), C.c_float(scale),
                                   C.c_float(esf), _p(noise), C.c_int(grid_size), C.c_int(max_samples), C.c_int64(n),
                                   _p(rays_a), _p(xs), _p(ds), _p(dl), _p(ts), _p(counter))
    lib().orc_raymarching_train(*args(None, None, None, None))
    N = int(counter[0])
    xyzs = torch.empty(N, 3); dirs = torch.empty(N, 3); deltas = torch.empty(N); ts = torch.empty(N)
    lib().orc_raymarching_train(*args(xyzs, dirs, deltas, ts))
    return rays_a, xyzs, dirs, deltas, ts, counter


def raymarching_test(o, d, hits_t, alive, bitfield, cascades, scale, esf, grid_size, max_samples, n_samples):
    o, d = _c(o, torch.float32), _c(d, torch.float32)
    assert hits_t.dtype == torch.float32 and hits_t.is_contiguous()
    a = alive.shape[0]
    xyzs = torch.empty(a, n_samples, 3); dirs = torch.empty(a, n_samples, 3)
    deltas = torch.empty(a, n_samples); ts = torch.empty(a, n_samples); neff = torch.empty(a, dtype=torch.int32)
    lib().orc_raymarching_test(_p(o), _p(d), _p(hits_t), _p(alive), _p(bitfield), C.c_int(cascades), C.c_float(scale),
                               C.c_float(esf), C.c_int(grid_size), C.c_int(max_samples), C.c_int(n_samples),
                               C.c_int64(a), _p(xyzs), _p(dirs), _p(deltas), _p(ts), _p(neff))
    return xyzs, dirs, deltas, ts, neff


def composite_train_fw(sigmas, rgbs, deltas, ts, rays_a, T_threshold):
    sigmas, rgbs, deltas, ts = (_c(v, torch.float32) for v in (sigmas, rgbs, deltas, ts))
    n = rays_a.shape[0]
    op = torch.zeros(n); dp = torch.zeros(n); d2 = torch.zeros(n); rgb = torch.zeros(n, 3)
    lib().orc_composite_train_fw(_p(sigmas), _p(rgbs), _p(deltas), _p(ts), _p(rays_a), C.c_float(T_threshold),
                                 C.c_int64(n), _p(op), _p(dp), _p(d2), _p(rgb))
    return op, dp, d2, rgb


def composite_train_bw(gO, gD, gD2, gRGB, sigmas, rgbs, deltas, ts, rays_a, op, dp, d2, rgb, T_threshold):
    f = lambda v: _c(v, torch.float32)
    sigmas, rgbs, deltas, ts = f(sigmas), f(rgbs), f(deltas), f(ts)
    n, N = rays_a.shape[0], sigmas.shape[0]
    ds = torch.empty(N); dc = torch.empty(N, 3)
    lib().orc_composite_train_bw(_p(f(gO)), _p(f(gD)), _p(f(gD2)), _p(f(gRGB)), _p(sigmas), _p(rgbs), _p(deltas),
                                 _p(ts), _p(rays_a), _p(f(op)), _p(f(dp)), _p(f(d2)), _p(f(rgb)),
                                 C.c_float(T_threshold), C.c_int64(n), C.c_int64(N), _p(ds), _p(dc))
    return ds, dc


def composite_test_fw(sigmas, rgbs, deltas, ts, hits_t, alive, T_threshold, neff, opacity, depth, rgb):
    f = lambda v: _c(v, torch.float32)
    a, s = sigmas.shape
    lib().orc_composite_test_fw(_p(f(sigmas)), _p(f(rgbs)), _p(f(deltas)), _p(f(ts)), _p(hits_t), _p(alive),
                                C.c_float(T_threshold), _p(neff), C.c_int(s), C.c_int64(a), _p(opacity), _p(depth), _p(rgb))
