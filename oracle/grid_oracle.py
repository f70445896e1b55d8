"""numpy front-end of oracle/grid_oracle.c (TEST INFRASTRUCTURE — see the header of the C file)."""
import ctypes as C

import numpy as np

from . import build_oracle

_lib = None


def _l():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_oracle.build())
        _lib.oracle_smoothness.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def affine_base(n):
    """linspace(-1,1,n)*(n-1)/n exactly as ATen builds it (fp32 ops in this order)."""
    import torch
    return (torch.linspace(-1, 1, n) * (n - 1) / n).numpy()


def affine_grid(theta, h, w):
    theta = _f(theta)
    n = theta.shape[0]
    grid = np.empty((n, h, w, 2), np.float32)
    _l().oracle_affine_grid(_p(theta), _p(_f(affine_base(w))), _p(_f(affine_base(h))), n, h, w, _p(grid))
    return grid


def flow_grid(off_nchw):
    import torch
    off = _f(off_nchw)
    n, _, h, w = off.shape
    grid = np.empty((n, h, w, 2), np.float32)
    _l().oracle_flow_grid(_p(off), _p(_f(torch.linspace(-1, 1, w).numpy())), _p(_f(torch.linspace(-1, 1, h).numpy())), n, h,
                          w, _p(grid))
    return grid


def grid_sample_fwd(img, grid):
    img, grid = _f(img), _f(grid)
    n, c, h, w = img.shape
    ho, wo = grid.shape[1:3]
    out = np.empty((n, c, ho, wo), np.float32)
    idx = np.empty((n, ho, wo, 2), np.int32)
    _l().oracle_grid_sample_fwd(_p(img), n, c, h, w, _p(grid), ho, wo, _p(out), _p(idx))
    return out, idx


def grid_sample_bwd(img, grid, dout, need_dimg=True):
    img, grid, dout = _f(img), _f(grid), _f(dout)
    n, c, h, w = img.shape
    ho, wo = grid.shape[1:3]
    dimg = np.zeros_like(img) if need_dimg else None
    dgrid = np.empty((n, ho, wo, 2), np.float32)
    _l().oracle_grid_sample_bwd(_p(img), n, c, h, w, _p(grid), ho, wo, _p(dout), _p(dimg), _p(dgrid))
    return dimg, dgrid


def smoothness(def_nchw, img=None, alpha=0.0):
    d = _f(def_nchw)
    n, _, h, w = d.shape
    im = _f(img) if img is not None else None
    return float(_l().oracle_smoothness(_p(d), _p(im), im.shape[1] if im is not None else 0, C.c_float(alpha), n, h, w))
