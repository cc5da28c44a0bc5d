"""TEST INFRASTRUCTURE ONLY — ctypes loader of the plain-C oracle (oracle/nerf_oracle_c.c, built by oracle/Makefile or
__graft_entry__.build()).  Same stages as the bit-exact part of oracle/nerf_oracle.py, written without numpy / torch."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "build", "libnerf_oracle_c.so")
_lib = None


def build():
    subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def ray_batch(H, W, focal, c2w, near, far, use_viewdirs=True, c2w_static=None, patch=None, ndc=False, ndc_near=1.):
    c2w = _f(np.asarray(c2w)[:3, :4])
    cs = None if c2w_static is None else _f(np.asarray(c2w_static)[:3, :4])
    i0, j0, h, w = (0, 0, H, W) if patch is None else [int(v) for v in patch]
    out = np.empty((h * w, 11 if use_viewdirs else 8), np.float32)
    load().mvo_ray_batch(_p(c2w), _p(cs), int(H), int(W), ctypes.c_double(focal), ctypes.c_float(near), ctypes.c_float(far), i0, j0,
                         h, w, int(use_viewdirs), int(ndc), ctypes.c_double(ndc_near), _p(out))
    return out


def sample_coarse(rays, t_vals, t_rand=None, lindisp=False):
    rays, t_vals = _f(rays), _f(t_vals)
    t_rand = None if t_rand is None else _f(t_rand)
    z = np.empty((rays.shape[0], t_vals.size), np.float32)
    load().mvo_sample_coarse(_p(rays), rays.shape[1], ctypes.c_int64(rays.shape[0]), _p(t_vals), _p(t_rand), t_vals.size,
                             int(lindisp), _p(z))
    return z


def fine_samples(z_vals, weights, u):
    z_vals, weights, u = _f(z_vals), _f(weights), _f(u)
    n, S = z_vals.shape
    M = u.shape[-1]
    samples = np.empty((n, M), np.float32)
    inds = np.empty((n, M), np.int64)
    cdf = np.empty((n, S - 1), np.float32)
    merged = np.empty((n, S + M), np.float32)
    load().mvo_sample_fine(_p(z_vals), _p(weights), _p(u), int(u.ndim == 1), ctypes.c_int64(n), S, M, _p(samples), _p(inds),
                           _p(cdf), _p(merged))
    return {"z_samples": samples, "inds": inds, "cdf": cdf, "z_merged": merged}


def raw2outputs(raw, z_vals, rays_d, noise=None, white_bkgd=False):
    raw, z_vals, rays_d = _f(raw), _f(z_vals), _f(rays_d)
    noise = None if noise is None else _f(noise)
    n, S = z_vals.shape
    rgb, disp, acc, depth = np.empty((n, 3), np.float32), np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.float32)
    weights, alpha = np.empty((n, S), np.float32), np.empty((n, S), np.float32)
    load().mvo_raw2outputs(_p(raw), _p(z_vals), _p(rays_d), _p(noise), ctypes.c_int64(n), S, int(white_bkgd), _p(rgb), _p(disp),
                           _p(acc), _p(depth), _p(weights), _p(alpha))
    return {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, "depth_map": depth, "weights": weights, "alpha": alpha}
