"""Generates tests/golden/ray_batch.npz by EXECUTING the unmodified reference's render() on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference exists):

    python oracle/make_golden_rays.py

render() (DS_NeRF/run.py:1143-1219) is called with `batchify_rays` replaced by a recorder, so the fixture is the
exact [N, 8|11] ray batch the reference assembles (get_rays, patch crop, c2w_staticcam, viewdir normalisation,
ndc_rays, near/far columns, cat) for both entry forms (`c2w=` and `rays=`).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = [   # name, H, W, focal, near, far, ndc, use_viewdirs, static, patch
    ("plain", 37, 53, 41.7, 1.2, 7.7369, False, True, False, None),
    ("ndc", 37, 53, 41.7, 0.0, 1.0, True, True, False, None),
    ("ndc_static", 37, 53, 41.7, 0.0, 1.0, True, True, True, None),
    ("ndc_patch", 37, 53, 41.7, 0.0, 1.0, True, True, False, (3, 5, 20, 31)),
    ("ndc_nodirs", 24, 40, 767.2935 / 16, 0.0, 1.0, True, False, False, None),
    ("static", 37, 53, 41.7, 1.2, 7.7369, False, True, True, None),   # (static + patch raises in the reference: run.py:1207)
    ("patch", 37, 53, 41.7, 1.2, 7.7369, False, True, False, (3, 5, 20, 31)),
]


def pose(rng):
    """forward-facing pose (camera looks down -z, small rotation): d.z stays away from 0 as ndc_rays needs"""
    a = rng.randn(3) * 0.15
    rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    rz = np.array([[np.cos(a[2]), -np.sin(a[2]), 0], [np.sin(a[2]), np.cos(a[2]), 0], [0, 0, 1]])
    t = rng.randn(3, 1) * 0.3
    return torch.from_numpy(np.concatenate([rx @ ry @ rz, t], 1).astype(np.float32))


def main():
    run, helpers = ref_import.load()
    rng = np.random.RandomState(11)
    fx = {"cases": np.array([c[0] for c in CASES])}
    captured = {}

    def recorder(rays_flat, chunk=0, **kw):
        captured["rays"] = rays_flat.detach().clone()
        n = rays_flat.shape[0]
        return {"rgb_map": torch.zeros(n, 3), "disp_map": torch.zeros(n), "acc_map": torch.zeros(n), "depth_map": torch.zeros(n)}

    orig = run.batchify_rays
    run.batchify_rays = recorder
    try:
        for name, H, W, focal, near, far, ndc, use_vd, static, patch in CASES:
            c2w, c2s = pose(rng), (pose(rng) if static else None)
            run.render(H, W, focal, c2w=c2w, ndc=ndc, near=near, far=far, use_viewdirs=use_vd, c2w_staticcam=c2s, patch=patch)
            fx[name + "_c2w"] = c2w.numpy()
            if static:
                fx[name + "_c2w_static"] = c2s.numpy()
            fx[name + "_args"] = np.array([H, W, focal, near, far, int(ndc), int(use_vd)] + list(patch or (0, 0, H, W)), dtype=np.float64)
            fx[name + "_batch"] = captured["rays"].numpy()
            if not static:    # the `rays=` entry on a seeded subset of the same view (training form, run.py:914)
                ro, rd = helpers.get_rays(H, W, focal, c2w)
                idx = torch.from_numpy(rng.permutation(H * W)[:200])
                ro, rd = ro.reshape(-1, 3)[idx].contiguous(), rd.reshape(-1, 3)[idx].contiguous()
                run.render(H, W, focal, rays=torch.stack([ro, rd], 0), ndc=ndc, near=near, far=far, use_viewdirs=use_vd)
                fx[name + "_rays_o"], fx[name + "_rays_d"] = ro.numpy(), rd.numpy()
                fx[name + "_rays_batch"] = captured["rays"].numpy()
    finally:
        run.batchify_rays = orig
    path = os.path.join(OUT, "ray_batch.npz")
    np.savez_compressed(path, **fx)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
