"""Generates tests/golden/cfg1_view.npz by EXECUTING the unmodified reference on CPU for BASELINE cfg 1:
config_1.txt on data/1 (factor 4), random-init NeRF, render-only forward of one training view.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference exists):

    python oracle/make_golden_cfg1.py

load_llff_data cannot run as shipped (data/1/images/ and sparse/0/*.bin are absent), so the pose pipeline
of load_llff.py is replayed on data/1/poses_bounds.npy with the reference's OWN functions (recenter_poses / poses_avg):
[-u, r, -t] -> [r, u, -t] reorder (load_llff.py:323-325), sc = 1/(bds.min() * 0.75) (:337-339), recenter (:343),
training poses = poses[40:] (:427), hwf from the images_4 size 567 x 1008 and focal / factor (:73-74, 123-125),
near = bds.min() * .9, far = bds.max() (run.py:417-418, no_ndc).  The view is rendered with render_kwargs_test on a
16 x 24 patch and on a strided set of rays through the reference's render(); inputs + outputs are stored.
"""
import importlib
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import  # noqa: E402
from oracle.make_golden import load_seeded, nerf_args  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    run, helpers = ref_import.load()
    llff = importlib.import_module("load_llff")
    arr = np.load(os.path.join(ref_import.REF_ROOT, "data", "1", "poses_bounds.npy"))
    poses = arr[:, :-2].reshape([-1, 3, 5]).transpose([1, 2, 0])
    bds = arr[:, -2:].transpose([1, 0])
    factor, H, W = 4, 567, 1008                                  # images_4/*.png
    poses[:2, 4, :] = np.array([H, W]).reshape([2, 1])
    poses[2, 4, :] = poses[2, 4, :] * 1. / factor
    poses = np.concatenate([poses[:, 1:2, :], -poses[:, 0:1, :], poses[:, 2:, :]], 1)
    poses = np.moveaxis(poses, -1, 0).astype(np.float32)
    bds = np.moveaxis(bds, -1, 0).astype(np.float32)
    sc = 1. / (bds.min() * .75)
    poses[:, :3, 3] *= sc
    bds *= sc
    poses = llff.recenter_poses(poses).astype(np.float32)
    hwf = poses[0, :3, -1]
    train_poses = poses[40:, :, :]
    focal = float(hwf[2])
    near, far = float(np.ndarray.min(bds) * .9), float(np.ndarray.max(bds) * 1.)
    print("hwf", hwf, "near/far", near, far)

    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "exp"))
        torch.manual_seed(0)
        kw_train, kw_test, _, _, _ = run.create_nerf(nerf_args(td, "exp"))
    load_seeded(kw_train["network_fn"], 200)
    load_seeded(kw_train["network_fine"], 201)
    c2w = torch.from_numpy(train_poses[0, :3, :4].copy())
    fx = {"coarse_seed": np.int64(200), "fine_seed": np.int64(201), "c2w": c2w.numpy(), "H": np.int64(H), "W": np.int64(W),
          "focal": np.float64(focal), "near": np.float64(near), "far": np.float64(far)}
    patch = (270, 492, 16, 24)
    fx["patch"] = np.array(patch)

    def far_sigma(ro, rd, net):
        """density the reference's network predicts at the LAST sample of each ray (z = far).  raw2outputs gives that sample
        dist = 1e10, so alpha_last jumps 0 -> 1 where this value crosses zero: rays with |sigma_far| below the bf16 error
        of the MLP are ill-conditioned for ANY reduced-precision implementation and are excluded from comparisons."""
        vd = rd / torch.norm(rd, dim=-1, keepdim=True)
        pts = (ro + rd * far)[:, None, :]
        return kw_test["network_query_fn"](pts, vd, net)[:, 0, 3].numpy()

    with torch.no_grad():
        rgb, disp, acc, depth, extras = run.render(H, W, focal, chunk=32768, c2w=c2w, patch=patch, near=near, far=far,
                                                   retraw=True, **kw_test)
        fx["patch_rgb"], fx["patch_disp"], fx["patch_acc"], fx["patch_depth"] = [t.numpy() for t in (rgb, disp, acc, depth)]
        fx["patch_rgb0"], fx["patch_z_vals"] = extras["rgb0"].numpy(), extras["z_vals"].numpy()
        fx["patch_sigma_far_fine"] = extras["raw"][..., -1, 3].numpy()
        ro_p, rd_p = helpers.get_rays(H, W, focal, c2w)
        i0, j0, h, w = patch
        ro_p, rd_p = ro_p[i0:i0 + h, j0:j0 + w].reshape(-1, 3), rd_p[i0:i0 + h, j0:j0 + w].reshape(-1, 3)
        fx["patch_sigma_far_coarse"] = far_sigma(ro_p, rd_p, kw_test["network_fn"]).reshape(h, w)
        # a strided set of rays of the whole view through the `rays=` entry
        ro, rd = helpers.get_rays(H, W, focal, c2w)
        ro, rd = ro[::40, ::41].reshape(-1, 3).contiguous(), rd[::40, ::41].reshape(-1, 3).contiguous()
        rgb, disp, acc, depth, extras = run.render(H, W, focal, chunk=32768, rays=torch.stack([ro, rd], 0), near=near, far=far,
                                                   retraw=True, **kw_test)
        fx["rays_sigma_far_fine"] = extras["raw"][..., -1, 3].numpy()
        fx["rays_sigma_far_coarse"] = far_sigma(ro, rd, kw_test["network_fn"])
        fx["rays_o"], fx["rays_d"] = ro.numpy(), rd.numpy()
        fx["rays_rgb"], fx["rays_disp"], fx["rays_acc"], fx["rays_depth"] = [t.numpy() for t in (rgb, disp, acc, depth)]
    path = os.path.join(OUT, "cfg1_view.npz")
    np.savez_compressed(path, **fx)
    print(path, os.path.getsize(path), "patch rgb mean", float(fx["patch_rgb"].mean()), "n strided rays", ro.shape[0])


if __name__ == "__main__":
    main()
