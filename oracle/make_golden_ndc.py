"""Generates tests/golden/render_ndc.npz by EXECUTING the unmodified reference on CPU in its forward-facing (NDC)
configuration: create_nerf with no_ndc=False (so render_kwargs carry ndc=True and no lindisp key), render(c2w=...) with the
default near=0 / far=1, linear-in-depth sampling, black background.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference exists):  python oracle/make_golden_ndc.py
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import  # noqa: E402
from oracle.make_golden import load_seeded, nerf_args  # noqa: E402
from oracle.make_golden_rays import pose  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    run, helpers = ref_import.load()
    a = nerf_args(None, "exp")
    a.no_ndc, a.white_bkgd, a.lindisp = False, False, False
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "exp"))
        a.basedir = td
        torch.manual_seed(0)
        kw_train, kw_test, _, _, _ = run.create_nerf(a)
    assert kw_test["ndc"] is True and "lindisp" not in kw_test
    load_seeded(kw_train["network_fn"], 200)
    load_seeded(kw_train["network_fine"], 201)
    H, W, focal = 24, 32, 28.0
    c2w = pose(np.random.RandomState(5))
    with torch.no_grad():
        rgb, disp, acc, depth, extras = run.render(H, W, focal, chunk=4096, c2w=c2w, retraw=True, **kw_test)
        ro, rd = helpers.get_rays(H, W, focal, c2w)
        vd = (rd / torch.norm(rd, dim=-1, keepdim=True)).reshape(-1, 3)
        ro, rd = helpers.ndc_rays(H, W, focal, 1., ro, rd)
        pts_far = (ro + rd * 1.0).reshape(-1, 1, 3)           # last coarse sample: z = near (1 - t) + far t = 1
        sigma_far_coarse = kw_test["network_query_fn"](pts_far, vd, kw_test["network_fn"])[:, 0, 3].reshape(H, W).numpy()
    fx = {"sigma_far_coarse": sigma_far_coarse, "coarse_seed": np.int64(200), "fine_seed": np.int64(201), "c2w": c2w.numpy(), "H": np.int64(H), "W": np.int64(W),
          "focal": np.float64(focal), "rgb": rgb.numpy(), "disp": disp.numpy(), "acc": acc.numpy(), "depth": depth.numpy(),
          "rgb0": extras["rgb0"].numpy(), "acc0": extras["acc0"].numpy(), "z_vals": extras["z_vals"].numpy(),
          "sigma_far_fine": extras["raw"][..., -1, 3].numpy()}
    path = os.path.join(OUT, "render_ndc.npz")
    np.savez_compressed(path, **fx)
    print(path, os.path.getsize(path), "acc range", float(acc.min()), float(acc.max()), "acc0", float(extras["acc0"].min()),
          float(extras["acc0"].max()), "sigma_far", float(fx["sigma_far_fine"].min()), float(fx["sigma_far_fine"].max()))


if __name__ == "__main__":
    main()
