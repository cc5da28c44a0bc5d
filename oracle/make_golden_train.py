"""Generates the round-2 fixtures by EXECUTING the unmodified reference on CPU (TEST INFRASTRUCTURE ONLY).

    python oracle/make_golden_train.py trajectory   # tests/golden/train_trajectory.npz
    python oracle/make_golden_train.py fullsize     # tests/golden/fullsize_cfg3.npz, fullsize_cfg5.npz
    python oracle/make_golden_train.py trained      # tests/golden/trained_ckpt.npz (briefly trained on data/1, + its render)

trajectory: 30 optimisation steps of the reference's own training body (render -> img2mse(rgb) + img2mse(rgb0) -> backward ->
            Adam, DS_NeRF/run.py:914-1029 without the guidance terms) on 256 fixed rays with pytest=True random streams: the
            loss curve, the rendered colours along the way and the parameter update of every tensor.
fullsize  : BASELINE cfg 3 (1008 x 756 view) on a 4096-ray strided subset, and one full 512 x 512 guidance view of cfg 5
            (rgb / disp / acc / depth on a strided subset, the FULL depth map, and the normal map depth2normal_geo makes of it).
trained   : the reference NeRF trained for a few hundred steps on the real data/1 images (poses 40.., as load_llff does), so
            that density has structure (the random-init fixtures all have sigma ~ 0 at the far sample); stores the weights
            and the reference's render of a strided ray set of a training view.
"""
import argparse
import importlib
import os
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import  # noqa: E402
from oracle.make_golden import load_seeded, nerf_args, sub_grad, synth_rays  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
H3, W3, FOCAL3, NEAR, FAR = 756, 1008, 767.2935, 1.2, 7.7369


def make_nets(run, seeds):
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "exp"))
        torch.manual_seed(0)
        kw_train, kw_test, _, grad_vars, optimizer = run.create_nerf(nerf_args(td, "exp"))
    load_seeded(kw_train["network_fn"], seeds[0])
    load_seeded(kw_train["network_fine"], seeds[1])
    return kw_train, kw_test, grad_vars, optimizer


def far_sigma(kw, ro, rd, far, net):
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    pts = (ro + rd * far)[:, None, :]
    with torch.no_grad():
        return kw["network_query_fn"](pts, vd, net)[:, 0, 3].numpy()


def smooth_target(rd):
    d = rd / torch.norm(rd, dim=-1, keepdim=True)
    return torch.stack([0.5 + 0.45 * torch.sin(9 * d[:, 0]), 0.5 + 0.45 * torch.cos(7 * d[:, 1]),
                        0.5 + 0.45 * torch.sin(5 * d[:, 0] + 4 * d[:, 1])], -1)


def trajectory(run, helpers, steps=30, n_rays=256):
    kw_train, _, grad_vars, optimizer = make_nets(run, (11, 12))
    nets = {"coarse": kw_train["network_fn"], "fine": kw_train["network_fine"]}
    init = {nm: {k.replace("module.", ""): v.detach().clone() for k, v in net.named_parameters()} for nm, net in nets.items()}
    ro, rd = synth_rays(helpers, n_rays, seed=5)
    target = smooth_target(rd)
    fx = {"coarse_seed": np.int64(11), "fine_seed": np.int64(12), "rays_o": ro.numpy(), "rays_d": rd.numpy(),
          "target": target.numpy(), "lrate": np.float64(5e-4), "near": np.float64(NEAR), "far": np.float64(FAR)}
    losses, rgbs = [], []
    t0 = time.time()
    for it in range(steps):
        rgb, disp, acc, depth, extras = run.render(H3, W3, FOCAL3, chunk=32768, rays=torch.stack([ro, rd], 0), near=NEAR, far=FAR,
                                                   pytest=True, **kw_train)
        optimizer.zero_grad()
        loss = helpers.img2mse(rgb, target) + helpers.img2mse(extras["rgb0"], target)       # run.py:1000, 1024-1026
        loss.backward()
        optimizer.step()
        losses.append(float(loss.item()))
        if it in (0, 9, 19, steps - 1):
            rgbs.append(rgb.detach().numpy().copy())
        print("step %d loss %.6f (%.1fs)" % (it, losses[-1], time.time() - t0), flush=True)
    fx["loss"] = np.array(losses, np.float64)
    fx["rgb_steps"] = np.array([0, 9, 19, steps - 1])
    fx["rgb"] = np.stack(rgbs, 0)
    for nm, net in nets.items():
        for k, v in net.named_parameters():
            k = k.replace("module.", "")
            fx["delta.%s.%s" % (nm, k)] = sub_grad((v.detach() - init[nm][k]).numpy())
    path = os.path.join(OUT, "train_trajectory.npz")
    np.savez_compressed(path, **fx)
    print(path, os.path.getsize(path))


def fullsize(run, helpers):
    kw_train, kw_test, _, _ = make_nets(run, (200, 201))
    # ---- cfg 3: 4096 strided rays of the 1008 x 756 identity-pose view ------------------------------------------
    c2w = torch.eye(4)[:3, :4]
    ro, rd = helpers.get_rays(H3, W3, FOCAL3, c2w)
    idx = torch.arange(4096) * 186 + 55
    ro_s, rd_s = ro.reshape(-1, 3)[idx].contiguous(), rd.reshape(-1, 3)[idx].contiguous()
    t0 = time.time()
    with torch.no_grad():
        rgb, disp, acc, depth, extras = run.render(H3, W3, FOCAL3, chunk=32768, rays=torch.stack([ro_s, rd_s], 0), near=NEAR,
                                                   far=FAR, retraw=True, **kw_test)
    fx = {"coarse_seed": np.int64(200), "fine_seed": np.int64(201), "idx": idx.numpy(), "H": np.int64(H3), "W": np.int64(W3),
          "focal": np.float64(FOCAL3), "near": np.float64(NEAR), "far": np.float64(FAR),
          "rgb": rgb.numpy(), "disp": disp.numpy(), "acc": acc.numpy(), "depth": depth.numpy(), "rgb0": extras["rgb0"].numpy(),
          "acc0": extras["acc0"].numpy(), "sigma_far_fine": extras["raw"][:, -1, 3].numpy(),
          "sigma_far_coarse": far_sigma(kw_test, ro_s, rd_s, FAR, kw_test["network_fn"])}
    path = os.path.join(OUT, "fullsize_cfg3.npz")
    np.savez_compressed(path, **fx)
    print(path, os.path.getsize(path), "%.1fs" % (time.time() - t0), flush=True)

    # ---- cfg 5: one full 512 x 512 guidance view + its normal map (run.py:948-965) ---------------------------------
    Hg = Wg = 512
    fg = FOCAL3 * Wg / W3
    pose = torch.eye(4)[:3, :4].clone()
    pose[0, 3] = 0.1
    t0 = time.time()
    with torch.no_grad():
        rgb, disp, acc, depth, extras = run.render(Hg, Wg, fg, chunk=32768, c2w=pose, near=NEAR, far=FAR, retraw=True, **kw_test)
        K = torch.Tensor([[fg, 0, Wg / 2], [0, fg, Hg / 2], [0, 0, 1]])
        xyz = run.depth2xyz_torch(depth, K)
        normal = run.depth2normal_geo(xyz.permute(2, 0, 1).unsqueeze(0))[0]       # [3,H,W], not normalised
    ro, rd = helpers.get_rays(Hg, Wg, fg, pose)
    sfc = far_sigma(kw_test, ro.reshape(-1, 3), rd.reshape(-1, 3), FAR, kw_test["network_fn"]).reshape(Hg, Wg)
    st = 8
    fx = {"coarse_seed": np.int64(200), "fine_seed": np.int64(201), "c2w": pose.numpy(), "H": np.int64(Hg), "W": np.int64(Wg),
          "focal": np.float64(fg), "near": np.float64(NEAR), "far": np.float64(FAR), "stride": np.int64(st),
          "rgb": rgb[::st, ::st].numpy(), "disp": disp[::st, ::st].numpy(), "acc": acc[::st, ::st].numpy(),
          "rgb0": extras["rgb0"][::st, ::st].numpy(),
          "depth_full": depth.numpy(), "normal": normal[:, ::st, ::st].numpy(), "normal_rows": normal[:, 250:258, :].numpy(),
          "sigma_far_fine": extras["raw"][..., -1, 3].numpy()[::st, ::st], "sigma_far_coarse": sfc[::st, ::st],
          "frac_ill_fine_full": np.float64((np.abs(extras["raw"][..., -1, 3].numpy()) <= 5e-3).mean()),
          "frac_ill_coarse_full": np.float64((np.abs(sfc) <= 5e-3).mean())}
    path = os.path.join(OUT, "fullsize_cfg5.npz")
    np.savez_compressed(path, **fx)
    print(path, os.path.getsize(path), "%.1fs" % (time.time() - t0), flush=True)


def trained(run, helpers, steps=400, n_rand=1024):
    from PIL import Image
    llff = importlib.import_module("load_llff")
    root = os.path.join(ref_import.REF_ROOT, "data", "1")
    arr = np.load(os.path.join(root, "poses_bounds.npy"))
    poses = arr[:, :-2].reshape([-1, 3, 5]).transpose([1, 2, 0])
    bds = arr[:, -2:].transpose([1, 0])
    Hh, Ww = 567, 1008
    poses[:2, 4, :] = np.array([Hh, Ww]).reshape([2, 1])
    poses[2, 4, :] = poses[2, 4, :] * 1. / 4
    poses = np.concatenate([poses[:, 1:2, :], -poses[:, 0:1, :], poses[:, 2:, :]], 1)
    poses = np.moveaxis(poses, -1, 0).astype(np.float32)
    bds = np.moveaxis(bds, -1, 0).astype(np.float32)
    sc = 1. / (bds.min() * .75)
    poses[:, :3, 3] *= sc
    bds *= sc
    poses = llff.recenter_poses(poses).astype(np.float32)
    focal = float(poses[0, 2, -1])
    near, far = float(bds.min() * .9), float(bds.max() * 1.)
    files = sorted(f for f in os.listdir(os.path.join(root, "images_4")) if f.endswith(".png"))
    assert len(files) == poses.shape[0], (len(files), poses.shape)
    views = list(range(40, poses.shape[0], 6))                   # every 6th training view
    imgs = [torch.from_numpy(np.asarray(Image.open(os.path.join(root, "images_4", files[v])).convert("RGB"), np.float32) / 255.)
            for v in views]
    rays = [helpers.get_rays(Hh, Ww, focal, torch.from_numpy(poses[v, :3, :4].copy())) for v in views]
    ro_all = torch.cat([r[0].reshape(-1, 3) for r in rays], 0)
    rd_all = torch.cat([r[1].reshape(-1, 3) for r in rays], 0)
    rgb_all = torch.cat([im.reshape(-1, 3) for im in imgs], 0)
    print("views", len(views), "rays", ro_all.shape[0], "focal", focal, "near/far", near, far, flush=True)

    kw_train, kw_test, grad_vars, optimizer = make_nets(run, (300, 301))
    g = torch.Generator().manual_seed(7)
    torch.manual_seed(7)
    t0 = time.time()
    losses = []
    for it in range(steps):
        sel = torch.randint(0, ro_all.shape[0], (n_rand,), generator=g)
        rgb, disp, acc, depth, extras = run.render(Hh, Ww, focal, chunk=32768, rays=torch.stack([ro_all[sel], rd_all[sel]], 0),
                                                   near=near, far=far, **kw_train)
        optimizer.zero_grad()
        loss = helpers.img2mse(rgb, rgb_all[sel]) + helpers.img2mse(extras["rgb0"], rgb_all[sel])
        loss.backward()
        optimizer.step()
        losses.append(float(loss.item()))
        if it % 20 == 0:
            print("step %d loss %.5f (%.0fs)" % (it, losses[-1], time.time() - t0), flush=True)
    fx = {"H": np.int64(Hh), "W": np.int64(Ww), "focal": np.float64(focal), "near": np.float64(near), "far": np.float64(far),
          "train_loss": np.array(losses), "c2w": poses[views[1], :3, :4].copy()}
    for nm, net in (("coarse", kw_train["network_fn"]), ("fine", kw_train["network_fine"])):
        for k, v in net.state_dict().items():
            fx["%s.%s" % (nm, k.replace("module.", ""))] = v.detach().numpy().copy()
    ro, rd = rays[1]
    ro, rd = ro[::9, ::15].reshape(-1, 3).contiguous(), rd[::9, ::15].reshape(-1, 3).contiguous()
    with torch.no_grad():
        rgb, disp, acc, depth, extras = run.render(Hh, Ww, focal, chunk=32768, rays=torch.stack([ro, rd], 0), near=near, far=far,
                                                   retraw=True, **kw_test)
    fx.update({"rays_o": ro.numpy(), "rays_d": rd.numpy(), "rgb": rgb.numpy(), "disp": disp.numpy(), "acc": acc.numpy(),
               "depth": depth.numpy(), "rgb0": extras["rgb0"].numpy(), "acc0": extras["acc0"].numpy(),
               "sigma_far_fine": extras["raw"][:, -1, 3].numpy(),
               "sigma_far_coarse": far_sigma(kw_test, ro, rd, far, kw_test["network_fn"]),
               "target": imgs[1][::9, ::15].reshape(-1, 3).numpy()})
    path = os.path.join(OUT, "trained_ckpt.npz")
    np.savez_compressed(path, **fx)
    print(path, os.path.getsize(path), "n rays", ro.shape[0], "acc mean", float(acc.mean()),
          "psnr vs image %.2f" % float(-10 * np.log10(((rgb.numpy() - fx["target"]) ** 2).mean())))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["trajectory", "fullsize", "trained"])
    ap.add_argument("--steps", type=int, default=None)
    a = ap.parse_args()
    run_mod, helpers_mod = ref_import.load()
    if a.what == "trajectory":
        trajectory(run_mod, helpers_mod, steps=a.steps or 30)
    elif a.what == "fullsize":
        fullsize(run_mod, helpers_mod)
    else:
        trained(run_mod, helpers_mod, steps=a.steps or 400)
