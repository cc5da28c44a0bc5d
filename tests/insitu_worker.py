"""In-situ drop-in run (worker process; started by tests/test_gpu_insitu.py).

1. imports the UNMODIFIED reference `run` module (oracle/_ref or /root/reference) the way its __main__ runs it on a GPU
   (`torch.set_default_tensor_type('torch.cuda.FloatTensor')`, DS_NeRF/run.py:1978) and executes, with the reference's own
   functions, the render calls of ONE `train()` iteration (run.py:919, 960-965, 973-976, 978, 982, 1000-1029) plus
   `render_path(..., savedir=)` (run.py:1222) -> the stock-PyTorch-on-B200 result;
2. executes the python block of INTEGRATION.md §1 inside the reference module's namespace (exactly what a maintainer
   appends to run.py), rebuilds the networks through the rebound `create_nerf`, loads the stock weights, and executes the SAME
   reference code again — `render_path`, `render_path_4view` and the iteration body are still the reference's, every
   `render` / `depth2normal_geo` / `create_nerf` / `raw2outputs` / `sample_pdf` underneath is ours;
3. prints one JSON line with the measured differences.

Both passes see the same CUDA random stream (same seeds, same draw order and shapes).
"""
import argparse
import json
import os
import re
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def nerf_args(basedir):
    return argparse.Namespace(
        multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=64, N_samples=64, netdepth=8,
        netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=65536, alpha_model_path=None, no_coarse=False,
        lrate=5e-4, basedir=basedir, expname="insitu", ft_path=None, no_reload=True, perturb=1.0, white_bkgd=True,
        raw_noise_std=1.0, dataset_type="llff", no_ndc=True, lindisp=True, sigma_loss=False,
        chunk=32768, normalmap_render_factor=7, depth_lambda=0.1, sds_loss_weight=1e-4)


def integration_snippet():
    """the python block of INTEGRATION.md §1 (without its sys.path line: the repo root is already importable)"""
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n# at the end of DS_NeRF/run.py.*?\n(.*?)```", md, re.S).group(1)
    return "\n".join(ln for ln in block.splitlines() if not ln.startswith("import sys; sys.path.insert"))


def iteration(run, args, kw_train, kw_test, optimizer, data, savedir):
    """The render calls and the loss of one second-stage train() iteration (run.py:914-1029), written against the reference
    module's OWN global names (run.render, run.render_path_4view, run.depth2normal_geo, ...)."""
    H, W, focal = data["hwf"]
    out = {}
    i = 3
    # run.py:919  masked-region rays (rgb SDS branch)
    rgb, disp, _, _, _ = run.render(H, W, focal, chunk=args.chunk, rays=data["masked_batch_rays"], verbose=False, retraw=True, **kw_train)
    # run.py:948-965  normal-map view
    H_r, W_r, focal_r = H // args.normalmap_render_factor, W // args.normalmap_render_factor, focal / args.normalmap_render_factor
    K = torch.Tensor(np.array([[focal_r, 0, W_r / 2], [0, focal_r, H_r / 2], [0, 0, 1]]))
    pose_i = data["poses"][i]
    _, _, _, depth1, _ = run.render(H_r, W_r, focal_r, chunk=args.chunk, c2w=pose_i[:3, :4], verbose=False, retraw=True, **kw_train)
    points = run.depth2xyz_torch(depth1.reshape(H_r, W_r), K)
    points_tensor = points.unsqueeze(0).transpose(2, 3).transpose(1, 2)
    normal = (run.depth2normal_geo(points_tensor) + 1) / 2
    # run.py:973-976  collaborative views (the reference's own render_path_4view)
    rgbs4, _, mask4 = run.render_path_4view(i, data["masks"], data["poses"], data["hwf"], args.chunk, kw_test,
                                            render_factor=args.normalmap_render_factor, need_alpha=True)
    # run.py:978, 982  unmasked RGB-D supervision
    rgb2, _, _, _, extras2 = run.render(H, W, focal, chunk=args.chunk, rays=data["batch_rays_clf"], verbose=False, retraw=True, **kw_train)
    _, disp2, _, _, _ = run.render(H, W, focal, chunk=args.chunk, rays=data["batch_inp"], verbose=False, retraw=True, **kw_train)
    # run.py:1000-1029 (the SDS term replaced by a fixed differentiable image functional of the same inputs)
    optimizer.zero_grad()
    img_loss = run.img2mse(rgb2, data["target_clf"])
    depth_loss = run.img2mse(disp2, data["target_inp"])
    loss = img_loss + args.depth_lambda * depth_loss
    loss = loss + run.img2mse(extras2["rgb0"], data["target_clf"])
    sds = ((rgb - 0.4) ** 2).mean() + ((normal - 0.5) ** 2).mean() + ((rgbs4 - 0.6) ** 2).mean()
    loss = loss + 100 * args.sds_loss_weight * sds
    loss.backward()
    grads = {}
    for nm in ("network_fn", "network_fine"):
        for k, v in kw_train[nm].named_parameters():
            grads[nm + "." + k.replace("module.", "")] = v.grad.detach().float().cpu().numpy().copy()
    optimizer.step()
    out.update(rgb=rgb, disp=disp, depth1=depth1, normal=normal, rgbs4=rgbs4, rgb2=rgb2, disp2=disp2, rgb0=extras2["rgb0"])
    out = {k: v.detach().float().cpu().numpy() for k, v in out.items()}
    out["loss"] = float(loss.item())
    # run.py:1222  render_path with the on-disk writers (reference code; test kwargs)
    with torch.no_grad():
        rgbs, disps, _ = run.render_path(data["poses"][:2], data["hwf"], args.chunk, kw_test, savedir=savedir, render_factor=6, need_alpha=True)
    out["path_rgbs"], out["path_disps"] = np.asarray(rgbs), np.asarray(disps)
    out["files"] = sorted(os.path.join(os.path.relpath(d, savedir), f) for d, _, fs in os.walk(savedir) for f in fs)
    # second iteration's loss after the Adam step
    with torch.no_grad():
        rgb3, _, _, _, ex3 = run.render(H, W, focal, chunk=args.chunk, rays=data["batch_rays_clf"], **kw_test)
    out["loss_after"] = float((run.img2mse(rgb3, data["target_clf"]) + run.img2mse(ex3["rgb0"], data["target_clf"])).item())
    return out, grads


def main():
    from oracle import ref_import
    torch.set_default_tensor_type('torch.cuda.FloatTensor')          # DS_NeRF/run.py:1978
    run, helpers = ref_import.load()
    sys.modules["imageio"].imwrite = lambda path, img: open(path, "wb").write(np.ascontiguousarray(img).tobytes()[:64])
    dev = torch.device("cuda")
    H, W, focal = 378, 504, 383.6
    g = torch.Generator(device="cpu").manual_seed(3)
    poses = []
    for v in range(8):
        p = torch.eye(4, device="cpu")
        p[0, 3], p[1, 3] = 0.05 * v, -0.02 * v
        poses.append(p)
    poses = torch.stack(poses, 0).to(dev)
    masks = np.zeros((8, H, W), np.float32)
    masks[:, 100:200, 150:300] = 1

    def pick(n):
        ro, rd = helpers.get_rays(H, W, focal, poses[3][:3, :4])
        idx = torch.randperm(H * W, generator=g, device="cpu")[:n].to(dev)
        return torch.stack([ro.reshape(-1, 3)[idx], rd.reshape(-1, 3)[idx]], 0)
    data = {"hwf": [H, W, focal], "poses": poses, "masks": masks, "masked_batch_rays": pick(1024), "batch_rays_clf": pick(1024),
            "batch_inp": pick(512), "target_clf": torch.rand(1024, 3, generator=g, device="cpu").to(dev),
            "target_inp": (torch.rand(512, generator=g, device="cpu") * 0.5).to(dev)}

    res = {}
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "insitu"))
        args = nerf_args(td)
        # ---- pass 1: the stock reference on this GPU (fp32) -------------------------------------------------------------
        torch.manual_seed(0)
        kw_train_ref, kw_test_ref, _, _, opt_ref = run.create_nerf(args)
        bds_dict = {"near": 1.2, "far": 7.7369}                         # run.py:554-559
        kw_train_ref.update(bds_dict)
        kw_test_ref.update(bds_dict)
        with torch.no_grad():       # a positive density offset, as a trained scene has: every ray hits something (acc > 0, finite disp)
            for k in ("network_fn", "network_fine"):
                kw_train_ref[k].module.alpha_linear.bias += 0.3
        state = {k: kw_train_ref[k].state_dict() for k in ("network_fn", "network_fine")}
        state = {k: {n: t.detach().clone() for n, t in sd.items()} for k, sd in state.items()}
        torch.manual_seed(1)
        os.makedirs(os.path.join(td, "ref_out"))
        ref_out, ref_grads = iteration(run, args, kw_train_ref, kw_test_ref, opt_ref, data, os.path.join(td, "ref_out"))
        stock_render = run.render
        # ---- pass 2: INTEGRATION.md §1 applied to the reference module --------------------------------------------------
        exec(integration_snippet(), run.__dict__)
        assert run.render is not stock_render and run.render.__module__ == "mvip_nerf_b200.run"
        torch.manual_seed(0)
        kw_train, kw_test, _, _, opt = run.create_nerf(args)            # ours now (same call, same Namespace)
        kw_train.update(bds_dict)
        kw_test.update(bds_dict)
        assert type(opt).__name__ == "FusedAdam"
        for k in ("network_fn", "network_fine"):
            kw_train[k].load_state_dict(state[k])                       # `module.`-prefixed keys of the stock nn.DataParallel
        from mvip_nerf_b200 import ops
        l0 = ops.launch_count
        torch.manual_seed(1)
        os.makedirs(os.path.join(td, "our_out"))
        our_out, our_grads = iteration(run, args, kw_train, kw_test, opt, data, os.path.join(td, "our_out"))
        res["launches"] = ops.launch_count - l0

    def cmp(key):
        a, b = ref_out[key], our_out[key]
        fin = np.isfinite(a) & np.isfinite(b)
        d = np.abs(a - b)[fin]
        return {"max_abs": float(d.max()) if d.size else None, "mean_abs": float(d.mean()) if d.size else None,
                "ref_mean": float(np.abs(a[fin]).mean()) if d.size else None, "nan_mismatch": int((np.isnan(a) != np.isnan(b)).sum()),
                "nan_ref": int(np.isnan(a).sum()), "nan_ours": int(np.isnan(b).sum()), "n": int(a.size)}
    for k in ("rgb", "disp", "depth1", "normal", "rgbs4", "rgb2", "disp2", "rgb0", "path_rgbs", "path_disps"):
        res[k] = cmp(k)
    res["loss_ref"], res["loss_ours"] = ref_out["loss"], our_out["loss"]
    res["loss_after_ref"], res["loss_after_ours"] = ref_out["loss_after"], our_out["loss_after"]
    res["files_equal"] = ref_out["files"] == our_out["files"]
    res["n_files"] = len(our_out["files"])
    cos = {}
    for k in ref_grads:
        a, b = ref_grads[k].ravel().astype(np.float64), our_grads[k].ravel().astype(np.float64)
        cos[k] = {"cos": float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300)),
                  "rel_l2": float(np.linalg.norm(a - b) / (np.linalg.norm(a) + 1e-300))}
    res["grad"] = cos
    res["grad_min_cos"] = min(v["cos"] for v in cos.values())
    res["grad_max_rel_l2"] = max(v["rel_l2"] for v in cos.values())
    print("INSITU " + json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
