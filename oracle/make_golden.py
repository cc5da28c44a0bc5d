"""Generates tests/golden/*.npz by EXECUTING the unmodified reference on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

Each fixture stores the seeded inputs and the reference's outputs for one stage of the
hot path; tests replay the inputs through oracle/nerf_oracle.py (CPU suite) and through
the CUDA path via the C-ABI (GPU suite).  The reference has no golden vectors of its own
for this path — these are outputs of the reference itself.
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402  (only for init_params: seeded weights)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def nerf_args(basedir, expname):
    return argparse.Namespace(
        multires=10, multires_views=4, i_embed=0, use_viewdirs=True, N_importance=64, N_samples=64,
        netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256, netchunk=65536,
        alpha_model_path=None, no_coarse=False, lrate=5e-4, basedir=basedir, expname=expname,
        ft_path=None, no_reload=True, perturb=1.0, white_bkgd=True, raw_noise_std=1.0,
        dataset_type="llff", no_ndc=True, lindisp=True, sigma_loss=False)


def load_seeded(module, seed):
    """Overwrite a reference NeRF (possibly DataParallel-wrapped) with orc.init_params(seed)."""
    p = orc.init_params(seed)
    sd = module.state_dict()
    new = {k: torch.from_numpy(p[k.replace("module.", "")]) for k in sd}
    module.load_state_dict(new)


def sub_grad(g):
    """Fixture-size control: 256-wide matrices keep every 8th row, the rest are stored whole."""
    return g[::8].copy() if (g.ndim == 2 and g.shape[0] == 256 and g.shape[1] >= 256) else g.copy()


def state_to_np(module):
    sd = module.state_dict()
    return {k.replace("module.", ""): v.detach().cpu().numpy().copy() for k, v in sd.items()}


def synth_rays(run_helpers, n, seed, H=756, W=1008, focal=767.2935):
    """cfg-2 style rays: a seeded permutation of the pinhole grid, identity pose."""
    c2w = torch.eye(4)[:3, :4]
    ro, rd = run_helpers.get_rays(H, W, focal, c2w)
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(H * W, generator=g)[:n]
    return ro.reshape(-1, 3)[idx].contiguous(), rd.reshape(-1, 3)[idx].contiguous()


def main():
    os.makedirs(OUT, exist_ok=True)
    run, helpers = ref_import.load()
    torch.manual_seed(0)
    np.random.seed(0)
    near, far = 1.2, 7.7369

    # ---------------------------------------------------------------- stage: coarse sampling
    fx = {}
    for lindisp in (False, True):
        ro, rd = synth_rays(helpers, 96, seed=1)
        nears = near * (1.0 + 0.3 * torch.rand(96, 1))
        fars = far * (1.0 + 0.3 * torch.rand(96, 1))
        vd = rd / torch.norm(rd, dim=-1, keepdim=True)
        ray_batch = torch.cat([ro, rd, nears, fars, vd], -1)
        for perturb in (0.0, 1.0):
            ret = run.render_rays(ray_batch, network_fn=lambda *a: None, network_query_fn=lambda pts, v, fn: torch.zeros(pts.shape[:-1] + (4,)),
                                  N_samples=64, lindisp=lindisp, perturb=perturb, N_importance=0, pytest=True)
            key = "lindisp%d_perturb%d" % (int(lindisp), int(perturb))
            fx[key + "_rays"] = ray_batch.numpy()
            fx[key + "_z"] = ret["z_vals"].numpy()
    np.random.seed(0)
    fx["t_rand"] = torch.Tensor(np.random.rand(96, 64)).numpy()       # what pytest=True draws (run.py:1777-1779)
    fx["t_vals"] = torch.linspace(0., 1., steps=64).numpy()
    np.savez_compressed(os.path.join(OUT, "sample_coarse.npz"), **fx)

    # ---------------------------------------------------------------- stage: sample_pdf
    fx = {}
    N = 160
    g = torch.Generator().manual_seed(2)
    z = torch.sort(near + (far - near) * torch.rand(N, 64, generator=g), -1)[0]
    bins = .5 * (z[:, 1:] + z[:, :-1])
    dists = {
        "uniform": 0.01 + 0.001 * torch.rand(N, 62, generator=g),
        "peaky": torch.rand(N, 62, generator=g) ** 12,
        "sparse": torch.rand(N, 62, generator=g) * (torch.rand(N, 62, generator=g) > 0.95),
        "zeros": torch.zeros(N, 62),
    }
    fx["bins"] = bins.numpy()
    fx["u_det"] = torch.linspace(0., 1., steps=64).numpy()
    np.random.seed(0)
    fx["u_rand"] = torch.Tensor(np.random.rand(N, 64)).numpy()        # pytest=True stream (helpers:320-327)
    for name, w in dists.items():
        fx["w_" + name] = w.numpy()
        for det in (True, False):
            tag = "%s_%s" % (name, "det" if det else "rand")
            # replicate the internals to capture cdf and inds as well as the samples
            s = helpers.sample_pdf(bins, w, 64, det=det, pytest=(not det))   # det: torch.linspace u (helpers:313)
            ww = w + 1e-5
            pdf = ww / torch.sum(ww, -1, keepdim=True)
            cdf = torch.cat([torch.zeros(N, 1), torch.cumsum(pdf, -1)], -1)
            u = torch.from_numpy(fx["u_det"]).expand(N, 64).contiguous() if det else torch.from_numpy(fx["u_rand"])
            inds = torch.searchsorted(cdf, u, right=True)
            fx["cdf_" + tag] = cdf.numpy()
            fx["inds_" + tag] = inds.numpy()
            fx["samples_" + tag] = s.numpy()
    # ties: handcrafted cdf steps (weights that make exact repeats in cdf are rare; cover via zeros + det)
    np.savez_compressed(os.path.join(OUT, "sample_pdf.npz"), **fx)

    # ---------------------------------------------------------------- stage: merge (sort of cat)
    fx = {}
    zs = helpers.sample_pdf(bins, dists["peaky"], 64, det=False, pytest=True)
    merged, _ = torch.sort(torch.cat([z, zs], -1), -1)
    fx["z"] = z.numpy(); fx["z_samples"] = zs.numpy(); fx["merged"] = merged.numpy()
    fx["z_std"] = torch.std(zs, dim=-1, unbiased=False).numpy()
    np.savez_compressed(os.path.join(OUT, "merge.npz"), **fx)

    # ---------------------------------------------------------------- stage: raw2outputs fwd + autograd bwd
    fx = {}
    for S in (64, 128):
        N = 48
        g = torch.Generator().manual_seed(3 + S)
        raw = (torch.randn(N, S, 4, generator=g) * 1.5)
        raw[0, :, 3] = -1.0                                            # a ray with every sigma <= 0 (disp = NaN)
        raw[1, :, 3] = 30.0                                            # opaque at the first sample
        zz = torch.sort(near + (far - near) * torch.rand(N, S, generator=g), -1)[0]
        rd = torch.randn(N, 3, generator=g)
        noise = torch.randn(N, S, generator=g)
        for white in (False, True):
            for use_noise in (False, True):
                tag = "S%d_w%d_n%d" % (S, int(white), int(use_noise))
                r = raw.clone().requires_grad_(True)
                rr = r if not use_noise else torch.cat([r[..., :3], (r[..., 3] + noise)[..., None]], -1)
                rgb, disp, acc, wts, depth, alpha = helpers.raw2outputs(rr, zz, rd, 0, white, need_alpha=True)
                g_rgb = torch.randn(N, 3, generator=g); g_disp = torch.randn(N, generator=g)
                g_acc = torch.randn(N, generator=g); g_depth = torch.randn(N, generator=g)
                g_w = torch.randn(N, S, generator=g)
                # NaN disp rays stay in the graph: the reference then yields NaN grads for exactly those rays
                loss = (rgb * g_rgb).sum() + (disp * g_disp).sum() + (acc * g_acc).sum() + \
                    (depth * g_depth).sum() + (wts * g_w).sum()
                loss.backward()
                fx[tag + "_rgb"] = rgb.detach().numpy(); fx[tag + "_disp"] = disp.detach().numpy()
                fx[tag + "_acc"] = acc.detach().numpy(); fx[tag + "_weights"] = wts.detach().numpy()
                fx[tag + "_depth"] = depth.detach().numpy(); fx[tag + "_alpha"] = alpha.detach().numpy()
                fx[tag + "_g_rgb"] = g_rgb.numpy(); fx[tag + "_g_disp"] = g_disp.numpy()
                fx[tag + "_g_acc"] = g_acc.numpy(); fx[tag + "_g_depth"] = g_depth.numpy()
                fx[tag + "_g_weights"] = g_w.numpy(); fx[tag + "_d_raw"] = r.grad.numpy()
        fx["S%d_raw" % S] = raw.numpy(); fx["S%d_z" % S] = zz.numpy()
        fx["S%d_rays_d" % S] = rd.numpy(); fx["S%d_noise" % S] = noise.numpy()
    np.savez_compressed(os.path.join(OUT, "raw2outputs.npz"), **fx)

    # ---------------------------------------------------------------- stage: embedder + NeRF MLP fwd/bwd
    fx = {}
    torch.manual_seed(4)
    embed_fn, in_ch = helpers.get_embedder(10, 0)
    embedd_fn, in_chv = helpers.get_embedder(4, 0)
    model = helpers.NeRF(D=8, W=256, input_ch=in_ch, output_ch=5, skips=[4], input_ch_views=in_chv, use_viewdirs=True)
    load_seeded(model, 104)
    fx["param_seed"] = np.int64(104)
    P = 256
    pts = (torch.rand(P, 3) * 2 - 1) * 4.0
    vd = torch.randn(P, 3); vd = vd / vd.norm(dim=-1, keepdim=True)
    x = torch.cat([embed_fn(pts), embedd_fn(vd)], -1)
    out = model(x)
    d_out = torch.randn(P, 4)
    (out * d_out).sum().backward()
    fx["pts"] = pts.numpy(); fx["viewdirs"] = vd.numpy(); fx["embedded"] = x.detach().numpy()
    fx["out"] = out.detach().numpy(); fx["d_out"] = d_out.numpy()
    for k, v in model.named_parameters():
        fx["grad." + k] = v.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "nerf_mlp.npz"), **fx)

    # ---------------------------------------------------------------- stage: normal map fwd + autograd bwd
    fx = {}
    g = torch.Generator().manual_seed(5)
    Hn, Wn = 40, 56
    yy, xx = torch.meshgrid(torch.arange(Hn, dtype=torch.float32), torch.arange(Wn, dtype=torch.float32), indexing="ij")
    depth = (3.0 + 0.02 * xx - 0.015 * yy + 0.3 * torch.sin(xx / 7.0) * torch.cos(yy / 5.0) + 0.05 * torch.rand(Hn, Wn, generator=g))
    fr = 60.0
    Kmat = torch.Tensor([[fr, 0, Wn / 2], [0, fr, Hn / 2], [0, 0, 1]])
    for dt in (torch.float64, torch.float32):
        torch.set_default_dtype(dt)
        d = depth.to(dt).clone().requires_grad_(True)
        xyz = run.depth2xyz_torch(d, Kmat.to(dt)).to(dt)
        nrm = run.depth2normal_geo(xyz.permute(2, 0, 1).unsqueeze(0), k=31)[0]
        g_n = torch.randn(3, Hn, Wn, generator=g).to(dt)
        (nrm * g_n).sum().backward()
        tag = "f64" if dt == torch.float64 else "f32"
        fx["normal_" + tag] = nrm.detach().numpy(); fx["g_normal_" + tag] = g_n.numpy()
        fx["d_depth_" + tag] = d.grad.numpy()
    torch.set_default_dtype(torch.float32)
    fx["depth"] = depth.numpy(); fx["K"] = Kmat.numpy()
    np.savez_compressed(os.path.join(OUT, "normal_map.npz"), **fx)

    # ---------------------------------------------------------------- end to end: render() on 64 rays, fwd (+bwd)
    fx = {}
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "exp"))
        torch.manual_seed(0)
        kw_train, kw_test, _, grad_vars, _ = run.create_nerf(nerf_args(td, "exp"))
    nets = {"coarse": kw_train["network_fn"], "fine": kw_train["network_fine"]}
    load_seeded(nets["coarse"], 200); load_seeded(nets["fine"], 201)
    fx["coarse_seed"] = np.int64(200); fx["fine_seed"] = np.int64(201)
    ro, rd = synth_rays(helpers, 64, seed=1)
    fx["rays_o"] = ro.numpy(); fx["rays_d"] = rd.numpy()
    fx["near"] = np.float32(near); fx["far"] = np.float32(far)
    # (i) render kwargs: perturb=0, noise=0 (run.py:1590-1591)
    with torch.no_grad():
        rgb, disp, acc, depth, extras = run.render(756, 1008, 767.2935, chunk=32768, rays=torch.stack([ro, rd], 0),
                                                   near=near, far=far, retraw=True, need_alpha=True, **kw_test)
    fx["test_rgb"] = rgb.numpy(); fx["test_disp"] = disp.numpy(); fx["test_acc"] = acc.numpy()
    fx["test_depth"] = depth.numpy()
    for k in ("weights", "z_vals", "raw", "rgb0", "disp0", "acc0", "z_std", "alpha", "alpha0"):
        fx["test_" + k] = extras[k].numpy()
    # (ii) train kwargs with pytest=True randoms + backward of a simple loss
    for v in grad_vars:
        v.grad = None
    rgb, disp, acc, depth, extras = run.render(756, 1008, 767.2935, chunk=32768, rays=torch.stack([ro, rd], 0),
                                               near=near, far=far, retraw=True, pytest=True, **kw_train)
    loss = helpers.img2mse(rgb, torch.full_like(rgb, 0.5)) + helpers.img2mse(extras["rgb0"], torch.full_like(rgb, 0.5)) \
        + 0.1 * helpers.img2mse(disp, torch.full_like(disp, 0.3))
    loss.backward()
    fx["train_rgb"] = rgb.detach().numpy(); fx["train_disp"] = disp.detach().numpy()
    fx["train_acc"] = acc.detach().numpy(); fx["train_depth"] = depth.detach().numpy()
    fx["train_loss"] = np.float32(loss.item())
    for k in ("weights", "z_vals", "raw", "rgb0", "disp0", "acc0", "z_std"):
        fx["train_" + k] = extras[k].detach().numpy()
    for nm, net in nets.items():
        for k, v in net.named_parameters():
            fx["train_grad.%s.%s" % (nm, k.replace("module.", ""))] = sub_grad(v.grad.numpy())
    # the pytest=True streams (every draw re-seeds MT19937 with 0; noise is UNIFORM, helpers:377-381)
    np.random.seed(0); fx["train_t_rand"] = torch.Tensor(np.random.rand(64, 64)).numpy()
    np.random.seed(0); fx["train_noise0"] = torch.Tensor(np.random.rand(64, 64) * 1.0).numpy()
    np.random.seed(0); fx["train_u"] = torch.Tensor(np.random.rand(64, 64)).numpy()
    np.random.seed(0); fx["train_noise1"] = torch.Tensor(np.random.rand(64, 128) * 1.0).numpy()
    np.savez_compressed(os.path.join(OUT, "render_e2e.npz"), **fx)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
