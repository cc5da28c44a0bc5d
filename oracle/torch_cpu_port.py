"""CPU baseline arm — TEST / BENCH INFRASTRUCTURE ONLY (never imported by mvip_nerf_b200/).

A functional PyTorch-CPU port of the reference's render path (same op sequence as
DS_NeRF/run.py:1703-1847 + run_nerf_helpers.py, fp32, autograd, all host threads), used by bench.py for
`cpu_baseline` and `--impl reference` (kind "port"): the reference itself is a Python tree under
/root/reference that does not exist on the GPU box, so the same torch CPU kernels it would call
(mm/addmm, sin/cos, cat, relu, sigmoid, exp, cumsum/cumprod, searchsorted, sort) are timed through this
port instead.  tests/test_oracle_golden.py pins it against the reference-generated golden vectors.
"""
import torch


def pe(x, n_freqs):
    """[x, sin(2^k x), cos(2^k x)]_k   (run_nerf_helpers.py:39-52)"""
    bands = 2. ** torch.linspace(0., n_freqs - 1, steps=n_freqs)
    parts = [x]
    for f in bands:
        parts += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(parts, -1)


def mlp(p, x):
    """NeRF.forward with use_viewdirs (run_nerf_helpers.py:104-127); p maps names to tensors."""
    lin = lambda name, h: torch.nn.functional.linear(h, p[name + ".weight"], p[name + ".bias"])  # noqa: E731
    pts, views = x[:, :63], x[:, 63:]
    h = pts
    for i in range(8):
        h = torch.relu(lin("pts_linears.%d" % i, h))
        if i == 4:
            h = torch.cat([pts, h], -1)
    alpha = lin("alpha_linear", h)
    h = torch.cat([lin("feature_linear", h), views], -1)
    h = torch.relu(lin("views_linears.0", h))
    return torch.cat([lin("rgb_linear", h), alpha], -1)


def query(p, pts, viewdirs, netchunk=65536):
    """run_network (run.py:1108-1124)"""
    flat = pts.reshape(-1, 3)
    dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)
    x = torch.cat([pe(flat, 10), pe(dirs, 4)], -1)
    out = torch.cat([mlp(p, x[i:i + netchunk]) for i in range(0, x.shape[0], netchunk)], 0)
    return out.reshape(pts.shape[0], pts.shape[1], 4)


def composite(raw, z, rays_d, noise, white_bkgd):
    """raw2outputs (run_nerf_helpers.py:350-404)"""
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], 1e10)], -1) * torch.norm(rays_d[:, None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    sigma = raw[..., 3] if noise is None else raw[..., 3] + noise
    alpha = 1. - torch.exp(-torch.relu(sigma) * dists)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :1]), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    w = alpha * trans
    rgb_map = (w[..., None] * rgb).sum(-2)
    depth = (w * z).sum(-1)
    acc = w.sum(-1)
    disp = 1. / torch.max(1e-10 * torch.ones_like(depth), depth / acc)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc[:, None])
    return rgb_map, disp, acc, w, depth


def importance(z, w, u):
    """z_mid + sample_pdf + sort-merge (run.py:1809-1814, run_nerf_helpers.py:304-347)"""
    bins = .5 * (z[:, 1:] + z[:, :-1])
    ww = w[:, 1:-1] + 1e-5
    pdf = ww / ww.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)
    u = u.expand(z.shape[0], u.shape[-1]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below, above = (inds - 1).clamp(min=0), inds.clamp(max=cdf.shape[-1] - 1)
    cb, ca = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bb, ba = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    den = ca - cb
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    samples = (bb + (u - cb) / den * (ba - bb)).detach()
    return torch.sort(torch.cat([z, samples], -1), -1)[0], samples


def render_rays(rays, pc, pf, t_vals, t_rand=None, u=None, noise0=None, noise1=None, lindisp=True, white_bkgd=True,
                n_importance=64):
    o, d, near, far, vd = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8], rays[:, -3:]
    if lindisp:
        z = 1. / (1. / near * (1. - t_vals) + 1. / far * t_vals)
    else:
        z = near * (1. - t_vals) + far * t_vals
    z = z.expand(rays.shape[0], t_vals.shape[0])
    if t_rand is not None:
        mids = .5 * (z[:, 1:] + z[:, :-1])
        upper, lower = torch.cat([mids, z[:, -1:]], -1), torch.cat([z[:, :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    raw0 = query(pc, o[:, None] + d[:, None] * z[..., None], vd)
    rgb0, disp0, acc0, w0, depth0 = composite(raw0, z, d, noise0, white_bkgd)
    if u is None:
        u = torch.linspace(0., 1., steps=n_importance)
    z1, samples = importance(z, w0, u)
    raw1 = query(pf, o[:, None] + d[:, None] * z1[..., None], vd)
    rgb, disp, acc, w, depth = composite(raw1, z1, d, noise1, white_bkgd)
    return {"rgb_map": rgb, "disp_map": disp, "acc_map": acc, "depth_map": depth, "weights": w, "z_vals": z1,
            "rgb0": rgb0, "disp0": disp0, "acc0": acc0, "z_samples": samples}


def params_from_numpy(p_np, requires_grad=False):
    return {k: torch.from_numpy(v.copy()).requires_grad_(requires_grad) for k, v in p_np.items()}


def train_step(rays, pc, pf, t_vals, t_rand, u, noise0, noise1, target):
    """forward + mse(rgb) + mse(rgb0) + backward; returns the loss (grads land in pc / pf tensors)."""
    for p in list(pc.values()) + list(pf.values()):
        p.grad = None
    out = render_rays(rays, pc, pf, t_vals, t_rand, u, noise0, noise1)
    loss = ((out["rgb_map"] - target) ** 2).mean() + ((out["rgb0"] - target) ** 2).mean()
    loss.backward()
    return loss
