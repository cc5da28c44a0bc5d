"""Drop-in for the render path of the reference's DS_NeRF/run.py:
batchify / run_network (:1096-1124), batchify_rays (:1127), render (:1143), create_nerf (:1474),
render_rays (:1703), depth2xyz_torch (:1909), depth2normal_geo (:1924).

Same names, arguments and return structures; the per-ray math runs in the sm_100a kernels of
libmvip_nerf.so.  The training driver, data loaders, SDS guidance, GUI and the tcnn model are out of scope
 and stay with the reference, which can import these functions in place of its own.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .optim import FusedAdam
from .run_nerf_helpers import (NeRF, _NormalFromXYZ, get_embedder, get_rays, ndc_rays, raw2outputs, sample_pdf)

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
DEBUG = False


class DataParallel(nn.Module):
    """Stand-in for the nn.DataParallel wrapper of run.py:1491/1527: keeps the `module.` prefix of the
    reference's checkpoints (run.py:1043-1053) and the `.module` attribute, without the per-call
    scatter/broadcast/gather.  Scale-out is one process per GPU (mvip_nerf_b200.dist)."""

    def __init__(self, module, device_ids=None):
        super().__init__()
        self.module = module

    def forward(self, *a, **kw):
        return self.module(*a, **kw)


def _unwrap(net):
    return net.module if isinstance(net, (DataParallel, nn.DataParallel)) else net


def batchify(fn, chunk):
    """(run.py:1096-1105)"""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def run_network(inputs2, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """(run.py:1108-1124) inputs2 [N,S,3], viewdirs [N,3] -> [N,S,4].

    When `fn` is our NeRF and the embedders are the standard ones, the embedding is never materialised: the
    fused kernel encodes on chip (netchunk is then irrelevant — the kernel is persistent and tiles internally).
    Otherwise the reference's sequence is followed with our embed kernel and `fn` applied in netchunk pieces."""
    net = _unwrap(fn)
    flat = torch.reshape(inputs2, [-1, inputs2.shape[-1]])
    fused = (isinstance(net, NeRF) and viewdirs is not None and getattr(embed_fn, "multires", None) == 10 and
             getattr(embeddirs_fn, "multires", None) == 4 and flat.shape[-1] == 3)
    if fused:
        dirs = viewdirs[:, None].expand(inputs2.shape).reshape(-1, 3)
        out = net.query_points(flat.contiguous(), dirs.contiguous())
        return torch.reshape(out, list(inputs2.shape[:-1]) + [4])
    embedded = embed_fn(flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs2.shape)
        embedded = torch.cat([embedded, embeddirs_fn(torch.reshape(input_dirs, [-1, input_dirs.shape[-1]]))], -1)
    outputs_flat = batchify(fn, netchunk)(embedded)
    return torch.reshape(outputs_flat, list(inputs2.shape[:-1]) + [outputs_flat.shape[-1]])


class _FusedQuery:
    """network_query_fn built by create_nerf: callable like the reference's closure (run.py:1530-1533) and
    recognised by render_rays, which then feeds rays + depths straight to the fused kernel."""

    def __init__(self, embed_fn, embeddirs_fn, netchunk):
        self.embed_fn, self.embeddirs_fn, self.netchunk = embed_fn, embeddirs_fn, netchunk

    def __call__(self, inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, embed_fn=self.embed_fn, embeddirs_fn=self.embeddirs_fn,
                           netchunk=self.netchunk)


def batchify_rays(rays_flat, chunk=1024 * 32, need_alpha=False, detach_weights=False, **kwargs):
    """(run.py:1127-1140) render in chunks, concatenate per key."""
    pieces = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], need_alpha=need_alpha, detach_weights=detach_weights, **kwargs)
        for k, v in ret.items():
            pieces.setdefault(k, []).append(v)
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in pieces.items()}


def _is_number(x):
    return isinstance(x, (int, float)) or (isinstance(x, np.generic) and np.ndim(x) == 0)


def render(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, depths=None, need_alpha=False, detach_weights=False, patch=None, **kwargs):
    """(run.py:1143-1219) -> [rgb_map, disp_map, acc_map, depth_map, extras].

    The ray batch [N, 8|11] is written by ONE kernel (get_rays / patch / viewdir normalisation / ndc_rays / near / far /
    cat): mvip_rays_from_pose for `c2w=`, mvip_rays_pack for `rays=`.  Only the `depths=` column (used by the out-of-scope
    sigma loss) and non-scalar near / far / focal take the elementwise torch route below."""
    fusable = depths is None and torch.cuda.is_available() and _is_number(near) and _is_number(far) and _is_number(focal)
    rays_flat = None
    if fusable and c2w is not None:
        if patch is not None and (patch[0] + patch[2] > H or patch[1] + patch[3] > W):
            raise RuntimeError("patch outside the image")
        rays_flat = ops.rays_from_pose(H, W, focal, c2w, near, far, use_viewdirs=use_viewdirs, c2w_staticcam=c2w_staticcam,
                                       patch=patch, ndc=ndc)
        sh = ((H, W) if patch is None else (int(patch[2]), int(patch[3]))) + (3,)
    elif fusable and c2w is None and (c2w_staticcam is None or not use_viewdirs):
        rays_o, rays_d = rays
        sh = rays_d.shape
        rays_flat = ops.rays_pack(rays_o, rays_d, near, far, use_viewdirs=use_viewdirs, ndc=ndc, H=H, W=W, focal=focal)
    if rays_flat is not None:
        all_ret = batchify_rays(rays_flat, chunk, need_alpha=need_alpha, detach_weights=detach_weights, **kwargs)
        for k in all_ret:
            all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
        k_extract = ['rgb_map', 'disp_map', 'acc_map', 'depth_map']
        return [all_ret[k] for k in k_extract] + [{k: all_ret[k] for k in all_ret if k not in k_extract}]
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, focal, c2w)
        if patch is not None:
            i, j, len1, len2 = patch
            rays_o = rays_o[i:i + len1, j:j + len2, :]
            rays_d = rays_d[i:i + len1, j:j + len2, :]
    else:
        rays_o, rays_d = rays
    viewdirs = None
    if use_viewdirs:
        viewdirs = rays_d
        if c2w_staticcam is not None:
            rays_o, rays_d = get_rays(H, W, focal, c2w_staticcam)
        viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
        viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    sh = rays_d.shape
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, focal, 1., rays_o, rays_d)
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    ones = torch.ones_like(rays_d[..., :1])
    cols = [rays_o, rays_d, near * ones, far * ones]
    if depths is not None:
        cols.append(depths.reshape(-1, 1).to(rays_d))
    if use_viewdirs:
        cols.append(viewdirs)
    all_ret = batchify_rays(torch.cat(cols, -1), chunk, need_alpha=need_alpha, detach_weights=detach_weights, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    k_extract = ['rgb_map', 'disp_map', 'acc_map', 'depth_map']
    return [all_ret[k] for k in k_extract] + [{k: all_ret[k] for k in all_ret if k not in k_extract}]


def create_nerf(args):
    """(run.py:1474-1599) -> (render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer)."""
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed)
    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    if getattr(args, "alpha_model_path", None) is not None:
        raise NotImplementedError("alpha_model_path / NeRF_RGB is out of scope (unused by config_1, run.py:173)")
    output_ch = 5 if args.N_importance > 0 else 4
    skips = [4]
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    model = DataParallel(NeRF(D=args.netdepth, W=args.netwidth, input_ch=input_ch, output_ch=output_ch, skips=skips,
                              input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs).to(dev))
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = NeRF(D=args.netdepth_fine, W=args.netwidth_fine, input_ch=input_ch, output_ch=output_ch,
                          skips=skips, input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs).to(dev)
        grad_vars += list(model_fine.parameters())
        model_fine = DataParallel(model_fine)

    network_query_fn = _FusedQuery(embed_fn, embeddirs_fn, args.netchunk)
    optimizer = FusedAdam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))   # torch.optim.Adam semantics, one launch per step

    start = 0
    basedir, expname = args.basedir, args.expname
    if args.ft_path is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    else:
        d = os.path.join(basedir, expname)
        ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if 'tar' in f]
    print('Found ckpts', ckpts)
    if len(ckpts) > 0 and not args.no_reload:
        ckpt_path = ckpts[-1]
        print('Reloading from', ckpt_path)
        ckpt = torch.load(ckpt_path, map_location=dev)
        start = ckpt['global_step']
        optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        model.load_state_dict(ckpt['network_fn_state_dict'])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt['network_fine_state_dict'])

    render_kwargs_train = {
        'network_query_fn': network_query_fn, 'perturb': args.perturb, 'N_importance': args.N_importance,
        'network_fine': model_fine, 'N_samples': args.N_samples, 'network_fn': model,
        'use_viewdirs': args.use_viewdirs, 'white_bkgd': args.white_bkgd, 'raw_noise_std': args.raw_noise_std,
    }
    if args.dataset_type != 'llff' or args.no_ndc:
        print('Not ndc!')
        render_kwargs_train['ndc'] = False
        render_kwargs_train['lindisp'] = args.lindisp
    else:
        render_kwargs_train['ndc'] = True
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    if getattr(args, "sigma_loss", False):
        raise NotImplementedError("SigmaLoss (loss.py) is out of scope: unset in config_1 (SURVEY.md §2 row 13)")
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer


def _pytest_uniform(shape, dev):
    np.random.seed(0)
    return torch.Tensor(np.random.rand(*list(shape))).to(dev)


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., pytest=False, sigma_loss=None,
                verbose=False, need_alpha=False, detach_weights=False, _randoms=None):
    """(run.py:1703-1847) volumetric rendering of one chunk of rays -> dict with the reference's keys.

    Random numbers are drawn on the host side in the reference's order (t_rand, coarse noise, u, fine noise)
    and handed to the kernels; `_randoms` (dict with any of t_rand / noise0 / u / noise1) overrides them so
    tests and multi-GPU runs can share one stream."""
    if network_fn is None:
        raise NotImplementedError("render_rays without a coarse network (alpha_model path) is out of scope")
    if sigma_loss is not None:
        raise NotImplementedError("sigma_loss is out of scope (SURVEY.md §2 row 13)")
    rnd = _randoms or {}
    ray_batch = ray_batch.float().contiguous()
    dev = ray_batch.device
    N_rays = ray_batch.shape[0]
    rays_d = ray_batch[:, 3:6]
    has_dirs = ray_batch.shape[-1] > 9
    coarse = _unwrap(network_fn)
    fused = isinstance(network_query_fn, _FusedQuery) and isinstance(coarse, NeRF) and has_dirs

    # ---- stratified samples (run.py:1759-1781) --------------------------------------------------------
    t_vals = torch.linspace(0., 1., steps=N_samples, device=dev)
    t_rand = None
    if perturb > 0.:
        t_rand = rnd.get("t_rand")
        if t_rand is None:
            t_rand = _pytest_uniform((N_rays, N_samples), dev) if pytest else torch.rand((N_rays, N_samples), device=dev)
    z_vals = ops.sample_coarse(ray_batch, t_vals, t_rand, lindisp)

    def query(net, z):
        if fused and isinstance(_unwrap(net), NeRF):
            return _unwrap(net).query_rays(ray_batch, z)
        pts = ray_batch[:, None, 0:3] + ray_batch[:, None, 3:6] * z[..., :, None]
        return network_query_fn(pts, ray_batch[:, -3:] if has_dirs else None, net)

    def noise_for(shape, key):
        if key in rnd:
            return rnd[key]
        if raw_noise_std > 0.:
            if pytest:
                return _pytest_uniform(shape, dev) * raw_noise_std
            n = torch.randn(shape, device=dev)
            return n if raw_noise_std == 1. else n * raw_noise_std     # x * 1.0 == x: skip the launch
        return None

    raw = query(network_fn, z_vals)
    rgb_map, disp_map, acc_map, weights, depth_map, alpha = raw2outputs(
        raw, z_vals, rays_d, raw_noise_std, white_bkgd, need_alpha=need_alpha, detach_weights=detach_weights,
        _noise=noise_for((N_rays, N_samples), "noise0"))

    z_std = None
    if N_importance > 0:
        rgb_map_0, disp_map_0, acc_map_0, alpha0 = rgb_map, disp_map, acc_map, alpha
        # ---- hierarchical samples: z_mid, sample_pdf, detach, sort-merge (run.py:1809-1816) ------------
        u = rnd.get("u")
        if u is None:
            if perturb == 0.:
                u = torch.linspace(0., 1., steps=N_importance, device=dev)
                if pytest:
                    u = torch.Tensor(np.linspace(0., 1., N_importance)).to(dev)
            else:
                u = _pytest_uniform((N_rays, N_importance), dev) if pytest else torch.rand((N_rays, N_importance), device=dev)
        fs = ops.sample_fine(z_vals, weights.detach(), u, want_samples=False)
        z_vals, z_std = fs["z_merged"], fs["z_std"]
        run_fn = network_fn if network_fine is None else network_fine
        raw = query(run_fn, z_vals)
        rgb_map, disp_map, acc_map, weights, depth_map, alpha = raw2outputs(
            raw, z_vals, rays_d, raw_noise_std, white_bkgd, need_alpha=need_alpha, detach_weights=detach_weights,
            _noise=noise_for((N_rays, N_samples + N_importance), "noise1"))

    ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map, 'depth_map': depth_map,
           'weights': weights, 'z_vals': z_vals}
    if retraw:
        ret['raw'] = raw
    if need_alpha:
        ret['alpha'] = alpha
        ret['alpha0'] = alpha0      # NameError when N_importance == 0, exactly as in the reference (run.py:1831)
    if N_importance > 0:
        ret['rgb0'] = rgb_map_0
        ret['disp0'] = disp_map_0
        ret['acc0'] = acc_map_0
        ret['z_std'] = z_std
    if DEBUG:
        for k in ret:
            if torch.isnan(ret[k]).any() or torch.isinf(ret[k]).any():
                print(f"! [Numerical Error] {k} contains nan or inf.")
    return ret


# ---------------------------------------------------------------------------------------------------
# normal map from depth (run.py:1909-1940; call site :948-965)
# ---------------------------------------------------------------------------------------------------
def depth2xyz_torch(depth_map, depth_cam_matrix, depth_scale=1.0):
    """tensor(h,w), tensor(3,3) -> tensor(h,w,3)  (elementwise glue; the fused route is run_nerf_helpers.depth2normal)"""
    fx, fy = depth_cam_matrix[0, 0], depth_cam_matrix[1, 1]
    cx, cy = depth_cam_matrix[0, 2], depth_cam_matrix[1, 2]
    hh, ww = torch.meshgrid(torch.arange(depth_map.shape[0], device=depth_map.device, dtype=torch.float32),
                            torch.arange(depth_map.shape[1], device=depth_map.device, dtype=torch.float32), indexing="ij")
    z = depth_map / depth_scale
    return torch.stack([(ww - cx) * z / fx, (hh - cy) * z / fy, z], -1)


def depth2normal_geo(depth, k=31):
    """tensor(b,3,h,w) point map -> tensor(b,3,h,w) least-squares normals (not normalised)."""
    return torch.stack([_NormalFromXYZ.apply(depth[b].contiguous(), int(k)) for b in range(depth.shape[0])], 0)
