"""GPU suite: the callers either side of the hot path (SURVEY.md §8 f2 / f4 and BASELINE cfg 5): deferred back-propagation,
render_path / render_path_4view / checkpoints, the multi-view guidance batch with normal maps."""
import os
import tempfile

import numpy as np
import pytest
import torch

from test_gpu_render import cu, load_seeded, nerf_args

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nerf():
    from mvip_nerf_b200 import run
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "exp"))
        kw_train, kw_test, start, grad_vars, opt = run.create_nerf(nerf_args(td))
    load_seeded(kw_train["network_fn"], 200)
    load_seeded(kw_train["network_fine"], 201)
    return run, kw_train, kw_test, grad_vars, opt


def pose(tx=0.0, ry=0.0):
    c, s = np.cos(ry), np.sin(ry)
    m = np.array([[c, 0, s, tx], [0, 1, 0, 0.05], [-s, 0, c, 0.1]], dtype=np.float32)
    return torch.from_numpy(m).cuda()


def image_loss(rgb, disp, depth, extras):
    w = torch.linspace(0.5, 1.5, rgb.shape[0] * rgb.shape[1], device=rgb.device).view(rgb.shape[:2])
    return ((rgb - 0.4) ** 2 * w[..., None]).mean() + ((extras["rgb0"] - 0.6) ** 2).mean() + 0.05 * (disp * w).mean() + \
        0.01 * (depth * w).mean()


@pytest.mark.parametrize("mode", ["test", "train"])
def test_deferred_backprop_matches_direct(nerf, mode):
    """render_deferred == render: identical outputs (bitwise: same kernels, same chunking, replayed randoms) and parameter
    gradients equal up to the fp32 accumulation order across chunks."""
    run, kw_train, kw_test, grad_vars, opt = nerf
    kw = kw_test if mode == "test" else kw_train
    H, W, focal, chunk = 20, 28, 26.0, 192          # 560 rays -> 3 chunks, the last ragged
    c2w = pose(0.1, 0.05)
    torch.manual_seed(7)
    rgb, disp, acc, depth, ex = run.render(H, W, focal, chunk=chunk, c2w=c2w, near=1.2, far=7.7, **kw)
    gd = torch.autograd.grad(image_loss(rgb, disp, depth, ex), grad_vars)
    torch.manual_seed(7)
    rgb2, disp2, acc2, depth2, ex2 = run.render_deferred(H, W, focal, chunk=chunk, c2w=c2w, near=1.2, far=7.7, **kw)
    assert torch.equal(rgb2, rgb.detach()) and torch.equal(depth2, depth.detach()) and torch.equal(ex2["rgb0"], ex["rgb0"].detach())
    assert torch.equal(ex2["z_vals"], ex["z_vals"]) and not ex2["weights"].requires_grad
    mem0 = torch.cuda.memory_allocated()
    gq = torch.autograd.grad(image_loss(rgb2, disp2, depth2, ex2), grad_vars)
    assert torch.cuda.memory_allocated() - mem0 < 64 << 20          # nothing but the gradients survives the backward
    for a, b in zip(gd, gq):
        scale = float(a.abs().max()) + 1e-12
        assert float((a - b).abs().max()) <= 2e-4 * scale


def test_deferred_rays_entry_and_loss_backward(nerf):
    run, kw_train, kw_test, grad_vars, opt = nerf
    g = torch.Generator().manual_seed(0)
    ro = torch.zeros(300, 3).cuda()
    rd = torch.nn.functional.normalize(torch.randn(300, 3, generator=g) * 0.2 + torch.tensor([0., 0., -1.]), dim=-1).cuda()
    for v in grad_vars:
        v.grad = None
    rgb, disp, acc, depth, ex = run.render_deferred(756, 1008, 767.2935, chunk=128, rays=torch.stack([ro, rd], 0), near=1.2, far=7.7,
                                                    **kw_test)
    assert rgb.shape == (300, 3) and rgb.requires_grad
    ((rgb - 0.5) ** 2).mean().backward()
    # only the fine network feeds rgb_map (z_samples is detached, run.py:1812): as with render(), the coarse one gets no gradient
    assert all(p.grad is None for p in kw_test["network_fn"].parameters())
    fine = list(kw_test["network_fine"].parameters())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in fine)
    assert any(float(p.grad.abs().max()) > 0.0 for p in fine)


def test_deferred_accepts_what_render_accepts(nerf):
    """render_deferred shares render()'s ray-batch assembly: per-ray near / far tensors (the reference allows them,
    run.py:1201-1203) and a static camera with view directions take the torch route instead of raising, and give render()'s
    numbers."""
    run, kw_train, kw_test, grad_vars, opt = nerf
    g = torch.Generator().manual_seed(1)
    ro = torch.zeros(200, 3).cuda()
    rd = torch.nn.functional.normalize(torch.randn(200, 3, generator=g) * 0.2 + torch.tensor([0., 0., -1.]), dim=-1).cuda()
    rays = torch.stack([ro, rd], 0)
    near = (1.0 + 0.4 * torch.rand(200, 1, generator=g)).cuda()
    far = (7.0 + torch.rand(200, 1, generator=g)).cuda()
    with torch.no_grad():
        want = run.render(756, 1008, 767.2935, chunk=128, rays=rays, near=near, far=far, **kw_test)
    got = run.render_deferred(756, 1008, 767.2935, chunk=128, rays=rays, near=near, far=far, **kw_test)
    assert torch.equal(got[0].detach(), want[0]) and torch.equal(got[3].detach(), want[3])
    c2w, cam = pose(0.1, 0.05), pose(-0.2, 0.0)
    with torch.no_grad():
        want = run.render(12, 16, 20.0, chunk=128, c2w=c2w, c2w_staticcam=cam, near=1.2, far=7.7, **kw_test)
    got = run.render_deferred(12, 16, 20.0, chunk=128, c2w=c2w, c2w_staticcam=cam, near=1.2, far=7.7, **kw_test)
    assert torch.equal(got[0].detach(), want[0])
    got[0].mean().backward()


def test_render_path_writes_reference_layout(nerf, tmp_path):
    run, kw_train, kw_test, grad_vars, opt = nerf
    poses = torch.stack([pose(0.0), pose(0.2, 0.1)], 0)
    hwf = [40, 56, 52.0]
    kw = dict(kw_test, near=1.2, far=7.7)                    # train() merges the bounds into the kwargs (run.py:554-559)
    rgbs, disps, (Xs, Ys) = run.render_path(poses, hwf, 4096, kw, savedir=str(tmp_path), render_factor=2, need_alpha=True)
    assert isinstance(rgbs, np.ndarray) and rgbs.shape == (2, 20, 28, 3) and disps.shape == (2, 20, 28) and Xs == []
    for sub, ext in (("rgb", "png"), ("depth", "npy"), ("disp", "npy"), ("weight", "npy"), ("z", "npy"), ("alpha", "npy"), ("pose", "txt")):
        assert os.path.isfile(os.path.join(str(tmp_path), sub, "000001." + ext)), sub
    K = np.loadtxt(os.path.join(str(tmp_path), "intrinsics.txt"))
    assert np.allclose(K, [[26, 0, 14], [0, 26, 10], [0, 0, 1]])
    assert np.load(os.path.join(str(tmp_path), "weight", "000000.npy")).shape == (20, 28, 128)
    with torch.no_grad():
        direct = run.render(20, 28, 26.0, chunk=4096, c2w=poses[1][:3, :4], **kw)
    assert np.array_equal(rgbs[1], direct[0].cpu().numpy())
    assert np.array_equal(np.load(os.path.join(str(tmp_path), "disp", "000001.npy")), direct[1].cpu().numpy())
    # gradient-carrying patch renders (the SDS branch of the reference's train loop)
    masks = np.zeros((2, 40, 56), np.uint8)
    masks[:, 10:30, 12:44] = 1
    rg, dg, (Xs, Ys) = run.render_path(poses, hwf, 4096, kw, render_factor=2, rgb_require_grad=True, disp_require_grad=True,
                                       patch_len=(6, 8), masks=masks, deferred_backprop=True)
    assert rg.shape == (2, 6, 8, 3) and rg.requires_grad and dg.shape == (2, 6, 8) and len(Xs) == 2
    rg.sum().backward()


def test_render_path_4view_and_projection(nerf):
    run, kw_train, kw_test, grad_vars, opt = nerf
    poses = torch.stack([pose(0.02 * i, 0.01 * i) for i in range(12)], 0)
    masks = np.zeros((12, 16, 24), np.uint8)
    kw = dict(kw_test, near=1.2, far=7.7)
    rgbs, disps, sel = run.render_path_4view(65, masks, poses, [16, 24, 20.0], 2048, kw, need_alpha=True, deferred_backprop=True)
    # iter = 65 % 60 = 5 -> poses [1:10:2] = 5 views
    assert rgbs.shape == (5, 16, 24, 3) and disps.shape == (5, 16, 24) and len(sel) == 5 and rgbs.requires_grad
    with torch.no_grad():
        want = run.render(16, 24, 20.0, chunk=2048, c2w=poses[3][:3, :4], **kw)
    assert torch.equal(rgbs[1].detach(), want[0])
    z, w, c2ws, K = run.render_path_projection(poses[:2], [16, 24, 20.0], 2048, kw)
    assert z[0].shape == (16, 24, 128) and w[1].shape == (16, 24, 128) and c2ws[0].shape == (4, 4) and c2ws[0][1, 1] == -poses[0][1, 1].item()


def test_checkpoint_roundtrip_with_reference_keys(nerf):
    run, kw_train, kw_test, grad_vars, opt = nerf
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "exp"))
        path = os.path.join(td, "exp", "{:06d}.tar".format(1234))
        run.save_checkpoint(path, 1234, kw_train, opt)
        ck = torch.load(path, map_location="cpu")
        assert set(ck) == {"global_step", "network_fn_state_dict", "network_fine_state_dict", "optimizer_state_dict"}
        assert list(ck["network_fn_state_dict"])[0] == "module.pts_linears.0.weight"
        kw2, _, start, gv2, opt2 = run.create_nerf(nerf_args(td, no_reload=False))
    assert start == 1234
    for a, b in zip(grad_vars, gv2):
        assert torch.equal(a, b)
    lr = run.update_learning_rate(opt2, nerf_args("x", lrate_decay=250), 125000)
    assert abs(lr - 5e-4 * 0.1 ** 0.5) < 1e-12 and opt2.param_groups[0]["lr"] == lr


def test_guidance_views_with_normal_maps(nerf):
    """cfg-5 shape at reduced size: V views -> rgb / disp / acc / depth [V,H,W,...] + normal maps [V,3,H,W] in [0, 1] units,
    equal to per-view render() + depth2normal."""
    run, kw_train, kw_test, grad_vars, opt = nerf
    from mvip_nerf_b200 import dist as md
    from mvip_nerf_b200.run_nerf_helpers import depth2normal
    poses = [pose(0.05 * i, 0.02 * i) for i in range(3)]
    H, W, focal = 48, 64, 60.0

    def fn(*a, **k):
        with torch.no_grad():
            return run.render(*a, chunk=8192, **k)
    out = md.render_views_sharded(fn, poses, H, W, focal, 1.2, 7.7, **kw_test)
    assert out["rgb_map"].shape == (3, H, W, 3) and out["normal"].shape == (3, 3, H, W)
    with torch.no_grad():
        want = run.render(H, W, focal, chunk=8192, c2w=poses[2], near=1.2, far=7.7, **kw_test)
    assert torch.equal(out["rgb_map"][2], want[0]) and torch.equal(out["depth_map"][2], want[3])
    K = [[focal, 0., W / 2], [0., focal, H / 2], [0., 0., 1.]]
    assert torch.equal(out["normal"][2:3], (depth2normal(want[3].contiguous(), K, 31) + 1) / 2)
    assert torch.isfinite(out["normal"]).all()


def fresh_nerf(run):
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "exp"))
        kw_train, kw_test, start, grad_vars, opt = run.create_nerf(nerf_args(td))
    load_seeded(kw_train["network_fn"], 200)
    load_seeded(kw_train["network_fine"], 201)
    return kw_train, kw_test, grad_vars, opt


def test_graphed_train_step_matches_eager():
    """GraphedTrainStep (two CUDA graphs per step) == the eager step: with the deterministic kwargs (no random draws) the
    parameters after 4 steps with a decaying learning rate agree to fp32 round-off, capture itself trains nothing, and the
    optimizer state / step counts interchange with the eager path."""
    from mvip_nerf_b200 import run
    from mvip_nerf_b200.graph import GraphedTrainStep, default_loss
    g = torch.Generator().manual_seed(3)
    N = 512
    ro = torch.zeros(N, 3)
    rd = torch.nn.functional.normalize(torch.randn(N, 3, generator=g) * 0.2 + torch.tensor([0., 0., -1.]), dim=-1)
    rays = torch.stack([ro, rd], 0).cuda()
    tgt = torch.rand(N, 3, generator=g).cuda()
    lrs = [5e-4, 4e-4, 3e-4, 2e-4]

    kw_a, _, gv_a, opt_a = fresh_nerf(run)
    kw_a = dict(kw_a, perturb=0., raw_noise_std=0.)
    losses_a = []
    for lr in lrs:
        opt_a.param_groups[0]["lr"] = lr
        opt_a.zero_grad(set_to_none=True)
        out = run.render(756, 1008, 767.2935, chunk=32768, rays=rays, near=1.2, far=7.7, **kw_a)
        loss = default_loss(*out, tgt, 1.0)
        loss.backward()
        opt_a.step()
        losses_a.append(float(loss.detach()))

    kw_b, _, gv_b, opt_b = fresh_nerf(run)
    kw_b = dict(kw_b, perturb=0., raw_noise_std=0.)
    before = [p.detach().clone() for p in gv_b]
    step = GraphedTrainStep(kw_b, opt_b, 756, 1008, 767.2935, N, 1.2, 7.7)
    step.rays.copy_(rays)
    step.target.copy_(tgt)
    step.capture()
    assert all(torch.equal(a, b) for a, b in zip(before, gv_b))            # capture trains nothing
    losses_b = []
    for lr in lrs:
        opt_b.param_groups[0]["lr"] = lr
        losses_b.append(float(step(rays, tgt)))
    np.testing.assert_allclose(losses_b, losses_a, rtol=1e-5)
    for a, b in zip(gv_a, gv_b):
        assert float((a - b).abs().max()) <= 1e-6 + 1e-4 * float(a.abs().max())
    sd = opt_b.state_dict()
    assert float(sd["state"][0]["step"]) == 4.0
    assert all(p.grad is g for p, g in zip(step._params, step._static_grads))
    # and back to the eager path: the fifth step continues from the graphed four
    opt_b.zero_grad(set_to_none=True)
    out = run.render(756, 1008, 767.2935, chunk=32768, rays=rays, near=1.2, far=7.7, **kw_b)
    default_loss(*out, tgt, 1.0).backward()
    opt_b.step()
    assert float(opt_b.state_dict()["state"][0]["step"]) == 5.0


def test_graphed_train_step_random_draws_and_learning():
    """train kwargs (perturb = 1, raw_noise_std = 1): every replay draws fresh randoms (losses differ step to step on fixed
    rays) and the loss goes down."""
    from mvip_nerf_b200 import run
    from mvip_nerf_b200.graph import GraphedTrainStep
    g = torch.Generator().manual_seed(4)
    N = 1024
    rd = torch.nn.functional.normalize(torch.randn(N, 3, generator=g) * 0.2 + torch.tensor([0., 0., -1.]), dim=-1)
    rays = torch.stack([torch.zeros(N, 3), rd], 0).pin_memory()
    tgt = torch.full((N, 3), 0.2).pin_memory()
    kw, _, gv, opt = fresh_nerf(run)
    step = GraphedTrainStep(dict(kw, near=1.2, far=7.7), opt, 756, 1008, 767.2935, N)     # bounds merged as train() does
    losses = [float(step(rays, tgt)) for _ in range(30)]
    assert len(set(losses[:5])) == 5
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < 0.7 * np.mean(losses[:5])


def guidance_loss(out):
    V, H, W = out["disp_map"].shape
    wgt = torch.linspace(0.5, 1.5, V * H * W, device=out["disp_map"].device).view(V, H, W)
    loss = ((out["rgb_map"] - 0.4) ** 2 * wgt[..., None]).mean() + 0.02 * (out["depth_map"] * wgt).mean()
    if "normal" in out:
        loss = loss + 0.05 * ((out["normal"] - 0.5) ** 2).mean()
    return loss


def test_sharded_guidance_views_gradients_match_direct(nerf):
    """dist.ShardedGuidanceViews (gather -> loss on rank 0 -> scatter -> deferred backward -> allreduce), world size 1:
    images, normal maps and parameter gradients equal the direct render() + depth2normal autograd graph."""
    run, kw_train, kw_test, grad_vars, opt = nerf
    from mvip_nerf_b200 import dist as md
    from mvip_nerf_b200.run_nerf_helpers import depth2normal
    poses = [pose(0.05 * i, 0.02 * i) for i in range(2)]
    H, W, focal = 20, 28, 26.0
    K = [[focal, 0., W / 2], [0., focal, H / 2], [0., 0., 1.]]
    # direct
    outs = [run.render(H, W, focal, chunk=4096, c2w=p, near=1.2, far=7.7, **kw_test) for p in poses]
    direct = {"rgb_map": torch.stack([o[0] for o in outs]), "disp_map": torch.stack([o[1] for o in outs]),
              "acc_map": torch.stack([o[2] for o in outs]), "depth_map": torch.stack([o[3] for o in outs])}
    direct["normal"] = torch.cat([(depth2normal(direct["depth_map"][v].contiguous(), K, 31) + 1) / 2 for v in range(2)], 0)
    gd = torch.autograd.grad(guidance_loss(direct), grad_vars, allow_unused=True)
    # sharded (deferred)
    for v in grad_vars:
        v.grad = None
    g = md.ShardedGuidanceViews(kw_test, poses, H, W, focal, 1.2, 7.7, chunk=256)
    out = g.forward()
    assert torch.equal(out["rgb_map"], direct["rgb_map"].detach()) and torch.equal(out["normal"], direct["normal"].detach())
    guidance_loss(out).backward()
    g.backward()
    for p, a in zip(grad_vars, gd):
        if a is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
        else:
            assert float((p.grad - a).abs().max()) <= 2e-4 * float(a.abs().max()) + 1e-12
