"""CPU suite: world_size-2 gloo tests of the N>1 host logic (sharding, gather, gradient allreduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from mvip_nerf_b200 import dist as md
    r, w, _ = md.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # --- ray sharding + single gather: rank-local "render" = a pure per-ray function
    rays = torch.arange(n * 11, dtype=torch.float32).reshape(n, 11)

    def fake_render(rows):
        return {"rgb_map": rows[:, 0:3] * 2, "disp_map": rows[:, 3], "acc_map": rows[:, 4] + 1, "depth_map": rows[:, 5]}
    out = md.render_sharded(fake_render, rays)
    if rank == 0:
        want = fake_render(rays)
        for k in want:
            assert torch.equal(out[k], want[k]), k
    else:
        assert out is None
    lo, hi = md.shard_bounds(n)
    assert hi - lo in (n // world, n // world + 1)
    # --- gradient allreduce: each rank holds the gradient of its shard; the sum equals the full-batch gradient
    torch.manual_seed(0)
    lin_a, lin_b = torch.nn.Linear(11, 4), torch.nn.Linear(11, 2)
    full = (lin_a(rays) ** 2).sum() / n + (lin_b(rays) ** 2).sum() / n
    ga_full = torch.autograd.grad(full, list(lin_a.parameters()) + list(lin_b.parameters()))
    mine = md.shard_rows(rays)
    loss = (lin_a(mine) ** 2).sum() / n + (lin_b(mine) ** 2).sum() / n      # scaled by the GLOBAL count
    loss.backward()
    md.allreduce_grads([list(lin_a.parameters()), list(lin_b.parameters())])
    for p, g in zip(list(lin_a.parameters()) + list(lin_b.parameters()), ga_full):
        torch.testing.assert_close(p.grad, g, rtol=1e-5, atol=1e-5)
    # --- in-place path: gradients that are consecutive views of one flat buffer (what the MLP backward produces)
    ps = list(lin_a.parameters())
    flat = torch.full((sum(p.numel() for p in ps),), float(rank + 1))
    off = 0
    for p in ps:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    assert md._flat_view([p.grad for p in ps]).data_ptr() == flat.data_ptr()
    red = md.GradAllReducer(ps)
    red.start()
    assert red.in_place
    red.finish()
    assert torch.equal(flat, torch.full_like(flat, float(sum(range(1, world + 1)))))     # reduced in place, no copies
    assert all(p.grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr() for p in ps)
    # --- guidance batch: V views x H rows sharded as row bands (a rank's range may straddle two views), one gather
    V, H, W = 3, n, 4
    poses = [torch.eye(4)[:3, :4] * float(v + 1) for v in range(V)]

    def fake_view(H_, W_, focal, c2w=None, patch=None, near=0., far=1.):
        i0, j0, h, w_ = patch
        assert (j0, w_) == (0, W_)
        rows = torch.arange(i0, i0 + h, dtype=torch.float32)[:, None].expand(h, w_)
        cols = torch.arange(w_, dtype=torch.float32)[None, :].expand(h, w_)
        val = c2w[0, 0] * 1000 + rows * 10 + cols
        return [torch.stack([val, val + .25, val + .5], -1), val * 2, val * 3, val * 4, {}]
    out = md.render_views_sharded(fake_view, poses, H, W, 50., 1., 6., with_normals=False)
    bands = [b for r_ in range(world) for b in md.view_row_bands(V, H, r_, world)]
    assert sum(b[2] for b in bands) == V * H and bands[0][:2] == (0, 0)
    if rank == 0:
        for v in range(V):
            want = fake_view(H, W, 50., c2w=poses[v], patch=(0, 0, H, W))
            assert torch.equal(out["rgb_map"][v], want[0]) and torch.equal(out["disp_map"][v], want[1])
            assert torch.equal(out["acc_map"][v], want[2]) and torch.equal(out["depth_map"][v], want[3])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


class _Field(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(11, 12)
        self.b = torch.nn.Linear(12, 6)

    def _ordered_params(self):
        return list(self.parameters())


def _fake_rays_from_pose(H, W, focal, c2w, near, far, use_viewdirs=True, c2w_staticcam=None, patch=None, device=None, ndc=False,
                         ndc_near=1.):
    i0, j0, h, w = patch
    ii, jj = torch.meshgrid(torch.arange(i0, i0 + h, dtype=torch.float32), torch.arange(j0, j0 + w, dtype=torch.float32), indexing="ij")
    cols = [c2w[0, 3] + 0 * ii, ii / H, jj / W, torch.sin(ii + c2w[0, 3]), torch.cos(jj), ii * 0 + focal / 100, ii * 0 + near, ii * 0 + far,
            ii * jj / (H * W), ii * 0 + 1, jj * 0 - 1]
    return torch.stack(cols, -1).reshape(-1, 11)


def _fake_render_rays(ray_batch, network_fn=None, **kw):
    o = network_fn.b(torch.tanh(network_fn.a(ray_batch)))
    return {"rgb_map": torch.sigmoid(o[:, :3]), "disp_map": o[:, 3], "acc_map": torch.sigmoid(o[:, 4]), "depth_map": o[:, 5] + 3.}


def _image_loss(out):
    V, H, W = out["disp_map"].shape
    wgt = torch.linspace(0.5, 1.5, V * H * W).view(V, H, W)
    return ((out["rgb_map"] - 0.3) ** 2 * wgt[..., None]).mean() + 0.2 * (out["depth_map"] * wgt).mean() + 0.1 * (out["disp_map"] ** 2).mean()


def _guidance_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from mvip_nerf_b200 import dist as md
    from mvip_nerf_b200 import ops, run
    md.init_from_env(backend="gloo")
    ops.rays_from_pose = _fake_rays_from_pose
    run.render_rays = _fake_render_rays
    torch.manual_seed(0)
    net = _Field()
    V, H, W = 3, 7, 5                       # 21 rows over 2 ranks: 11 + 10, rank 0's range straddles views 0 and 1
    poses = [torch.eye(4)[:3, :4] * 1.0 + v for v in range(V)]
    kw = {"network_fn": net, "use_viewdirs": True, "ndc": False}
    g = md.ShardedGuidanceViews(kw, poses, H, W, 50., 1., 6., chunk=16, with_normals=False)
    out = g.forward()
    if rank == 0:
        assert out["rgb_map"].shape == (V, H, W, 3)
        _image_loss(out).backward()
    else:
        assert out is None
    g.backward()
    # single-process reference: all rows at once, plain autograd
    rays = torch.cat([_fake_rays_from_pose(H, W, 50., poses[v], 1., 6., patch=(0, 0, H, W)) for v in range(V)], 0)
    ret = _fake_render_rays(rays, network_fn=net)
    full = {k: v.view(V, H, W, *v.shape[1:]) for k, v in ret.items()}
    want = torch.autograd.grad(_image_loss(full), list(net.parameters()))
    for p, gw in zip(net.parameters(), want):
        torch.testing.assert_close(p.grad, gw, rtol=1e-5, atol=1e-7)          # every rank holds the FULL gradient
    dist.barrier()
    dist.destroy_process_group()


def _gradsync_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from mvip_nerf_b200 import dist as md
    md.init_from_env(backend="gloo")
    torch.manual_seed(0)
    fine, coarse, unused = torch.nn.Linear(5, 3), torch.nn.Linear(5, 3), torch.nn.Linear(2, 2)
    groups = [list(fine.parameters()), list(coarse.parameters()), list(unused.parameters())]
    sync = md.GradSync(groups)
    x = torch.arange(20, dtype=torch.float32).view(4, 5) / 10 + rank
    for step in range(2):                       # second step: hooks re-arm, .grad accumulates in place
        for ps in groups:
            for p in ps:
                p.grad = None
        loss = (fine(x) ** 2).mean() * (step + 1) + coarse(x).sum() * 0.5
        loss.backward()
        assert sync.started[:2] == [True, True] and not sync.started[2]      # issued from the hooks, before finish()
        sync.finish()
        want = []
        for r in range(world):
            xr = torch.arange(20, dtype=torch.float32).view(4, 5) / 10 + r
            want.append(torch.autograd.grad((fine(xr) ** 2).mean() * (step + 1) + coarse(xr).sum() * 0.5,
                                            groups[0] + groups[1]))
        for i, p in enumerate(groups[0] + groups[1]):
            torch.testing.assert_close(p.grad, sum(w[i] for w in want), rtol=1e-6, atol=1e-6)
        for p in groups[2]:                     # a network without gradients still takes part: zeros
            assert p.grad is not None and float(p.grad.abs().max()) == 0.0
    sync.remove()
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_gradsync_overlapped_buckets():
    mp.spawn(_gradsync_worker, args=(2, _free_port()), nprocs=2, join=True)


def _gradsync_single_bucket_worker(rank, world, port):
    """overlap=False + the gradients of both networks back to back in one buffer (what graph.GraphedTrainStep arranges through
    ops.grad_arena): finish() must issue ONE in-place collective over the whole range; a gap in the layout falls back to
    one bucket per network."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from mvip_nerf_b200 import dist as md
    md.init_from_env(backend="gloo")
    torch.manual_seed(0)
    fine, coarse = torch.nn.Linear(5, 3), torch.nn.Linear(5, 3)
    groups = [list(fine.parameters()), list(coarse.parameters())]
    params = groups[0] + groups[1]
    sync = md.GradSync(groups, overlap=False)
    calls = []
    real = dist.all_reduce

    def counting(t, *a, **k):
        calls.append(t.numel())
        return real(t, *a, **k)
    md.dist.all_reduce = counting
    n = sum(p.numel() for p in params)
    for gap in (0, 3):
        arena = torch.zeros(n + gap)
        off = 0
        for i, p in enumerate(params):
            if i == len(groups[0]):
                off += gap                      # gap > 0: the two networks are not adjacent
            p.grad = arena[off:off + p.numel()].view_as(p)
            p.grad.fill_(float(rank + 1) * (i + 1))
            off += p.numel()
        del calls[:]
        sync.finish()
        assert not any(sync.started)
        assert calls == ([n] if gap == 0 else [sum(p.numel() for p in groups[0]), sum(p.numel() for p in groups[1])]), calls
        for i, p in enumerate(params):
            assert float((p.grad - 3.0 * (i + 1)).abs().max()) == 0.0          # 1 + 2 = sum over the two ranks
            assert p.grad.untyped_storage().data_ptr() == arena.untyped_storage().data_ptr()      # in place
    md.dist.all_reduce = real
    sync.remove()
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_gradsync_single_bucket():
    mp.spawn(_gradsync_single_bucket_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_grad_arena_carves_consecutive_buffers():
    from mvip_nerf_b200 import ops
    buf = torch.zeros(10)
    ops.grad_arena.begin(buf)
    a = ops.grad_arena.take(4, buf.device, zero=False)
    b = ops.grad_arena.take(6, buf.device, zero=True)
    c = ops.grad_arena.take(1, buf.device, zero=True)          # exhausted: falls back to a fresh tensor
    ops.grad_arena.end()
    assert a.data_ptr() == buf.data_ptr() and b.data_ptr() == buf.data_ptr() + 16
    assert c.untyped_storage().data_ptr() != buf.untyped_storage().data_ptr()
    d = ops.grad_arena.take(3, buf.device, zero=False)         # no arena installed
    assert d.untyped_storage().data_ptr() != buf.untyped_storage().data_ptr()


def test_world2_gloo_sharded_guidance_views_with_gradients():
    mp.spawn(_guidance_worker, args=(2, _free_port()), nprocs=2, join=True)


@pytest.mark.parametrize("n", [10, 7])
def test_world2_gloo_shard_gather_allreduce(n):
    mp.spawn(_worker, args=(2, _free_port(), n), nprocs=2, join=True)


def test_shard_bounds_cover_everything():
    from mvip_nerf_b200 import dist as md
    for n in (0, 1, 7, 8, 762048, 65536):
        for w in (1, 2, 4, 8):
            b = [md.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
