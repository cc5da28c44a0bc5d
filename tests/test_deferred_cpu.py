"""Host logic of the deferred back-propagation wrapper (mvip_nerf_b200.run._DeferredRays) on CPU:
render_rays is replaced by a small differentiable torch stand-in (the real one needs the CUDA library), so what is checked
is the chunk loop, the gradient slicing / accumulation and the replay of the random streams."""
import numpy as np
import pytest
import torch

from mvip_nerf_b200 import run


class TinyField(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(11, 16)
        self.b = torch.nn.Linear(16, 5)


def fake_render_rays(ray_batch, network_fn=None, perturb=0., **kw):
    h = torch.relu(network_fn.a(ray_batch))
    if perturb > 0.:
        h = h * (1. + torch.rand(h.shape))            # a random stream that backward has to replay
    o = network_fn.b(h)
    return {"rgb_map": torch.sigmoid(o[:, :3]), "disp_map": o[:, 3], "acc_map": torch.sigmoid(o[:, 4]), "depth_map": o[:, 3] * 2.,
            "weights": o.detach() * 0., "rgb0": torch.tanh(o[:, :3])}


def run_case(monkeypatch, perturb, chunk):
    monkeypatch.setattr(run, "render_rays", fake_render_rays)
    torch.manual_seed(0)
    net = TinyField()
    rays = torch.randn(37, 11)
    tgt = torch.rand(37, 3)
    params = list(net.parameters())

    def loss_of(ret):
        return ((ret["rgb_map"] - tgt) ** 2).mean() + 0.3 * ((ret["rgb0"] - tgt) ** 2).mean() + 0.1 * ret["depth_map"].abs().mean()

    # direct: everything in one autograd graph, same chunking (the random stream is drawn chunk by chunk)
    torch.manual_seed(123)
    direct = run.batchify_rays(rays, chunk, network_fn=net, perturb=perturb)
    gd = torch.autograd.grad(loss_of(direct), params)
    # deferred
    torch.manual_seed(123)
    holder = {}
    outs = run._DeferredRays.apply(rays, chunk, dict(network_fn=net, perturb=perturb), holder, *params)
    ret = dict(zip(holder["keys"], outs))
    assert not ret["weights"].requires_grad and ret["rgb_map"].requires_grad
    for k in direct:
        assert torch.equal(ret[k], direct[k].detach()), k
    state_before = torch.get_rng_state()
    gq = torch.autograd.grad(loss_of(ret), params)
    assert torch.equal(torch.get_rng_state(), state_before)        # the replay does not disturb the caller's stream
    for a, b in zip(gd, gq):
        np.testing.assert_allclose(b.numpy(), a.numpy(), rtol=1e-5, atol=1e-7)


def test_deferred_matches_direct_multi_chunk(monkeypatch):
    run_case(monkeypatch, 0., 8)


def test_deferred_replays_random_streams(monkeypatch):
    run_case(monkeypatch, 1., 16)


def test_deferred_single_chunk(monkeypatch):
    run_case(monkeypatch, 1., 64)


def test_png_writer_roundtrip(tmp_path):
    import struct
    import zlib
    img = (np.random.RandomState(0).rand(9, 13, 3) * 255).astype(np.uint8)
    p = str(tmp_path / "a.png")
    run._write_png(p, img)
    b = open(p, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    w, h, depth, ctype = struct.unpack(">IIBB", b[16:26])
    assert (w, h, depth, ctype) == (13, 9, 8, 2)
    n = struct.unpack(">I", b[33:37])[0]
    assert b[37:41] == b"IDAT"
    rows = np.frombuffer(zlib.decompress(b[41:41 + n]), np.uint8).reshape(9, 1 + 13 * 3)
    assert np.array_equal(rows[:, 1:].reshape(9, 13, 3), img) and not rows[:, 0].any()


def test_view_loop_host_helpers():
    """hwf scaling / intrinsics (run.py:1225-1236), pose convention flip (run.py:1435-1440), learning-rate decay (run.py:1031-1039)"""
    import argparse
    H, W, focal, K = run._scaled_hwf([567, 1008, 767.2935], 4)
    assert (H, W) == (141, 252) and abs(focal - 767.2935 / 4) < 1e-12
    assert np.allclose(K, [[focal, 0, 126.0], [0, focal, 70.5], [0, 0, 1]])
    assert run._scaled_hwf([567, 1008, 767.2935], 0)[:3] == (567, 1008, 767.2935)
    c = run.convert_pose(np.arange(16, dtype=np.float64).reshape(4, 4))
    assert np.array_equal(c[:, 0], [0, 4, 8, 12]) and np.array_equal(c[:, 1], [-1, -5, -9, -13]) and np.array_equal(c[:, 2], [-2, -6, -10, -14])
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=5e-4)
    lr = run.update_learning_rate(opt, argparse.Namespace(lrate=5e-4, lrate_decay=250), 250000)
    assert abs(lr - 5e-5) < 1e-15 and opt.param_groups[0]["lr"] == lr


def test_render_path_4view_pose_selection(monkeypatch):
    """every second pose of the +-4 neighbourhood of pose iter % 60 (run.py:1388-1392), rendered through render()"""
    seen = []

    def fake_render(H, W, focal, chunk=0, c2w=None, **kw):
        seen.append(float(c2w[0, 3]))
        z = torch.zeros(H, W)
        return [torch.zeros(H, W, 3), z, z, z, {}]
    monkeypatch.setattr(run, "render", fake_render)
    poses = torch.zeros(70, 3, 4)
    poses[:, 0, 3] = torch.arange(70, dtype=torch.float32)
    masks = np.arange(70)
    rgbs, disps, sel = run.render_path_4view(62, masks, poses, [8, 12, 10.0], 64, {}, render_factor=2)
    assert seen == [0.0, 2.0, 4.0, 6.0] and list(sel) == [0, 2, 4, 6] and rgbs.shape == (4, 4, 6, 3)     # iter = 2
    seen.clear()
    run.render_path_4view(127, masks, poses, [8, 12, 10.0], 64, {})
    assert seen == [3.0, 5.0, 7.0, 9.0, 11.0]                                                          # iter = 7


def test_patch_is_clipped_like_the_reference_slicing():
    """render(patch=...) past the image border: the reference slices rays_o[i:i+len1, j:j+len2] (run.py:1174), which clips
    silently; run._clip_patch must give the same extent (the first version raised)."""
    from mvip_nerf_b200 import run
    rng = np.random.RandomState(0)
    H, W = 37, 53
    img = np.zeros((H, W))
    for _ in range(200):
        i, j = int(rng.randint(0, H)), int(rng.randint(0, W))
        l1, l2 = int(rng.randint(1, 64)), int(rng.randint(1, 64))
        ci, cj, c1, c2 = run._clip_patch((i, j, l1, l2), H, W)
        assert (ci, cj) == (i, j)
        assert (c1, c2) == img[i:i + l1, j:j + l2].shape


def test_graphed_adam_refuses_several_param_groups():
    """FusedAdam's CUDA-graph form keeps one device-resident (lr, step) pair: more than one group must raise
    before anything is captured, not silently train every group with group 0's learning rate."""
    from mvip_nerf_b200.optim import FusedAdam
    a, b = torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2))
    opt = FusedAdam([{"params": [a]}, {"params": [b], "lr": 1e-2}], lr=5e-4)
    with pytest.raises(RuntimeError, match="single param group"):
        opt.graph_begin(torch.device("cpu"))
