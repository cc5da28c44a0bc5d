"""GPU suite: fused PE + NeRF MLP (tcgen05) against the reference's NeRF.forward (golden) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc

pytestmark = pytest.mark.gpu

# bf16 operands / fp32 accumulate.  Stated tolerances (random-init network):
RAW_ATOL = 2e-2        # max-abs on raw (rgb logits, sigma)
RGB_ATOL = 2e-3        # max-abs on composited rgb
DEPTH_RTOL = 1e-2      # relative on composited depth


@pytest.fixture(scope="module")
def ops():
    from mvip_nerf_b200 import ops as o
    return o


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def pack(ops, p):
    return ops.mlp_pack([cu(p[n]) for n in ops.PARAM_ORDER])


def test_pack_layout_roundtrip(ops):
    p = orc.init_params(3)
    blob = pack(ops, p).cpu().numpy()
    w = blob[:34 * 32768 + 5 * 16384].view(np.uint16)
    # chunk 1 = pts_linears.1 columns 0..63: element (n, k) at byte (n>>3)*1024 + (n&7)*128 + ((k>>3 ^ n&7)<<4) + (k&7)*2
    W1 = torch.from_numpy(p["pts_linears.1.weight"]).bfloat16().view(torch.int16).numpy().view(np.uint16)
    for n, k in [(0, 0), (5, 9), (200, 63), (255, 17)]:
        off = 32768 + (n >> 3) * 1024 + (n & 7) * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2
        assert w[off // 2] == W1[n, k]
    small = blob[34 * 32768 + 5 * 16384 + 34 * 32768:].view(np.float32)
    assert np.array_equal(small[:256], p["pts_linears.0.bias"])
    assert np.array_equal(small[2432:2688], p["alpha_linear.weight"][0])
    assert np.array_equal(small[2692:2692 + 384], p["rgb_linear.weight"].ravel())


def test_mlp_forward_points_vs_reference(ops, golden):
    fx = golden("nerf_mlp")
    p = orc.init_params(int(fx["param_seed"]))
    emb = cu(fx["embedded"])
    # the drop-in NeRF.forward path: raw pts / viewdirs are columns 0:3 and 63:66 of the embedded input
    raw = ops.mlp_forward(pack(ops, p), pts=emb[:, 0:3], dirs=emb[:, 63:66])
    err = np.abs(raw.cpu().numpy() - fx["out"]).max()
    assert err < RAW_ATOL, err


@pytest.mark.parametrize("P", [1, 127, 128, 129, 300, 5000])
def test_mlp_forward_ragged_sizes_vs_oracle(ops, P):
    rng = np.random.RandomState(P)
    p = orc.init_params(11)
    pts = ((rng.rand(P, 3) * 2 - 1) * 6).astype(np.float32)
    vd = rng.randn(P, 3).astype(np.float32)
    vd /= np.linalg.norm(vd, axis=-1, keepdims=True)
    want = orc.nerf_forward(p, np.concatenate([orc.embed(pts, 10), orc.embed(vd, 4)], -1))
    raw = ops.mlp_forward(pack(ops, p), pts=cu(pts), dirs=cu(vd))
    assert raw.shape == (P, 4)
    err = np.abs(raw.cpu().numpy() - want).max()
    assert err < RAW_ATOL, err


def test_mlp_forward_rays_mode_and_stash_do_not_change_result(ops):
    rng = np.random.RandomState(5)
    N, S = 96, 64
    p = orc.init_params(12)
    ro, rd = orc.get_rays(756, 1008, 767.2935, np.eye(4, dtype=np.float32)[:3, :4])
    idx = rng.permutation(756 * 1008)[:N]
    rays = orc.make_ray_batch(ro.reshape(-1, 3)[idx], rd.reshape(-1, 3)[idx], 1.2, 7.7369)
    z = orc.sample_coarse(rays, orc.linspace_f32(0, 1, S), rng.rand(N, S).astype(np.float32), True)
    want = orc.run_network(p, orc.points(rays, z), rays[:, -3:])
    blob = pack(ops, p)
    raw = ops.mlp_forward(blob, rays=cu(rays), z_vals=cu(z))
    err = np.abs(raw.view(N, S, 4).cpu().numpy() - want).max()
    assert err < RAW_ATOL, err
    raw2, stash = ops.mlp_forward(blob, rays=cu(rays), z_vals=cu(z), want_stash=True)
    assert torch.equal(raw, raw2)
    # stash chunk 0 of tile 0 is the PE tile image: row 0, first 3 bf16 = the point itself
    pe0 = stash[:6].cpu().view(torch.bfloat16).float().numpy()
    np.testing.assert_allclose(pe0, orc.points(rays, z)[0, 0], rtol=1e-2)


def test_mlp_forward_is_deterministic_and_tile_independent(ops):
    # size-independent property at cfg-2 size: a point's output does not depend on its tile / neighbours
    rng = np.random.RandomState(9)
    P = 4096 * 64
    p = orc.init_params(13)
    pts = ((rng.rand(P, 3) * 2 - 1) * 4).astype(np.float32)
    vd = rng.randn(P, 3).astype(np.float32)
    blob = pack(ops, p)
    a = ops.mlp_forward(blob, pts=cu(pts), dirs=cu(vd))
    b = ops.mlp_forward(blob, pts=cu(pts), dirs=cu(vd))
    assert torch.equal(a, b)
    perm = torch.randperm(P, device="cuda")
    c = ops.mlp_forward(blob, pts=cu(pts)[perm].contiguous(), dirs=cu(vd)[perm].contiguous())
    assert torch.equal(a[perm], c)
    sub = rng.permutation(P)[:2048]
    want = orc.nerf_forward(p, np.concatenate([orc.embed(pts[sub], 10), orc.embed(vd[sub], 4)], -1))
    assert np.abs(a.cpu().numpy()[sub] - want).max() < RAW_ATOL
