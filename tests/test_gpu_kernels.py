"""GPU suite (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the
golden vectors produced by the reference and against the numpy oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from mvip_nerf_b200 import ops as o
    return o


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------ sampler
@pytest.mark.parametrize("lindisp", [0, 1])
@pytest.mark.parametrize("perturb", [0, 1])
def test_sample_coarse_bit_exact_vs_reference(ops, golden, lindisp, perturb):
    fx = golden("sample_coarse")
    key = "lindisp%d_perturb%d" % (lindisp, perturb)
    z = ops.sample_coarse(cu(fx[key + "_rays"]), cu(fx["t_vals"]), cu(fx["t_rand"]) if perturb else None, bool(lindisp))
    assert np.array_equal(npy(z), fx[key + "_z"])


@pytest.mark.parametrize("dist", ["uniform", "peaky", "sparse", "zeros"])
@pytest.mark.parametrize("mode", ["det", "rand"])
def test_sample_pdf_bit_exact_vs_reference(ops, golden, dist, mode):
    fx = golden("sample_pdf")
    u = fx["u_det"] if mode == "det" else fx["u_rand"]
    samples, inds, cdf = ops.sample_pdf(cu(fx["bins"]), cu(fx["w_" + dist]), cu(u), want_inds=True, want_cdf=True)
    tag = "%s_%s" % (dist, mode)
    assert np.array_equal(npy(cdf), fx["cdf_" + tag])
    assert inds.dtype == torch.int64
    assert np.array_equal(npy(inds), fx["inds_" + tag])          # sample indices: bit-exact
    assert np.array_equal(npy(samples), fx["samples_" + tag])


def test_sample_pdf_large_random_vs_oracle(ops):
    rng = np.random.RandomState(7)
    N = 20000
    z = np.sort(1.2 + 6.5 * rng.rand(N, 64).astype(np.float32), -1)
    bins = (0.5 * (z[:, 1:] + z[:, :-1])).astype(np.float32)
    w = (rng.rand(N, 62) ** 8).astype(np.float32)
    w[::7] *= (rng.rand(*w[::7].shape) > 0.9)
    w[::13] = 0
    w[5::17] *= 1e-12                                   # tiny pdf entries exercise the sequential-scan fallback
    w[5::17, 3] = 1.0
    u = rng.rand(N, 64).astype(np.float32)
    u[:, 0] = 0.0
    u[:, 1] = 1.0
    want = orc.sample_pdf(bins, w, u)
    samples, inds, cdf = ops.sample_pdf(cu(bins), cu(w), cu(u), want_inds=True, want_cdf=True)
    assert np.array_equal(npy(cdf), want["cdf"])
    assert np.array_equal(npy(inds), want["inds"])
    assert np.array_equal(npy(samples), want["samples"])


def test_sample_fine_merge_vs_reference(ops, golden):
    fx = golden("merge")
    sp = golden("sample_pdf")
    N = fx["z"].shape[0]
    w = np.zeros((N, 64), np.float32)
    w[:, 1:-1] = sp["w_peaky"]
    out = ops.sample_fine(cu(fx["z"]), cu(w), cu(sp["u_rand"]), want_inds=True)
    assert np.array_equal(npy(out["z_samples"]), fx["z_samples"])
    assert np.array_equal(npy(out["z_merged"]), fx["merged"])       # == torch.sort(cat) of the reference
    np.testing.assert_allclose(npy(out["z_std"]), fx["z_std"], rtol=2e-5, atol=1e-6)
    # det mode: u row shared by all rays, samples already sorted
    out = ops.sample_fine(cu(fx["z"]), cu(w), cu(sp["u_det"]), want_inds=True)
    want = orc.fine_samples(fx["z"], w, sp["u_det"])
    assert np.array_equal(npy(out["z_merged"]), want["z_merged"])
    assert np.array_equal(npy(out["inds"]), want["inds"])


@pytest.mark.parametrize("mode", ["rand", "det", "rand_sorted_rows"])
def test_sample_fine_64x64_fast_path_vs_oracle(ops, mode):
    """The register kernel of the render_rays shape (64 depths, 64 draws): samples / inds / merged depths bit-exact against the
    oracle on 20k fuzzed rays, including rows whose tiny pdf entries take the sequential-scan fallback, ties (u = 0, u = 1,
    repeated draws) and all-zero weights; and equal to the generic kernel on the same inputs."""
    rng = np.random.RandomState(11)
    N = 20000
    z = np.sort(1.2 + 6.5 * rng.rand(N, 64).astype(np.float32), -1)
    z[::11, 5] = z[::11, 4]                                 # repeated depths
    w = (rng.rand(N, 64) ** 8).astype(np.float32)
    w[::7] *= (rng.rand(*w[::7].shape) > 0.9)
    w[::13] = 0
    w[5::17] *= 1e-12                                       # tiny pdf entries exercise the sequential-scan fallback
    w[5::17, 3] = 1.0
    if mode == "det":
        u = orc.linspace_f32(0, 1, 64)
    else:
        u = rng.rand(N, 64).astype(np.float32)
        u[:, 0] = 0.0
        u[:, 1] = 1.0
        u[::5, 7] = u[::5, 6]
        if mode == "rand_sorted_rows":
            u = np.sort(u, -1)
    want = orc.fine_samples(z, w, u)
    out = ops.sample_fine(cu(z), cu(w), cu(u), want_inds=True)
    assert np.array_equal(npy(out["inds"]), want["inds"])
    assert np.array_equal(npy(out["z_samples"]), want["z_samples"])
    assert np.array_equal(npy(out["z_merged"]), want["z_merged"])
    np.testing.assert_allclose(npy(out["z_std"]), want["z_std"], rtol=2e-5, atol=1e-6)
    # ragged row counts: the kernel carries four rays per warp, the last warp may hold 1 - 3
    for n in (1, 2, 3, 5, 37, 255):
        part = ops.sample_fine(cu(z[:n]), cu(w[:n]), cu(u if mode == "det" else u[:n]), want_inds=True)
        for key in ("inds", "z_samples", "z_merged", "z_std"):
            assert torch.equal(part[key], out[key][:n]), (n, key)
    # the generic kernel (taken for any other shape) on an embedding of the same problem: 64 draws -> 65 with a duplicate
    if mode != "det":
        u65 = np.concatenate([u, u[:, -1:]], -1)
        gen = ops.sample_fine(cu(z[:512]), cu(w[:512]), cu(u65[:512]), want_inds=True)
        assert np.array_equal(npy(gen["z_samples"])[:, :64], npy(out["z_samples"])[:512])


def test_sample_coarse_shapes_vs_oracle(ops):
    """warp-per-ray kernels (S = 32 / 64 / 128) and the generic one (S = 48) against the oracle, both lindisp modes"""
    rng = np.random.RandomState(5)
    N = 3001
    rays = rng.randn(N, 11).astype(np.float32)
    rays[:, 6] = 0.5 + rng.rand(N)
    rays[:, 7] = rays[:, 6] + 0.1 + 6 * rng.rand(N)
    for S in (32, 48, 64, 128):
        t = orc.linspace_f32(0, 1, S)
        tr = rng.rand(N, S).astype(np.float32)
        for lindisp in (False, True):
            for use_tr in (False, True):
                want = orc.sample_coarse(rays, t, tr if use_tr else None, lindisp)
                got = ops.sample_coarse(cu(rays), cu(t), cu(tr) if use_tr else None, lindisp)
                assert np.array_equal(npy(got), want), (S, lindisp, use_tr)


def test_sample_shapes_rejected(ops):
    z = torch.zeros(4, 8, device="cuda")
    with pytest.raises(RuntimeError, match="n_samples"):
        ops.sample_fine(z, z, torch.linspace(0, 1, 8, device="cuda"))
    assert ops.sample_coarse(torch.zeros(0, 11, device="cuda"), torch.linspace(0, 1, 64, device="cuda")).shape == (0, 64)


# ------------------------------------------------------------------------------------------ compositing
@pytest.mark.parametrize("S", [64, 128])
@pytest.mark.parametrize("white", [0, 1])
@pytest.mark.parametrize("use_noise", [0, 1])
def test_composite_forward_vs_reference(ops, golden, S, white, use_noise):
    fx = golden("raw2outputs")
    tag = "S%d_w%d_n%d" % (S, white, use_noise)
    noise = cu(fx["S%d_noise" % S]) if use_noise else None
    rgb, disp, acc, w, depth, alpha = ops.composite_forward(cu(fx["S%d_raw" % S]), cu(fx["S%d_z" % S]),
                                                            cu(fx["S%d_rays_d" % S]), noise, bool(white), need_alpha=True)
    for got, name in [(rgb, "rgb"), (acc, "acc"), (w, "weights"), (depth, "depth"), (alpha, "alpha"), (disp, "disp")]:
        ref = fx[tag + "_" + name]
        got = npy(got)
        assert np.array_equal(np.isnan(ref), np.isnan(got)), name
        # fp32 compositing: 1e-5 relative (north_star); atol covers entries that are ~0
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6, equal_nan=True, err_msg=name)


@pytest.mark.parametrize("S", [64, 128])
@pytest.mark.parametrize("white", [0, 1])
@pytest.mark.parametrize("use_noise", [0, 1])
def test_composite_backward_vs_reference_autograd(ops, golden, S, white, use_noise):
    fx = golden("raw2outputs")
    tag = "S%d_w%d_n%d" % (S, white, use_noise)
    noise = cu(fx["S%d_noise" % S]) if use_noise else None
    d_raw = ops.composite_backward(cu(fx["S%d_raw" % S]), cu(fx["S%d_z" % S]), cu(fx["S%d_rays_d" % S]), noise,
                                   bool(white), False, cu(fx[tag + "_g_rgb"]), cu(fx[tag + "_g_disp"]),
                                   cu(fx[tag + "_g_acc"]), cu(fx[tag + "_g_depth"]), cu(fx[tag + "_g_weights"]))
    ref = fx[tag + "_d_raw"]
    got = npy(d_raw)
    assert np.array_equal(np.isnan(ref), np.isnan(got))
    ok = ~np.isnan(ref)
    assert np.abs(got[ok] - ref[ok]).max() <= 2e-5 * np.abs(ref[ok]).max()


@pytest.mark.parametrize("S", [2, 7, 33, 48, 200])
def test_composite_ragged_sample_counts_vs_oracle(ops, S):
    rng = np.random.RandomState(S)
    N = 37
    raw = rng.randn(N, S, 4).astype(np.float32)
    z = np.sort(1 + 5 * rng.rand(N, S).astype(np.float32), -1)
    rd = rng.randn(N, 3).astype(np.float32)
    want = orc.raw2outputs(raw, z, rd, None, True)
    rgb, disp, acc, w, depth, _ = ops.composite_forward(cu(raw), cu(z), cu(rd), None, True)
    np.testing.assert_allclose(npy(w), want["weights"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(npy(rgb), want["rgb_map"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(npy(depth), want["depth_map"], rtol=1e-5, atol=1e-6)
    g = [rng.randn(N, 3), rng.randn(N), rng.randn(N), rng.randn(N), rng.randn(N, S)]
    g = [x.astype(np.float32) for x in g]
    want_b = orc.raw2outputs_backward(raw, z, rd, None, True, *g)
    got_b = npy(ops.composite_backward(cu(raw), cu(z), cu(rd), None, True, False, *[cu(x) for x in g]))
    assert np.abs(got_b - want_b).max() <= 3e-5 * np.abs(want_b).max()


def test_composite_linearity_at_full_size(ops):
    # size-independent property at the cfg-3 chunk size: rgb_map is linear in sigmoid(rgb) for fixed sigma,
    # acc + transmittance of the last sample == 1, weights >= 0
    N, S = 32768, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    raw = torch.randn(N, S, 4, device="cuda", generator=g)
    z = torch.sort(1.2 + 6.5 * torch.rand(N, S, device="cuda", generator=g), -1)[0]
    rd = torch.randn(N, 3, device="cuda", generator=g)
    rgb, disp, acc, w, depth, alpha = ops.composite_forward(raw, z, rd, None, False, need_alpha=True)
    assert float(w.min()) >= 0 and float(acc.max()) <= 1 + 1e-5
    raw2 = raw.clone()
    raw2[..., :3] = 50.0                                  # sigmoid -> 1: rgb_map == acc
    rgb2, _, acc2, _, _, _ = ops.composite_forward(raw2, z, rd, None, False)
    torch.testing.assert_close(rgb2, acc2[:, None].expand(-1, 3), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(acc2, acc)
    torch.testing.assert_close(depth, (w * z).sum(-1), rtol=2e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------ normal map
def test_normal_map_vs_reference(ops, golden):
    fx = golden("normal_map")
    K = fx["K"]
    args = (float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]))
    n = npy(ops.normal_forward(cu(fx["depth"]), *args, k=31))
    # the reference's own fp32 path differs from its fp64 path by ~1e-5; tolerance 1e-4 abs
    np.testing.assert_allclose(n, fx["normal_f32"], atol=1e-4)
    np.testing.assert_allclose(n, fx["normal_f64"], atol=2e-5)
    gd = npy(ops.normal_backward(cu(fx["depth"]), *args, cu(fx["g_normal_f64"].astype(np.float32)), k=31))
    ref = fx["d_depth_f64"]
    assert np.abs(gd - ref).max() <= 1e-4 * np.abs(ref).max()


def test_normal_map_plane_property(ops):
    # a fronto-parallel plane z = c has normal M^-1 s = (0, 0, 1/c) everywhere, any window, any size
    H, W = 512, 512
    depth = torch.full((H, W), 4.0, device="cuda")
    n = ops.normal_forward(depth, 500.0, 500.0, W / 2, H / 2, k=31)
    torch.testing.assert_close(n[2], torch.full((H, W), 0.25, device="cuda"), atol=1e-5, rtol=0)
    assert float(n[:2].abs().max()) < 1e-4


# ------------------------------------------------------------------------------------------ tcgen05 building blocks
@pytest.mark.parametrize("N,K", [(256, 64), (256, 256), (128, 128), (64, 64)])
def test_umma_k_major(ops, N, K):
    g = torch.Generator(device="cuda").manual_seed(N + K)
    a = torch.randn(128, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g)
    out = ops.selftest_umma(0, a, b)
    ref = a.bfloat16().float() @ b.bfloat16().float().t()
    torch.testing.assert_close(out, ref, atol=1e-2, rtol=1e-3)


@pytest.mark.parametrize("N,K", [(256, 64), (256, 128), (64, 64), (128, 256)])
def test_umma_mn_major(ops, N, K):
    g = torch.Generator(device="cuda").manual_seed(N + K + 1)
    a = torch.randn(K, 128, device="cuda", generator=g)
    b = torch.randn(K, N, device="cuda", generator=g)
    ref = a.bfloat16().float().t() @ b.bfloat16().float()
    out1 = ops.selftest_umma(1, a, b)
    err1 = float((out1 - ref).abs().max())
    if err1 > 1e-2:                                      # diagnostic: which LBO/SBO convention the hardware uses
        out2 = ops.selftest_umma(2, a, b)
        err2 = float((out2 - ref).abs().max())
        pytest.fail("MN-major convention 1 wrong (err %.3g); swapped convention err %.3g" % (err1, err2))


@pytest.mark.gpu
def test_fused_adam_matches_torch_adam(ops):
    """mvip_adam_step / optim.FusedAdam == torch.optim.Adam (run.py:1536-1537) over several steps, incl. a changed lr
    (run.py:1031-1039) and state_dict interchange."""
    from mvip_nerf_b200.optim import FusedAdam
    g = torch.Generator(device="cuda").manual_seed(5)
    shapes = [(256, 63), (256,), (256, 319), (1, 256), (1,), (3, 128), (3,)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    oa = FusedAdam(ours, lr=5e-4, betas=(0.9, 0.999))
    ra = torch.optim.Adam(ref, lr=5e-4, betas=(0.9, 0.999))
    for it in range(6):
        if it == 3:
            for grp in oa.param_groups + ra.param_groups:
                grp["lr"] = 2e-4
            import copy
            sd = copy.deepcopy(oa.state_dict())       # torch.optim.Adam <-> FusedAdam checkpoints interchange (deep copy:
                                                      # load_state_dict would otherwise alias our moment buffers)
            ra2 = torch.optim.Adam(ref, lr=2e-4, betas=(0.9, 0.999))
            ra2.load_state_dict(sd)
            ra = ra2
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, device="cuda", generator=g) * (10.0 ** (it - 3))
            a.grad = gr.clone()
            b.grad = gr.clone()
        e0 = ops.param_epoch
        oa.step()
        ra.step()
        assert ops.param_epoch == e0 + 1
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), float((a - b).abs().max())
    assert int(oa.state[ours[0]]["step"]) == 6


@pytest.mark.gpu
@pytest.mark.parametrize("use_viewdirs,static,patch", [(True, False, None), (True, True, (5, 7, 40, 33)), (False, False, None)])
def test_rays_from_pose_bit_exact_vs_reference_cpu(ops, use_viewdirs, static, patch):
    """mvip_rays_from_pose == get_rays (run_nerf_helpers.py:249-260) + the batch assembly of render() (run.py:1171-1207),
    evaluated with the reference's torch ops on the CPU (the bit-exactness authority)."""
    from mvip_nerf_b200.run_nerf_helpers import get_rays
    H, W, focal, near, far = 61, 83, 77.31, 1.2, 7.7369
    rng = np.random.RandomState(3)
    def pose():
        q, _ = np.linalg.qr(rng.randn(3, 3))
        return torch.from_numpy(np.concatenate([q, rng.randn(3, 1)], 1).astype(np.float32))
    c2w, c2s = pose(), (pose() if static else None)
    ro, rd = get_rays(H, W, focal, c2w)                       # CPU tensors -> torch CPU arithmetic
    vd = rd / torch.norm(rd, dim=-1, keepdim=True)
    if static:
        ro, rd = get_rays(H, W, focal, c2s)
    if patch is not None:
        i, j, l1, l2 = patch
        ro, rd, vd = ro[i:i + l1, j:j + l2], rd[i:i + l1, j:j + l2], vd[i:i + l1, j:j + l2]
    ones = torch.ones_like(rd[..., :1])
    cols = [ro, rd, near * ones, far * ones] + ([vd] if use_viewdirs else [])
    want = torch.cat(cols, -1).reshape(-1, 11 if use_viewdirs else 8).numpy()
    got = ops.rays_from_pose(H, W, focal, c2w, near, far, use_viewdirs=use_viewdirs, c2w_staticcam=c2s, patch=patch).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want), float(np.abs(got - want).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["plain", "ndc", "ndc_static", "ndc_patch", "ndc_nodirs", "static", "patch"])
def test_ray_batch_bit_exact_vs_reference_golden(ops, golden, name):
    """mvip_rays_from_pose_ndc / mvip_rays_pack == the [N, 8|11] batch the unmodified reference render() assembled on CPU
    (tests/golden/ray_batch.npz, oracle/make_golden_rays.py): get_rays, patch, c2w_staticcam, viewdirs, ndc_rays, near/far."""
    from test_oracle_golden import ray_case
    g = golden("ray_batch")
    c = ray_case(g, name)
    whole = c["patch"] == (0, 0, c["H"], c["W"])
    got = ops.rays_from_pose(c["H"], c["W"], c["focal"], cu(c["c2w"]), c["near"], c["far"], use_viewdirs=c["use_viewdirs"],
                             c2w_staticcam=None if c["c2w_static"] is None else cu(c["c2w_static"]),
                             patch=None if whole else c["patch"], ndc=c["ndc"])
    assert np.array_equal(npy(got), g[name + "_batch"])
    if name + "_rays_batch" in g:
        got = ops.rays_pack(cu(g[name + "_rays_o"]), cu(g[name + "_rays_d"]), c["near"], c["far"], use_viewdirs=c["use_viewdirs"],
                            ndc=c["ndc"], H=c["H"], W=c["W"], focal=c["focal"])
        assert np.array_equal(npy(got), g[name + "_rays_batch"])


@pytest.mark.gpu
def test_render_ndc_and_rays_entry_use_fused_batch(ops, golden):
    """run.render(ndc=True, c2w=...) and run.render(rays=...) go through the one-launch batch kernels and match the torch route."""
    from mvip_nerf_b200 import run
    seen = {}

    def fake_batchify(rays_flat, chunk, **kw):
        seen["rays"] = rays_flat
        n = rays_flat.shape[0]
        z = torch.zeros(n, device=rays_flat.device)
        return {"rgb_map": torch.zeros(n, 3, device=rays_flat.device), "disp_map": z, "acc_map": z, "depth_map": z}
    orig, run.batchify_rays = run.batchify_rays, fake_batchify
    try:
        g = golden("ray_batch")
        c2w = cu(g["ndc_c2w"])
        l0 = ops.launch_count
        out = run.render(37, 53, 41.7, c2w=c2w, ndc=True, near=0., far=1., use_viewdirs=True)
        assert ops.launch_count == l0 + 1 and out[0].shape == (37, 53, 3)
        assert np.array_equal(npy(seen["rays"]), g["ndc_batch"])
        run.render(37, 53, 41.7, rays=torch.stack([cu(g["ndc_rays_o"]), cu(g["ndc_rays_d"])], 0), ndc=True, near=0., far=1., use_viewdirs=True)
        assert ops.launch_count == l0 + 2
        assert np.array_equal(npy(seen["rays"]), g["ndc_rays_batch"])
    finally:
        run.batchify_rays = orig


@pytest.mark.parametrize("S,white,noise", [(64, True, True), (128, True, False), (128, False, True)])
def test_fused_mse_losses_match_img2mse_autograd(ops, S, white, noise):
    """img2mse on rgb and on disp fused into the compositing kernels (DS_NeRF/run.py:1000-1027,
    run_nerf_helpers.py:15): loss value vs the fp64 oracle maps, gradient w.r.t. raw vs torch autograd of the unfused maps
    (the unfused backward itself is held against the reference's autograd above)."""
    from mvip_nerf_b200 import run_nerf_helpers as h
    rng = np.random.RandomState(S + white)
    N = 3001                                               # ragged: not a multiple of the rays per block
    raw_np = rng.randn(N, S, 4).astype(np.float32)
    raw_np[..., 3] += 0.5                                   # some density on every ray: finite disp
    z_np = np.sort(1.2 + 6.5 * rng.rand(N, S).astype(np.float32), -1)
    rd_np = rng.randn(N, 3).astype(np.float32)
    nz_np = rng.rand(N, S).astype(np.float32) if noise else None
    t_rgb, t_disp = rng.rand(N, 3).astype(np.float32), (0.2 + 0.3 * rng.rand(N)).astype(np.float32)
    w_rgb, w_disp = 1.0, 0.1                                # loss = img2mse(rgb, t) + depth_lambda * img2mse(disp, t_disp)
    z, rd, nz = cu(z_np), cu(rd_np), (cu(nz_np) if noise else None)

    raw_a = cu(raw_np).requires_grad_(True)
    rgb, disp, acc, wts, depth, _ = h.raw2outputs(raw_a, z, rd, 1.0 if noise else 0.0, white, _noise=nz)
    loss_a = w_rgb * h.img2mse(rgb, cu(t_rgb)) + w_disp * h.img2mse(disp, cu(t_disp)) + 0.01 * depth.mean()
    loss_a.backward()

    raw_b = cu(raw_np).requires_grad_(True)
    rgb_b, disp_b, acc_b, wts_b, depth_b, _, sq = h.raw2outputs(raw_b, z, rd, 1.0 if noise else 0.0, white, _noise=nz,
                                                                _mse=(cu(t_rgb), cu(t_disp)))
    loss_b = w_rgb * sq[0] / (3 * N) + w_disp * sq[1] / N + 0.01 * depth_b.mean()      # an ordinary upstream gradient rides along
    loss_b.backward()
    assert torch.equal(rgb, rgb_b) and torch.equal(disp, disp_b)
    want = orc.raw2outputs(raw_np, z_np, rd_np, nz_np, white)
    sq_ref = [float(((want["rgb_map"].astype(np.float64) - t_rgb) ** 2).sum()), float(((want["disp_map"].astype(np.float64) - t_disp) ** 2).sum())]
    assert abs(float(sq[0]) - sq_ref[0]) <= 2e-5 * sq_ref[0] and abs(float(sq[1]) - sq_ref[1]) <= 2e-5 * sq_ref[1]
    assert abs(float(loss_a) - float(loss_b)) <= 1e-5 * abs(float(loss_a))
    ga, gb = npy(raw_a.grad), npy(raw_b.grad)
    assert np.abs(ga - gb).max() <= 2e-5 * np.abs(ga).max()
    # bitwise reproducible (fixed-order reduction) and the workspace is left ready for the next call
    sq2 = h.raw2outputs(cu(raw_np), z, rd, 1.0 if noise else 0.0, white, _noise=nz, _mse=(cu(t_rgb), cu(t_disp)))[6]
    assert torch.equal(sq, sq2)
    # one target only
    sq3 = h.raw2outputs(cu(raw_np), z, rd, 1.0 if noise else 0.0, white, _noise=nz, _mse=(cu(t_rgb), None))[6]
    assert float(sq3[0]) == float(sq[0]) and float(sq3[1]) == 0.0
