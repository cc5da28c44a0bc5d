"""CPU suite: the numpy oracle replayed against outputs of the reference itself
(tests/golden/*.npz, produced by oracle/make_golden.py from /root/reference)."""
import numpy as np
import pytest

from oracle import nerf_oracle as orc


def test_linspace_tables(golden):
    fx = golden("sample_coarse")
    assert np.array_equal(orc.linspace_f32(0, 1, 64), fx["t_vals"])
    assert np.array_equal(orc.linspace_f32(0, 1, 64), golden("sample_pdf")["u_det"])


@pytest.mark.parametrize("lindisp", [0, 1])
@pytest.mark.parametrize("perturb", [0, 1])
def test_sample_coarse_bit_exact(golden, lindisp, perturb):
    fx = golden("sample_coarse")
    key = "lindisp%d_perturb%d" % (lindisp, perturb)
    z = orc.sample_coarse(fx[key + "_rays"], fx["t_vals"], fx["t_rand"] if perturb else None, bool(lindisp))
    assert np.array_equal(z, fx[key + "_z"])


@pytest.mark.parametrize("dist", ["uniform", "peaky", "sparse", "zeros"])
@pytest.mark.parametrize("mode", ["det", "rand"])
def test_sample_pdf_bit_exact(golden, dist, mode):
    fx = golden("sample_pdf")
    u = fx["u_det"] if mode == "det" else fx["u_rand"]
    out = orc.sample_pdf(fx["bins"], fx["w_" + dist], u)
    tag = "%s_%s" % (dist, mode)
    assert np.array_equal(out["cdf"], fx["cdf_" + tag])
    assert np.array_equal(out["inds"], fx["inds_" + tag])
    assert out["inds"].dtype == np.int64
    assert np.array_equal(out["samples"], fx["samples_" + tag])


def test_searchsorted_ties():
    # torch.searchsorted(right=True) == numpy side='right' == upper_bound
    cdf = np.array([0, .25, .25, .5, 1], dtype=np.float32)
    u = np.array([0, .25, .3, .5, .999, 1, 1.5], dtype=np.float32)
    assert list(np.searchsorted(cdf, u, side="right")) == [1, 3, 3, 4, 4, 5, 5]


def test_merge(golden):
    fx = golden("merge")
    w = np.zeros((fx["z"].shape[0], 64), np.float32)
    merged = np.sort(np.concatenate([fx["z"], fx["z_samples"]], -1), -1)
    assert np.array_equal(merged, fx["merged"])
    zstd = np.std(fx["z_samples"].astype(np.float64), -1)
    np.testing.assert_allclose(zstd, fx["z_std"], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("S", [64, 128])
@pytest.mark.parametrize("white", [0, 1])
@pytest.mark.parametrize("use_noise", [0, 1])
def test_raw2outputs_forward(golden, S, white, use_noise):
    fx = golden("raw2outputs")
    tag = "S%d_w%d_n%d" % (S, white, use_noise)
    noise = fx["S%d_noise" % S] if use_noise else None
    out = orc.raw2outputs(fx["S%d_raw" % S], fx["S%d_z" % S], fx["S%d_rays_d" % S], noise, bool(white))
    for k, name in [("rgb_map", "rgb"), ("acc_map", "acc"), ("weights", "weights"), ("depth_map", "depth"),
                    ("alpha", "alpha"), ("disp_map", "disp")]:
        ref = fx[tag + "_" + name]
        got = out[k]
        assert np.array_equal(np.isnan(ref), np.isnan(got)), k
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6, equal_nan=True, err_msg=k)
    if not use_noise:
        assert np.isnan(out["disp_map"][0])      # the all-sigma<=0 ray keeps its NaN (reference behaviour)


@pytest.mark.parametrize("S", [64, 128])
@pytest.mark.parametrize("white", [0, 1])
@pytest.mark.parametrize("use_noise", [0, 1])
def test_raw2outputs_backward(golden, S, white, use_noise):
    fx = golden("raw2outputs")
    tag = "S%d_w%d_n%d" % (S, white, use_noise)
    noise = fx["S%d_noise" % S] if use_noise else None
    d_raw = orc.raw2outputs_backward(fx["S%d_raw" % S], fx["S%d_z" % S], fx["S%d_rays_d" % S], noise, bool(white),
                                     fx[tag + "_g_rgb"], fx[tag + "_g_disp"], fx[tag + "_g_acc"],
                                     fx[tag + "_g_depth"], fx[tag + "_g_weights"])
    ref = fx[tag + "_d_raw"]
    assert np.array_equal(np.isnan(ref), np.isnan(d_raw))     # the disp=NaN ray poisons its own grads, as in the reference
    ok = ~np.isnan(ref)
    scale = np.abs(ref[ok]).max()
    assert np.abs(d_raw[ok] - ref[ok]).max() <= 2e-5 * scale


def test_embed_and_mlp(golden):
    fx = golden("nerf_mlp")
    p = orc.init_params(int(fx["param_seed"]))
    x = np.concatenate([orc.embed(fx["pts"], 10), orc.embed(fx["viewdirs"], 4)], -1)
    assert x.shape[1] == 90
    np.testing.assert_allclose(x, fx["embedded"], atol=2e-6, rtol=0)
    out, saved = orc.nerf_forward(p, fx["embedded"], keep=True)
    np.testing.assert_allclose(out, fx["out"], atol=2e-5, rtol=1e-5)
    g = orc.nerf_backward(p, saved, fx["d_out"])
    n = 0
    for k in g:
        ref = fx["grad." + k]
        assert np.abs(g[k] - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), k
        n += ref.size
    assert n == 595844       # params per network


def test_normal_map(golden):
    fx = golden("normal_map")
    K = fx["K"]
    n = orc.normal_from_depth(fx["depth"], K[0, 0], K[1, 1], K[0, 2], K[1, 2], k=31)
    np.testing.assert_allclose(n, fx["normal_f64"], atol=1e-9)
    np.testing.assert_allclose(n, fx["normal_f32"], atol=2e-4)          # reference fp32 unfold+inv path
    gd = orc.normal_from_depth_backward(fx["depth"], K[0, 0], K[1, 1], K[0, 2], K[1, 2], fx["g_normal_f64"], k=31)
    np.testing.assert_allclose(gd, fx["d_depth_f64"], atol=1e-9 * max(1.0, np.abs(fx["d_depth_f64"]).max()))


def test_render_rays_end_to_end(golden):
    fx = golden("render_e2e")
    pc = orc.init_params(int(fx["coarse_seed"]))
    pf = orc.init_params(int(fx["fine_seed"]))
    rays = orc.make_ray_batch(fx["rays_o"], fx["rays_d"], fx["near"], fx["far"])
    t_vals = orc.linspace_f32(0, 1, 64)
    # render kwargs (perturb=0, noise=0)
    out = orc.render_rays(rays, pc, pf, t_vals, lindisp=True, white_bkgd=True)
    np.testing.assert_allclose(out["rgb0"], fx["test_rgb0"], atol=2e-5)
    np.testing.assert_allclose(out["rgb_map"], fx["test_rgb"], atol=5e-5)
    np.testing.assert_allclose(out["depth_map"], fx["test_depth"], rtol=2e-4)
    np.testing.assert_allclose(out["acc_map"], fx["test_acc"], atol=2e-5)
    # fine z_vals follow the coarse weights: fp32 MLP rounding differences (BLAS order) may move a sample
    # across a bin edge, so compare the bulk
    close = np.isclose(out["z_vals"], fx["test_z_vals"], atol=1e-3)
    assert close.mean() > 0.995
    # train kwargs with the pytest=True streams
    out = orc.render_rays(rays, pc, pf, t_vals, t_rand=fx["train_t_rand"], u=fx["train_u"],
                          noise0=fx["train_noise0"], noise1=fx["train_noise1"], lindisp=True, white_bkgd=True)
    np.testing.assert_allclose(out["rgb0"], fx["train_rgb0"], atol=2e-5)
    np.testing.assert_allclose(out["rgb_map"], fx["train_rgb"], atol=5e-5)
    np.testing.assert_allclose(out["disp_map"], fx["train_disp"], rtol=2e-4)


def test_torch_cpu_port_matches_reference_golden(golden):
    # the bench's CPU baseline arm (oracle/torch_cpu_port.py) against the reference's own render() outputs
    import torch
    from oracle import torch_cpu_port as port
    fx = golden("render_e2e")
    pc = port.params_from_numpy(orc.init_params(int(fx["coarse_seed"])), True)
    pf = port.params_from_numpy(orc.init_params(int(fx["fine_seed"])), True)
    rays = torch.from_numpy(orc.make_ray_batch(fx["rays_o"], fx["rays_d"], fx["near"], fx["far"]))
    tv = torch.linspace(0., 1., 64)
    with torch.no_grad():
        out = port.render_rays(rays, pc, pf, tv)
    np.testing.assert_allclose(out["rgb_map"].numpy(), fx["test_rgb"], atol=1e-5)
    np.testing.assert_allclose(out["z_vals"].numpy(), fx["test_z_vals"], atol=1e-5)
    out = port.render_rays(rays, pc, pf, tv, torch.from_numpy(fx["train_t_rand"]), torch.from_numpy(fx["train_u"]),
                           torch.from_numpy(fx["train_noise0"]), torch.from_numpy(fx["train_noise1"]))
    np.testing.assert_allclose(out["rgb_map"].detach().numpy(), fx["train_rgb"], atol=1e-5)
    loss = ((out["rgb_map"] - 0.5) ** 2).mean() + ((out["rgb0"] - 0.5) ** 2).mean() + 0.1 * ((out["disp_map"] - 0.3) ** 2).mean()
    loss.backward()
    assert abs(loss.item() - float(fx["train_loss"])) < 1e-5
    g = pf["pts_linears.0.weight"].grad.numpy()
    ref = fx["train_grad.fine.pts_linears.0.weight"]
    assert np.abs(g - ref).max() <= 1e-3 * np.abs(ref).max()


def test_oracle_train_step_matches_torch_port():
    import torch
    from oracle import torch_cpu_port as port
    rng = np.random.RandomState(0)
    N = 24
    ro, rd = orc.get_rays(756, 1008, 767.2935, np.eye(4, dtype=np.float32)[:3, :4])
    idx = rng.permutation(756 * 1008)[:N]
    rays = orc.make_ray_batch(ro.reshape(-1, 3)[idx], rd.reshape(-1, 3)[idx], 1.2, 7.7369)
    pcn, pfn = orc.init_params(1), orc.init_params(2)
    tv = orc.linspace_f32(0, 1, 64)
    r = [rng.rand(N, 64).astype(np.float32), rng.rand(N, 64).astype(np.float32),
         rng.randn(N, 64).astype(np.float32), rng.randn(N, 128).astype(np.float32)]
    target = rng.rand(N, 3).astype(np.float32)
    loss, g0, g1 = orc.train_step(rays, pcn, pfn, tv, r[0], r[1], r[2], r[3], target, dtype=np.float64)
    pc, pf = port.params_from_numpy(pcn, True), port.params_from_numpy(pfn, True)
    lt = port.train_step(torch.from_numpy(rays), pc, pf, torch.from_numpy(tv), *[torch.from_numpy(x) for x in r],
                         torch.from_numpy(target))
    assert abs(loss - lt.item()) < 1e-5
    for name in ("pts_linears.0.weight", "pts_linears.5.weight", "rgb_linear.weight", "alpha_linear.bias"):
        for gn, pt in ((g0, pc), (g1, pf)):
            ref = pt[name].grad.numpy()
            assert np.abs(gn[name] - ref).max() <= 2e-3 * max(np.abs(ref).max(), 1e-12), name


RAY_CASES = ["plain", "ndc", "ndc_static", "ndc_patch", "ndc_nodirs", "static", "patch"]


def ray_case(g, name):
    H, W, focal, near, far, ndc, vd, i0, j0, h, w = g[name + "_args"]
    return dict(H=int(H), W=int(W), focal=float(focal), near=float(near), far=float(far), ndc=bool(ndc), use_viewdirs=bool(vd),
                patch=(int(i0), int(j0), int(h), int(w)), c2w=g[name + "_c2w"], c2w_static=g.get(name + "_c2w_static"))


@pytest.mark.parametrize("name", RAY_CASES)
def test_ray_batch_bit_exact(golden, name):
    """get_rays + ndc_rays + the batch assembly of render() (run.py:1171-1207) vs the batch the reference's render() built."""
    g = golden("ray_batch")
    c = ray_case(g, name)
    ro, rd = orc.get_rays(c["H"], c["W"], c["focal"], c["c2w"])
    view = rd
    if c["c2w_static"] is not None:
        ro, rd = orc.get_rays(c["H"], c["W"], c["focal"], c["c2w_static"])
    i0, j0, h, w = c["patch"]
    sl = (slice(i0, i0 + h), slice(j0, j0 + w))
    got = orc.make_ray_batch(ro[sl], rd[sl], c["near"], c["far"], c["use_viewdirs"], view[sl], c["ndc"], c["H"], c["W"], c["focal"])
    assert np.array_equal(got, g[name + "_batch"])
    if name + "_rays_batch" in g:
        got = orc.make_ray_batch(g[name + "_rays_o"], g[name + "_rays_d"], c["near"], c["far"], c["use_viewdirs"], None, c["ndc"],
                                 c["H"], c["W"], c["focal"])
        assert np.array_equal(got, g[name + "_rays_batch"])


def test_cfg1_real_view_oracle(golden):
    """BASELINE cfg 1: a training view of data/1 (pose pipeline of load_llff.py replayed by the reference's own functions,
    oracle/make_golden_cfg1.py), render kwargs: oracle vs the reference's render() on a 16 x 24 patch and on strided rays."""
    fx = golden("cfg1_view")
    H, W, focal, near, far = int(fx["H"]), int(fx["W"]), float(fx["focal"]), float(fx["near"]), float(fx["far"])
    pc, pf = orc.init_params(int(fx["coarse_seed"])), orc.init_params(int(fx["fine_seed"]))
    t_vals = orc.linspace_f32(0, 1, 64)
    ro, rd = orc.get_rays(H, W, focal, fx["c2w"])
    assert np.array_equal(rd[::40, ::41].reshape(-1, 3), fx["rays_d"])            # get_rays on the real pose: bit-exact
    i0, j0, h, w = [int(v) for v in fx["patch"]]
    rays = orc.make_ray_batch(ro[i0:i0 + h, j0:j0 + w], rd[i0:i0 + h, j0:j0 + w], near, far)
    out = orc.render_rays(rays, pc, pf, t_vals, lindisp=True, white_bkgd=True)
    np.testing.assert_allclose(out["rgb0"].reshape(h, w, 3), fx["patch_rgb0"], atol=2e-5)
    np.testing.assert_allclose(out["rgb_map"].reshape(h, w, 3), fx["patch_rgb"], atol=5e-5)
    np.testing.assert_allclose(out["depth_map"].reshape(h, w), fx["patch_depth"], rtol=2e-4)
    np.testing.assert_allclose(out["acc_map"].reshape(h, w), fx["patch_acc"], atol=2e-5)


def test_ndc_configuration_oracle(golden):
    """forward-facing configuration (ndc_rays, near 0 / far 1, linear sampling, black background): oracle vs the reference's render()"""
    fx = golden("render_ndc")
    H, W, focal = int(fx["H"]), int(fx["W"]), float(fx["focal"])
    pc, pf = orc.init_params(int(fx["coarse_seed"])), orc.init_params(int(fx["fine_seed"]))
    ro, rd = orc.get_rays(H, W, focal, fx["c2w"])
    rays = orc.make_ray_batch(ro, rd, 0., 1., True, None, True, H, W, focal)
    out = orc.render_rays(rays, pc, pf, orc.linspace_f32(0, 1, 64), lindisp=False, white_bkgd=False)
    ok = (np.abs(fx["sigma_far_fine"]) > 1e-4).reshape(-1)
    ok0 = (np.abs(fx["sigma_far_coarse"]) > 1e-4).reshape(-1)
    np.testing.assert_allclose(out["rgb_map"][ok], fx["rgb"].reshape(-1, 3)[ok], atol=5e-5)
    np.testing.assert_allclose(out["acc_map"][ok], fx["acc"].reshape(-1)[ok], atol=2e-5)
    np.testing.assert_allclose(out["rgb0"][ok0], fx["rgb0"].reshape(-1, 3)[ok0], atol=5e-5)
    np.testing.assert_allclose(out["depth_map"][ok], fx["depth"].reshape(-1)[ok], rtol=2e-4, atol=1e-5)
