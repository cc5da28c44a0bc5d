"""CPU suite: the plain-C oracle (oracle/nerf_oracle_c.c — no numpy, no torch in the arithmetic) against the golden vectors the
reference produced, bit for bit on the bit-exact stages, and against the numpy oracle on fuzzed inputs."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import nerf_oracle as orc
from test_oracle_golden import RAY_CASES, ray_case


@pytest.mark.parametrize("lindisp", [0, 1])
@pytest.mark.parametrize("perturb", [0, 1])
def test_c_sample_coarse_bit_exact(golden, lindisp, perturb):
    fx = golden("sample_coarse")
    key = "lindisp%d_perturb%d" % (lindisp, perturb)
    z = co.sample_coarse(fx[key + "_rays"], fx["t_vals"], fx["t_rand"] if perturb else None, bool(lindisp))
    assert np.array_equal(z, fx[key + "_z"])


@pytest.mark.parametrize("dist", ["uniform", "peaky", "sparse", "zeros"])
@pytest.mark.parametrize("mode", ["det", "rand"])
def test_c_sample_fine_bit_exact(golden, dist, mode):
    fx, mg = golden("sample_pdf"), golden("merge")
    z = mg["z"]
    assert np.array_equal((np.float32(0.5) * (z[:, 1:] + z[:, :-1])).astype(np.float32), fx["bins"])     # same rays in both fixtures
    w = np.zeros((z.shape[0], 64), np.float32)
    w[:, 1:-1] = fx["w_" + dist]
    u = fx["u_det"] if mode == "det" else fx["u_rand"]
    out = co.fine_samples(z, w, u)
    tag = "%s_%s" % (dist, mode)
    assert np.array_equal(out["cdf"], fx["cdf_" + tag])
    assert np.array_equal(out["inds"], fx["inds_" + tag])
    assert np.array_equal(out["z_samples"], fx["samples_" + tag])
    if tag == "peaky_rand":
        assert np.array_equal(out["z_merged"], mg["merged"])          # == torch.sort(cat) of the reference


@pytest.mark.parametrize("name", RAY_CASES)
def test_c_ray_batch_bit_exact(golden, name):
    g = golden("ray_batch")
    c = ray_case(g, name)
    whole = c["patch"] == (0, 0, c["H"], c["W"])
    got = co.ray_batch(c["H"], c["W"], c["focal"], c["c2w"], c["near"], c["far"], c["use_viewdirs"], c["c2w_static"],
                       None if whole else c["patch"], c["ndc"])
    assert np.array_equal(got, g[name + "_batch"])


@pytest.mark.parametrize("S", [64, 128])
@pytest.mark.parametrize("white", [0, 1])
@pytest.mark.parametrize("use_noise", [0, 1])
def test_c_raw2outputs_vs_reference(golden, S, white, use_noise):
    fx = golden("raw2outputs")
    tag = "S%d_w%d_n%d" % (S, white, use_noise)
    out = co.raw2outputs(fx["S%d_raw" % S], fx["S%d_z" % S], fx["S%d_rays_d" % S], fx["S%d_noise" % S] if use_noise else None, bool(white))
    for name, key in (("rgb", "rgb_map"), ("acc", "acc_map"), ("weights", "weights"), ("depth", "depth_map"), ("alpha", "alpha"),
                      ("disp", "disp_map")):
        ref = fx[tag + "_" + name]
        assert np.array_equal(np.isnan(ref), np.isnan(out[key])), name
        np.testing.assert_allclose(out[key], ref, rtol=1e-5, atol=1e-6, equal_nan=True, err_msg=name)


def test_c_oracle_equals_numpy_oracle_on_fuzz():
    rng = np.random.RandomState(9)
    N = 3000
    rays = rng.randn(N, 11).astype(np.float32)
    rays[:, 6] = 0.5 + rng.rand(N)
    rays[:, 7] = rays[:, 6] + 0.1 + 6 * rng.rand(N)
    for S in (32, 48, 64, 128):
        t = orc.linspace_f32(0, 1, S)
        tr = rng.rand(N, S).astype(np.float32)
        for lindisp in (False, True):
            assert np.array_equal(co.sample_coarse(rays, t, tr, lindisp), orc.sample_coarse(rays, t, tr, lindisp))
    z = np.sort(1.2 + 6.5 * rng.rand(N, 64).astype(np.float32), -1)
    w = (rng.rand(N, 64) ** 8).astype(np.float32)
    w[::7] *= (rng.rand(*w[::7].shape) > 0.9)
    w[5::17] *= 1e-12
    u = rng.rand(N, 64).astype(np.float32)
    u[:, 0], u[:, 1] = 0.0, 1.0
    a, b = co.fine_samples(z, w, u), orc.fine_samples(z, w, u)
    for k in ("inds", "z_samples", "z_merged"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["cdf"], b["cdf"])
    # other bin counts (generic kernels' shapes)
    z2 = np.sort(1.2 + 6.5 * rng.rand(500, 48).astype(np.float32), -1)
    w2 = rng.rand(500, 48).astype(np.float32)
    u2 = rng.rand(500, 32).astype(np.float32)
    a, b = co.fine_samples(z2, w2, u2), orc.fine_samples(z2, w2, u2)
    assert np.array_equal(a["inds"], b["inds"]) and np.array_equal(a["z_merged"], b["z_merged"])
