"""CPU suite: the third-party (PyTorch CPU) rounding behaviour the bit-exact stages depend on — the oracle's
restatements checked against the LIVE torch ops of this interpreter, so a torch upgrade that changes an accumulation order
shows up here, not as an unexplained index mismatch.  Also cross-checks the two independent restatements of the sampling
stage (numpy oracle vs the torch-op port) bit for bit on fuzzed inputs."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as orc
from oracle import torch_cpu_port as port


@pytest.mark.parametrize("K", [1, 2, 3, 4, 8, 9, 15, 16, 31, 32, 33, 62, 63, 64, 65, 126, 127, 128, 200, 255, 256, 512])
def test_aten_row_sum_order(K):
    rng = np.random.RandomState(K)
    x = (rng.rand(3000, K).astype(np.float32) ** 6) * np.float32(10.0) ** rng.randint(-3, 3, size=(3000, 1)).astype(np.float32)
    want = torch.sum(torch.from_numpy(x), -1).numpy()
    assert np.array_equal(orc.aten_sum_lastdim(x), want)


def test_cumsum_cumprod_accumulate_in_fp64():
    rng = np.random.RandomState(1)
    x = rng.rand(5000, 63).astype(np.float32) ** 4
    assert np.array_equal(orc.cumsum_f64_round_f32(x), torch.cumsum(torch.from_numpy(x), -1).numpy())
    q = (1.0 - 0.3 * rng.rand(5000, 128) ** 3).astype(np.float32)
    assert np.array_equal(orc.cumprod_f64_round_f32(q), torch.cumprod(torch.from_numpy(q), -1).numpy())


@pytest.mark.parametrize("steps", [2, 3, 4, 63, 64, 65, 128, 1000])
def test_linspace_matches_torch(steps):
    for a, b in ((0., 1.), (1.2, 7.7369), (-3.5, 0.25)):
        assert np.array_equal(orc.linspace_f32(a, b, steps), torch.linspace(a, b, steps).numpy())


def test_searchsorted_right_is_upper_bound():
    rng = np.random.RandomState(2)
    cdf = np.sort(rng.rand(2000, 63).astype(np.float32), -1)
    cdf[:, 0] = 0
    cdf[::3, 10:14] = cdf[::3, 10:11]                    # plateaus (zero-weight bins)
    u = rng.rand(2000, 64).astype(np.float32)
    u[:, :8] = cdf[:, 5:13]                               # exact ties
    want = torch.searchsorted(torch.from_numpy(cdf), torch.from_numpy(u), right=True).numpy()
    got = np.stack([np.searchsorted(c, v, side="right") for c, v in zip(cdf, u)])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("det", [False, True])
def test_sampling_stage_numpy_oracle_equals_torch_port(det):
    """z_mid / sample_pdf / sort-merge: numpy restatement == the same stage written with torch CPU ops, bit for bit"""
    rng = np.random.RandomState(3)
    N = 4000
    z = np.sort(1.2 + 6.5 * rng.rand(N, 64).astype(np.float32), -1)
    w = (rng.rand(N, 64) ** 8).astype(np.float32)
    w[::7] *= (rng.rand(*w[::7].shape) > 0.9)
    w[::13] = 0
    u = orc.linspace_f32(0, 1, 64) if det else rng.rand(N, 64).astype(np.float32)
    want = orc.fine_samples(z, w, u)
    ut = torch.from_numpy(u).expand(N, 64).contiguous() if det else torch.from_numpy(u)
    got = port.importance(torch.from_numpy(z), torch.from_numpy(w), ut)
    z_merged = got[0] if isinstance(got, (tuple, list)) else got
    assert np.array_equal(z_merged.numpy(), want["z_merged"])


def test_coarse_sampling_numpy_oracle_equals_torch_ops():
    """run.py:1759-1781 written with torch ops on CPU == the oracle, both lindisp modes, with jitter"""
    rng = np.random.RandomState(4)
    N = 3000
    near = torch.from_numpy((0.5 + rng.rand(N, 1)).astype(np.float32))
    far = near + torch.from_numpy((0.1 + 6 * rng.rand(N, 1)).astype(np.float32))
    t_rand = torch.from_numpy(rng.rand(N, 64).astype(np.float32))
    t_vals = torch.linspace(0., 1., steps=64)
    rays = np.zeros((N, 11), np.float32)
    rays[:, 6:7], rays[:, 7:8] = near.numpy(), far.numpy()
    for lindisp in (False, True):
        if not lindisp:
            z = near * (1. - t_vals) + far * (t_vals)
        else:
            z = 1. / (1. / near * (1. - t_vals) + 1. / far * (t_vals))
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        zj = lower + (upper - lower) * t_rand
        assert np.array_equal(orc.sample_coarse(rays, t_vals.numpy(), None, lindisp), z.numpy())
        assert np.array_equal(orc.sample_coarse(rays, t_vals.numpy(), t_rand.numpy(), lindisp), zj.numpy())
