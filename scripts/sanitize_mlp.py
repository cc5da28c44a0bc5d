"""Small forward(+stash) / backward of the tcgen05 MLP kernels (+ head grads, fine sampler) for compute-sanitizer (racecheck / synccheck / memcheck):

    compute-sanitizer --tool racecheck python scripts/sanitize_mlp.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvip_nerf_b200 import ops  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402  (seeded weights only)

dev = "cuda"
p = orc.init_params(1)
blob = ops.mlp_pack([torch.from_numpy(p[n]).to(dev) for n in ops.PARAM_ORDER])
g = torch.Generator(device=dev).manual_seed(3)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 9 + 17
pts = torch.rand(P, 3, device=dev, generator=g) * 4 - 2
dirs = torch.nn.functional.normalize(torch.randn(P, 3, device=dev, generator=g), dim=-1)
raw = ops.mlp_forward(blob, pts=pts, dirs=dirs)
raw2, stash = ops.mlp_forward(blob, pts=pts, dirs=dirs, want_stash=True)
grads = ops.mlp_backward(blob, torch.randn(P, 4, device=dev, generator=g), stash)
torch.cuda.synchronize()
print("sanitize_mlp: P=%d raw diff %.3e grad0 norm %.4f" % (P, float((raw - raw2).abs().max()), float(grads[0].norm())))
# the four-rays-per-warp fine sampler (ragged row count: the last warp holds 1 - 3 rays) and the coarse sampler
N = 4 * 37 + 3
z = torch.sort(1.2 + 6.5 * torch.rand(N, 64, device=dev, generator=g), -1).values
w = torch.rand(N, 64, device=dev, generator=g) ** 8
w[::5] *= 1e-12                                   # sequential-scan fallback rows
u = torch.rand(N, 64, device=dev, generator=g)
fs = ops.sample_fine(z, w, u, want_inds=True)
torch.cuda.synchronize()
print("sanitize_mlp: sample_fine N=%d merged sorted %s" % (N, bool((fs["z_merged"][:, 1:] >= fs["z_merged"][:, :-1]).all())))
