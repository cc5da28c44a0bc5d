"""Launches every HBM-bound stage kernel twice at an image-sized batch (262,144 rays) and the render-only MLP forward twice
(4.2 M points) in a fixed order: the target of the `ncu --set full` pass of scripts/profile.sh."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvip_nerf_b200 import ops
from oracle import nerf_oracle as orc

dev = "cuda"
N = 262144
g = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *sh: torch.rand(*sh, device=dev, generator=g)  # noqa: E731
for rep in range(2):
    for S in (64, 128):
        raw = rnd(N, S, 4) * 2 - 1
        z = torch.sort(rnd(N, S) * 6 + 1.2, -1)[0]
        rd = rnd(N, 3) - .5
        ops.composite_forward(raw, z, rd, None, True)
        ops.composite_backward(raw, z, rd, None, True, False, rnd(N, 3), rnd(N), rnd(N), rnd(N))
    z = torch.sort(rnd(N, 64) * 6 + 1.2, -1)[0]
    w = rnd(N, 64) ** 4
    ops.sample_fine(z, w, rnd(N, 64), want_samples=False)
    ops.sample_fine(z, w, torch.linspace(0, 1, 64, device=dev), want_samples=False)
    rays = rnd(N, 11) + 1
    ops.sample_coarse(rays, torch.linspace(0, 1, 64, device=dev), rnd(N, 64), True)
    ops.rays_from_pose(756, 1008, 767.2935, torch.eye(4, device=dev)[:3, :4].contiguous(), 1.2, 7.7)
    p = orc.init_params(1)
    blob = ops.mlp_pack([torch.from_numpy(p[n]).to(dev) for n in ops.PARAM_ORDER])
    P = 32768 * 128
    pts = rnd(P, 3) * 4 - 2
    dirs = torch.nn.functional.normalize(rnd(P, 3) - .5, dim=-1)
    ops.mlp_forward(blob, pts=pts, dirs=dirs)
torch.cuda.synchronize()
print("done")
