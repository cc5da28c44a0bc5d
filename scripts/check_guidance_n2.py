"""torchrun --nproc-per-node 2 scripts/check_guidance_n2.py : ShardedGuidanceViews on 2 GPUs (NCCL gather / scatter / allreduce)
against the un-sharded deferred render of the same views computed locally by every rank."""
import os
import sys
import tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from mvip_nerf_b200 import dist as md
from mvip_nerf_b200 import run
from test_gpu_render import load_seeded, nerf_args

rank, world, local = md.init_from_env()
with tempfile.TemporaryDirectory() as td:
    os.makedirs(os.path.join(td, "exp"))
    kw_train, kw_test, _, grad_vars, _ = run.create_nerf(nerf_args(td))
load_seeded(kw_train["network_fn"], 200)
load_seeded(kw_train["network_fine"], 201)
poses = []
for i in range(3):
    p = torch.eye(4, device="cuda")[:3, :4].clone()
    p[0, 3] = 0.05 * i
    poses.append(p)
H, W, focal = 36, 48, 40.0


def loss_of(rgb, depth):
    wgt = torch.linspace(0.5, 1.5, rgb.numel() // 3, device=rgb.device).view(rgb.shape[:-1])
    return ((rgb - 0.4) ** 2 * wgt[..., None]).mean() + 0.02 * (depth * wgt).mean()


outs = [run.render_deferred(H, W, focal, chunk=512, c2w=p, near=1.2, far=7.7, **kw_test) for p in poses]
loss_of(torch.stack([o[0] for o in outs]), torch.stack([o[3] for o in outs])).backward()
want = [None if v.grad is None else v.grad.clone() for v in grad_vars]
for v in grad_vars:
    v.grad = None
g = md.ShardedGuidanceViews(kw_test, poses, H, W, focal, 1.2, 7.7, chunk=512, with_normals=False)
out = g.forward()
if rank == 0:
    assert torch.equal(out["rgb_map"], torch.stack([o[0] for o in outs]).detach())
    loss_of(out["rgb_map"], out["depth_map"]).backward()
g.backward()
worst = 0.0
for v, a in zip(grad_vars, want):
    if a is not None and float(a.abs().max()) > 0:
        worst = max(worst, float((v.grad - a).abs().max()) / float(a.abs().max()))
    else:
        assert float(v.grad.abs().max()) == 0.0
assert worst < 2e-4, worst
print("rank %d: sharded guidance gradients match the local reference, worst rel-to-max %.2e" % (rank, worst))
torch.distributed.barrier()
torch.distributed.destroy_process_group()
