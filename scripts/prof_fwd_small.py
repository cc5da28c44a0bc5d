import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mvip_nerf_b200 import ops
from oracle import nerf_oracle as orc
dev = "cuda"
p = orc.init_params(1)
blob = ops.mlp_pack([torch.from_numpy(p[n]).to(dev) for n in ops.PARAM_ORDER])
P = 1048576
pts = torch.rand(P, 3, device=dev) * 4 - 2
dirs = torch.nn.functional.normalize(torch.randn(P, 3, device=dev), dim=-1)
for _ in range(2): ops.mlp_forward(blob, pts=pts, dirs=dirs)
ops.mlp_forward(blob, pts=pts[:524288].contiguous(), dirs=dirs[:524288].contiguous(), want_stash=True)
torch.cuda.synchronize()
