import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mvip_nerf_b200 import ops, _lib
from oracle import nerf_oracle as orc
if os.environ.get("MVIP_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MVIP_LIB"])
dev = "cuda"
p = orc.init_params(1)
blob = ops.mlp_pack([torch.from_numpy(p[n]).to(dev) for n in ops.PARAM_ORDER])
P = 4194304
pts = torch.rand(P, 3, device=dev) * 4 - 2
dirs = torch.nn.functional.normalize(torch.randn(P, 3, device=dev), dim=-1)
for stash in (False, True):
    if stash: P2 = 524288; pts, dirs = pts[:P2].contiguous(), dirs[:P2].contiguous()
    for _ in range(3): ops.mlp_forward(blob, pts=pts, dirs=dirs, want_stash=stash)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.mlp_forward(blob, pts=pts, dirs=dirs, want_stash=stash); b.record(); torch.cuda.synchronize()
    out = (ctypes.c_ulonglong * 16)()
    _lib.load().mvip_debug_profile(out)
    v = list(out)
    n = pts.shape[0]
    print("stash=%s P=%d  %.3f ms  %.1f TF" % (stash, n, a.elapsed_time(b), n * 1186816 / a.elapsed_time(b) / 1e9))
    print("  issuer X (needs a -DMVIP_PROF build): act-wait %.1f%%  weight-wait %.1f%%  total %d cyc" % (100 * v[0] / max(v[2], 1), 100 * v[1] / max(v[2], 1), v[2]))
    print("  epilogue (slot 0, warp 0): acc-wait %.1f%% of %d cyc" % (100 * v[3] / max(v[5], 1), v[5]))
