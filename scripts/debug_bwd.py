import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mvip_nerf_b200 import ops
from oracle import nerf_oracle as orc
def cu(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
for P in (256, 20000):
    rng = np.random.RandomState(100 + P)
    p = orc.init_params(21)
    pts = ((rng.rand(P, 3) * 2 - 1) * 5).astype(np.float32)
    vd = rng.randn(P, 3).astype(np.float32); vd /= np.linalg.norm(vd, axis=-1, keepdims=True)
    d_out = rng.randn(P, 4).astype(np.float32)
    x = np.concatenate([orc.embed(pts, 10), orc.embed(vd, 4)], -1)
    _, saved = orc.nerf_forward(p, x, keep=True, dtype=np.float64)
    want = orc.nerf_backward(p, saved, d_out, dtype=np.float64)
    blob = ops.mlp_pack([cu(p[n]) for n in ops.PARAM_ORDER])
    raw, stash = ops.mlp_forward(blob, pts=cu(pts), dirs=cu(vd), want_stash=True)
    grads = ops.mlp_backward(blob, cu(d_out), stash)
    print("P =", P)
    for g, name in zip(grads, ops.PARAM_ORDER):
        ref = want[name]; got = g.cpu().numpy()
        err = np.abs(got - ref).max() / np.abs(ref).max()
        cos = (got * ref).sum() / np.sqrt((got**2).sum() * (ref**2).sum())
        extra = ""
        if name == "pts_linears.5.weight":
            extra = " pe-part err %.3g h-part err %.3g" % (np.abs(got[:, :63] - ref[:, :63]).max() / np.abs(ref).max(), np.abs(got[:, 63:] - ref[:, 63:]).max() / np.abs(ref).max())
        if name == "pts_linears.0.weight":
            e = np.abs(got - ref).max(0) / np.abs(ref).max()
            extra = " per-col err: " + " ".join("%.2f" % v for v in e[:63:3])
        print("  %-24s relerr %.4f cos %.5f%s" % (name, err, cos, extra))
