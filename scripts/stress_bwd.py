"""Stress of the fused backward's inter-CTA protocol: many point counts (tiny, ragged, around the wave size of 148 tiles, large),
poisoned workspace, two launches each, gradients must be finite and bit-identical.  python scripts/stress_bwd.py [iterations]"""
import ctypes, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvip_nerf_b200 import _lib, ops  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402

dev = "cuda"
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.RandomState(0)
p = orc.init_params(3)
blob = ops.mlp_pack([torch.from_numpy(p[n]).to(dev) for n in ops.PARAM_ORDER])
lib = _lib.load()
sizes = [1, 127, 128, 129, 255, 256, 257, 128 * 73, 128 * 74 + 1, 128 * 147, 128 * 148, 128 * 148 + 1, 128 * 149, 128 * 296 + 5, 128 * 1000 + 77]
sizes += [int(rng.randint(1, 128 * 600)) for _ in range(max(0, iters - len(sizes) - 3))]
sizes += [786432, 1572864 + 11, 2 ** 21 + 3]
t_start = time.time()
for it, P in enumerate(sizes):
    g = torch.Generator(device=dev).manual_seed(it)
    pts = torch.rand(P, 3, device=dev, generator=g) * 4 - 2
    dirs = torch.nn.functional.normalize(torch.randn(P, 3, device=dev, generator=g), dim=-1)
    raw, stash = ops.mlp_forward(blob, pts=pts, dirs=dirs, want_stash=True)
    d = torch.randn(P, 4, device=dev, generator=g)
    ws = ops._aligned_bytes(lib.mvip_mlp_backward_workspace_bytes(P), dev)
    outs = []
    for rep in range(2):
        ws.fill_(0xFF)
        flat = torch.full((595844,), float("nan"), device=dev)
        grads, off = [], 0
        for shp in ops.PARAM_SHAPES:
            n = int(torch.Size(shp).numel())
            grads.append(flat[off:off + n].view(shp)); off += n
        arr = (ctypes.c_void_p * 24)(*[x.data_ptr() for x in grads])
        _lib.check(lib.mvip_mlp_backward(ops._ptr(blob), ops._ptr(d), P, ops._ptr(stash), ops._ptr(ws), arr, 0, ops._stream()), "bwd")
        torch.cuda.synchronize()
        outs.append(flat.clone())
    ok = bool(torch.isfinite(outs[0]).all()) and torch.equal(outs[0], outs[1])
    print("P=%8d tiles=%6d  finite+deterministic=%s  |g|=%.4e" % (P, (P + 127) // 128, ok, float(outs[0].norm())), flush=True)
    assert ok, P
    del stash, ws, raw
print("stress_bwd OK: %d sizes in %.1f s" % (len(sizes), time.time() - t_start))
