"""Hand-over timeline of one fused backward: publication / pick-up stamps of every (tile, dZ unit).  python scripts/bwd_timeline.py [P]"""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvip_nerf_b200 import _lib, ops  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402
if os.environ.get("MVIP_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MVIP_LIB"])
dev = "cuda"
P = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
p = orc.init_params(1)
blob = ops.mlp_pack([torch.from_numpy(p[n]).to(dev) for n in ops.PARAM_ORDER])
g = torch.Generator(device=dev).manual_seed(3)
pts = torch.rand(P, 3, device=dev, generator=g) * 4 - 2
dirs = torch.nn.functional.normalize(torch.randn(P, 3, device=dev, generator=g), dim=-1)
raw, stash = ops.mlp_forward(blob, pts=pts, dirs=dirs, want_stash=True)
d_raw = torch.randn(P, 4, device=dev, generator=g)
lib = _lib.load()
if os.environ.get("THROTTLE"):
    lib.mvip_debug_set_bwd_throttle(*[int(x) for x in os.environ["THROTTLE"].split(",")])
if os.environ.get("STAGGER"):
    lib.mvip_debug_set_bwd_stagger(int(os.environ["STAGGER"]))
ws = ops._aligned_bytes(lib.mvip_mlp_backward_workspace_bytes(P), dev)
flat = torch.zeros(595844, device=dev)
grads, off = [], 0
for shp in ops.PARAM_SHAPES:
    n = int(torch.Size(shp).numel())
    grads.append(flat[off:off + n].view(shp)); off += n
arr = (ctypes.c_void_p * 24)(*[x.data_ptr() for x in grads])
for _ in range(3):
    _lib.check(lib.mvip_mlp_backward_phases(ops._ptr(blob), ops._ptr(d_raw), P, ops._ptr(stash), ops._ptr(ws), arr, 0, 1, ops._stream()), "bwd")
torch.cuda.synchronize()
a, b = ctypes.c_size_t(), ctypes.c_size_t()
lib.mvip_debug_bwd_stamp_offsets(P, ctypes.byref(a), ctypes.byref(b))
T = (P + 127) // 128
raw_ws = ws.view(torch.uint8)
pub = raw_ws[a.value:a.value + T * 40].view(torch.int32).cpu().numpy().astype(np.int64).reshape(T, 10) & 0xffffffff
pick = raw_ws[b.value:b.value + T * 40].view(torch.int32).cpu().numpy().astype(np.int64).reshape(T, 10) & 0xffffffff
t0 = pub.min()
pub = (pub - t0) / 1e3
pick = (pick - t0) / 1e3          # us
print("P=%d tiles=%d  kernel span %.0f us" % (P, T, max(pub.max(), pick.max())))
print("unit:                 " + "  ".join("%6d" % u for u in range(10)))
for name, x in (("pick-up - publication (us) mean", (pick - pub).mean(0)), ("                           p90 ", np.percentile(pick - pub, 90, axis=0)),
                ("                           max ", (pick - pub).max(0))):
    print(name + "  " + "  ".join("%6.1f" % v for v in x))
nc = 74
for cl in (0, 1, 36, 73):
    tiles = [2 * (cl + nc * w) for w in range(0, T // (2 * nc) + 1) if 2 * (cl + nc * w) < T]
    print("chain of cluster %d: publication of unit 9 per wave (us): %s" % (cl, " ".join("%.0f" % pub[t, 9] for t in tiles[:12])))
    print("     unit times of wave 2 (us): %s" % " ".join("%.1f" % (pub[tiles[2], u] - pub[tiles[2], 0]) for u in range(10)))
w2 = [2 * (c + nc * 2) for c in range(nc)]
print("wave 2, publication of unit 5 by cluster (us): " + " ".join("%.0f" % pub[t, 5] for t in w2))
print("wave 2, pick-up of unit 5 by cluster (us):     " + " ".join("%.0f" % pick[t, 5] for t in w2))
np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "bwd_timeline.npz"), pub=pub, pick=pick)
