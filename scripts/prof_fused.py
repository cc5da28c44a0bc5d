"""Timing + role cycle counters (cluster 0) of backward_fused_kernel alone: python scripts/prof_fused.py [P]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvip_nerf_b200 import _lib, ops  # noqa: E402
from oracle import nerf_oracle as orc  # noqa: E402

if os.environ.get("MVIP_LIB"):          # experiment builds (scripts/build_variant.sh)
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


def run(mask):
    rc = lib.mvip_mlp_backward_phases(ops._ptr(blob), ops._ptr(d_raw), P, ops._ptr(stash), ops._ptr(ws), arr, 0, mask, ops._stream())
    _lib.check(rc, "phases")


if os.environ.get("TRACE"):
    run(1); run(1)
    tr = (ctypes.c_longlong * 240)()
    if lib.mvip_debug_bwd_trace(tr) == 0:
        import numpy as np
        a = np.array(list(tr), dtype=np.int64).reshape(20, 12)
        names = ["acc-wait>", "wake", "loaded", "arrived", "math done", "st hi done", "stage_out done", "so: buffer free", "so: STS done", "so: fenced", "so: issued"]
        t0 = a[0, 0]
        for hs in range(19):
            r = a[hs]
            print("hs %2d  start %7d | wait %5d  ld %5d  arrive %5d  math %5d  sthi %5d  stage_out %5d [free %5d sts %5d fence %5d issue %5d rest %5d] | period %6d"
                  % (hs, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[7] - r[5], r[8] - r[7], r[9] - r[8],
                     r[10] - r[9], r[6] - r[10], (a[hs + 1, 0] - r[0]) if hs < 18 else 0))
    sys.exit(0)

for mask, name in ((1, "backward_fused_kernel"), (4, "head_grads"), (8, "reduce"), (15, "all")):
    for _ in range(2):
        run(mask)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(mask); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("%-28s P=%d  best %.3f ms  median %.3f ms" % (name, P, min(ts), sorted(ts)[2]))
    if mask == 1:
        out = (ctypes.c_ulonglong * 8)()
        lib.mvip_debug_wgrad_profile(out)
        v = list(out)
        print("  cluster 0: chain issuer total %d  wait act/hi %d (%.0f%%)  wait weights %d (%.0f%%)" % (v[0], v[1], 100. * v[1] / max(v[0], 1), v[2], 100. * v[2] / max(v[0], 1)))
        print("             wgrad issuer total %d  wait full %d (%.0f%%)" % (v[3], v[4], 100. * v[4] / max(v[3], 1)))
        lag = (ctypes.c_ulonglong * 320)()
        lib.mvip_debug_bwd_lag(lag)
        L = [list(lag)[4 * i:4 * i + 4] for i in range(80)]
        print("  dZ hand-over per CTA pair (mean us / max us / %% already published): " + "  ".join(
            "%d:%.0f/%.0f/%d%%" % (i, l[0] / max(l[2], 1) / 1e3, l[1] / 1e3, 100 * l[3] / max(l[2], 1)) for i, l in enumerate(L) if l[2]))
        print("             wgrad producer total %d  flag wait %d (%.0f%%)  empty wait %d (%.0f%%)" % (v[7], v[5], 100. * v[5] / max(v[7], 1), v[6], 100. * v[6] / max(v[7], 1)))
