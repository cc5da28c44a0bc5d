"""Time the MLP backward (fine network size: 524,288 points) with CUDA events; env selects the variant."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mvip_nerf_b200 import ops, _lib
from oracle import nerf_oracle as orc
dev = "cuda"
p = orc.init_params(1)
blob = ops.mlp_pack([torch.from_numpy(p[n]).to(dev) for n in ops.PARAM_ORDER])
P = int(os.environ.get("P", 524288))
pts = torch.rand(P, 3, device=dev) * 4 - 2
dirs = torch.nn.functional.normalize(torch.randn(P, 3, device=dev), dim=-1)
raw, stash = ops.mlp_forward(blob, pts=pts, dirs=dirs, want_stash=True)
d_raw = torch.randn(P, 4, device=dev)
lib = _lib.load()
ws = ops._aligned_bytes(lib.mvip_mlp_backward_workspace_bytes(P), dev)
sizes = [int(torch.Size(shp).numel()) for shp in ops.PARAM_SHAPES]
grads = [torch.empty(n, device=dev) for n in sizes]
arr = (ctypes.c_void_p * len(grads))(*[g.data_ptr() for g in grads])
def run(mask):
    rc = lib.mvip_mlp_backward_phases(ops._ptr(blob), ops._ptr(d_raw), P, ops._ptr(stash), ops._ptr(ws), arr, 0, mask, ops._stream())
    assert rc == 0, lib.mvip_last_error()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(mask, iters=6):
    run(mask); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(mask); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)
tag = "fused=%s dgrad_sms=%s" % (os.environ.get("MVIP_BWD_FUSED", "1"), os.environ.get("MVIP_BWD_DGRAD_SMS", "auto"))
if os.environ.get("MVIP_BWD_FUSED", "1") == "0":
    print("%s: dgrad %.3f ms  wgrad %.3f ms  both %.3f ms" % (tag, t(1), t(2), t(3)))
else:
    print("%s: dgrad+wgrad %.3f ms" % (tag, t(3)))

out = (ctypes.c_ulonglong * 8)()
run(2 if os.environ.get("MVIP_BWD_FUSED", "0") != "1" else 3); lib.mvip_debug_wgrad_profile(out); v = list(out)
print("  wgrad CTA 0: producer empty-wait %.0f%% of %d cyc (flag wait %.0f%%); issuer full-wait %.0f%% of %d; bias warps full-wait %.0f%% of %d" % (
    100 * v[0] / max(v[1], 1), v[1], 100 * v[6] / max(v[1], 1), 100 * v[2] / max(v[3], 1), v[3], 100 * v[4] / max(v[5], 1), v[5]))
