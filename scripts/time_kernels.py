"""Ad-hoc kernel timings on one B200 (CUDA events, L2 flushed between iterations)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mvip_nerf_b200 import ops
from oracle import nerf_oracle as orc

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))

p = orc.init_params(1)
blob = ops.mlp_pack([torch.from_numpy(p[n]).to(dev) for n in ops.PARAM_ORDER])
for P in (4096 * 64, 4096 * 128, 32768 * 128):
    pts = torch.rand(P, 3, device=dev) * 4 - 2
    dirs = torch.nn.functional.normalize(torch.randn(P, 3, device=dev), dim=-1)
    med, best = timeit(lambda: ops.mlp_forward(blob, pts=pts, dirs=dirs))
    fl = P * 1186816
    print("mlp_forward P=%d: median %.3f ms best %.3f ms -> %.1f TFLOP/s (best %.1f) = %.1f%% of 1654" % (P, med, best, fl / med / 1e9, fl / best / 1e9, 100 * fl / best / 1e9 / 1654.1))
    if P <= 4096 * 128:
        med, best = timeit(lambda: ops.mlp_forward(blob, pts=pts, dirs=dirs, want_stash=True))
        print("   with stash: median %.3f ms best %.3f -> %.1f TFLOP/s" % (med, best, fl / med / 1e9))
for N, S in ((32768, 64), (32768, 128), (262144, 128)):
    raw = torch.randn(N, S, 4, device=dev); z = torch.sort(torch.rand(N, S, device=dev) * 6 + 1.2, -1)[0]; rd = torch.randn(N, 3, device=dev)
    med, best = timeit(lambda: ops.composite_forward(raw, z, rd, None, True))
    by = N * (S * 20 + 12 + S * 4 + 24)
    print("composite_fwd N=%d S=%d: median %.4f ms best %.4f -> %.0f GB/s (%.1f%% of 6545)" % (N, S, med, best, by / best / 1e6, 100 * by / best / 1e6 / 6545))
    g = [torch.randn(N, 3, device=dev), torch.randn(N, device=dev), torch.randn(N, device=dev), torch.randn(N, device=dev)]
    med, best = timeit(lambda: ops.composite_backward(raw, z, rd, None, True, False, *g))
    by = N * (S * 20 + 12 + 24 + S * 16)
    print("composite_bwd N=%d S=%d: median %.4f ms best %.4f -> %.0f GB/s" % (N, S, med, best, by / best / 1e6))
N = 262144
z = torch.sort(torch.rand(N, 64, device=dev) * 6 + 1.2, -1)[0]; w = torch.rand(N, 64, device=dev) ** 4
for name, u in (("det", torch.linspace(0, 1, 64, device=dev)), ("rand", torch.rand(N, 64, device=dev))):
    med, best = timeit(lambda: ops.sample_fine(z, w, u))
    by = N * (512 + (256 if u.dim() == 2 else 0) + 256 + 512 + 4)
    print("sample_fine %s N=%d: median %.4f ms best %.4f -> %.0f GB/s" % (name, N, med, best, by / best / 1e6))
rays = torch.rand(N, 11, device=dev) + 1; tv = torch.linspace(0, 1, 64, device=dev); tr = torch.rand(N, 64, device=dev)
med, best = timeit(lambda: ops.sample_coarse(rays, tv, tr, True))
print("sample_coarse N=%d: median %.4f ms best %.4f -> %.0f GB/s" % (N, med, best, N * (512 + 8) / best / 1e6))
d = torch.rand(512, 512, device=dev) + 3
med, best = timeit(lambda: ops.normal_forward(d, 500., 500., 256., 256., 31))
print("normal_fwd 512x512: median %.4f ms" % med)
