"""Write-only / read-only / copy HBM bandwidth on one B200 (torch elementwise kernels, CUDA events)."""
import torch
dev = "cuda"
n = 1 << 31   # 2 Gi bf16 = 4 GiB
a = torch.empty(n, dtype=torch.bfloat16, device=dev)
b = torch.empty(n, dtype=torch.bfloat16, device=dev)
def t(fn, iters=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.fill_(1.0)); print("fill  (write only): %.3f ms  %.0f GB/s" % (ms, n * 2 / ms / 1e6))
ms = t(lambda: a.zero_()); print("zero_ (memset)    : %.3f ms  %.0f GB/s" % (ms, n * 2 / ms / 1e6))
ms = t(lambda: b.copy_(a)); print("copy  (read+write): %.3f ms  %.0f GB/s total" % (ms, 2 * n * 2 / ms / 1e6))
ms = t(lambda: a.sum()); print("sum   (read only) : %.3f ms  %.0f GB/s" % (ms, n * 2 / ms / 1e6))
