"""Adam with ONE kernel launch per step over all parameters.

Subclass of (and drop-in for) the `torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))` the reference builds in
create_nerf (DS_NeRF/run.py:1536-1537): same constructor arguments, same `param_groups` (the training loop writes the
decayed learning rate into `param_group['lr']`, run.py:1031-1039), and the same per-parameter state layout
(`step`, `exp_avg`, `exp_avg_sq`), so `state_dict()` / `load_state_dict()` interchange with checkpoints written by
torch.optim.Adam (run.py:1043-1053, 1557).  weight_decay / amsgrad / maximize are not implemented (unused by the reference).
"""
import ctypes

import torch

from . import _lib, ops


class FusedAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("FusedAdam implements torch.optim.Adam with weight_decay=0, amsgrad=False")
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        # a torch.optim.Adam in every respect (param_groups, defaults, state layout, state_dict) except step()
        super().__init__(params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False)

    def _init_state(self, p):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            ps, gs, ms, vs, sizes = [], [], [], [], []
            step_no = None
            keep = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("FusedAdam needs dense fp32 CUDA parameters and gradients (no CPU fallback)")
                st = self._init_state(p)
                s = int(st["step"]) + 1 if torch.is_tensor(st["step"]) else int(st["step"]) + 1
                if step_no is None:
                    step_no = s
                elif s != step_no:
                    raise RuntimeError("FusedAdam: parameters of one group must share the step count")
                st["step"] = torch.tensor(float(s), dtype=torch.float32)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous parameters")
                keep.append(g)
                ps.append(p.data_ptr()); gs.append(g.data_ptr())
                ms.append(st["exp_avg"].data_ptr()); vs.append(st["exp_avg_sq"].data_ptr())
                sizes.append(p.numel())
            if not ps:
                continue
            n = len(ps)
            arr = lambda xs: (ctypes.c_void_p * n)(*xs)  # noqa: E731
            beta1, beta2 = group["betas"]
            rc = lib.mvip_adam_step(arr(ps), arr(gs), arr(ms), arr(vs), (ctypes.c_int64 * n)(*sizes), n, float(group["lr"]),
                                    float(beta1), float(beta2), float(group["eps"]), step_no, ops._stream())
            _lib.check(rc, "mvip_adam_step")
            ops.launch_count += (n + 63) // 64
            # the update happens outside torch's view (no version-counter bump): tell NeRF.packed() to re-pack its bf16 blob
            ops.param_epoch += 1
        return loss
