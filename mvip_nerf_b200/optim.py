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

    # ---- CUDA-graph form -----------------------------------------------------------------------------------------
    def graph_state(self, device):
        """Device-resident (lr, step) pair read by the captured Adam launch (mvip_adam_step_dev)."""
        if getattr(self, "_g_lr", None) is None:
            self._g_lr = torch.zeros(1, dtype=torch.float32, device=device)
            self._g_step = torch.zeros(1, dtype=torch.int64, device=device)
            self._g_lr_host = None
        return self._g_lr, self._g_step

    @torch.no_grad()
    def step_captured(self):
        """The update as ONE launch whose learning rate and step number come from device memory: call it inside CUDA-graph
        capture (graph.GraphedTrainStep does); every parameter must already have a gradient and all must share one step count.
        Python-side `state[p]['step']` is brought up to date by sync_graph_steps()."""
        lib = _lib.load()
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            lr_dev, step_dev = self.graph_state(params[0].device)
            step_dev.add_(1)
            ps, gs, ms, vs, sizes = [], [], [], [], []
            for p in params:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdam needs dense contiguous fp32 CUDA parameters and gradients")
                st = self._init_state(p)
                ps.append(p.data_ptr()); gs.append(p.grad.data_ptr())
                ms.append(st["exp_avg"].data_ptr()); vs.append(st["exp_avg_sq"].data_ptr())
                sizes.append(p.numel())
            n = len(ps)
            arr = lambda xs: (ctypes.c_void_p * n)(*xs)  # noqa: E731
            beta1, beta2 = group["betas"]
            rc = lib.mvip_adam_step_dev(arr(ps), arr(gs), arr(ms), arr(vs), (ctypes.c_int64 * n)(*sizes), n,
                                        ctypes.c_void_p(lr_dev.data_ptr()), float(beta1), float(beta2), float(group["eps"]),
                                        ctypes.c_void_p(step_dev.data_ptr()), ops._stream())
            _lib.check(rc, "mvip_adam_step_dev")

    def graph_begin(self, device):
        """Before capture: materialise the state tensors and seed the device step counter from the python-side count."""
        if len(self.param_groups) != 1:
            raise RuntimeError("FusedAdam CUDA-graph form keeps ONE device-resident (lr, step) pair: use a single param group "
                               "(create_nerf builds one, run.py:1536)")
        steps = set()
        for group in self.param_groups:
            for p in group["params"]:
                steps.add(int(self._init_state(p)["step"]))
        if len(steps) > 1:
            raise RuntimeError("FusedAdam: parameters must share the step count")
        lr_dev, step_dev = self.graph_state(device)
        step_dev.fill_(steps.pop() if steps else 0)

    def graph_set_lr(self):
        """Before each replay: push param_group['lr'] (the training loop decays it every step, run.py:1031-1039) if it changed."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._g_lr_host:
            self._g_lr.fill_(lr)
            self._g_lr_host = lr

    def sync_graph_steps(self):
        """Python-side step counts after n graph replays (state_dict() / checkpoints read them)."""
        if getattr(self, "_g_lr", None) is None or not getattr(self, "_g_dirty", False):
            return
        self._g_dirty = False
        s = float(int(self._g_step.item()))
        for group in self.param_groups:
            for p in group["params"]:
                if p in self.state:
                    self.state[p]["step"] = torch.tensor(s, dtype=torch.float32)
        ops.param_epoch += 1

    def state_dict(self):
        self.sync_graph_steps()
        return super().state_dict()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        self.sync_graph_steps()        # steps taken by graph replays, if any
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
            if getattr(self, "_g_lr", None) is not None:
                self._g_step.fill_(step_no)
            ops.launch_count += (n + 63) // 64
            # the update happens outside torch's view (no version-counter bump): tell NeRF.packed() to re-pack its bf16 blob
            ops.param_epoch += 1
        return loss
