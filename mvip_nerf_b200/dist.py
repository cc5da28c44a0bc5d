"""Multi-GPU plumbing: one process per GPU (torchrun), rays sharded across ranks.

Replaces the reference's nn.DataParallel (run.py:1491,1527), which re-broadcasts the weights and
scatters/gathers [P,90] inputs on every MLP call.  Here the weights are replicated once, each rank
renders its contiguous slice of the ray batch with no data-path communication, and

  * a render ends with ONE gather of (rgb, disp, acc, depth) = 24 B/ray to rank 0;
  * a training step ends with ONE sum-allreduce per network of the flattened fp32 parameter gradients
    (2 x 595,844 x 4 B = 4.77 MB) over NCCL/NVLink; the fine network's allreduce is issued as soon as
    its gradients exist, so it overlaps the coarse network's backward when autograd runs them in
    sequence (the two branches are independent: z_samples is detached, run.py:1812).

The host logic is backend-agnostic (gloo on CPU tensors in tests/, nccl on the GPU box).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def world():
    return dist.get_world_size() if dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def shard_bounds(n, r=None, w=None):
    """Contiguous, balanced slice [lo, hi) of n units for rank r of w (first n % w ranks get one extra)."""
    r = rank() if r is None else r
    w = world() if w is None else w
    base, extra = divmod(n, w)
    lo = r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def shard_rows(t, r=None, w=None):
    lo, hi = shard_bounds(t.shape[0], r, w)
    return t[lo:hi]


def gather_rows(local, n_total, dst=0):
    """Concatenates per-rank row slices (shard_bounds order) on rank dst; other ranks get None.
    One collective: all ranks pad to the largest shard and all_gather into a single buffer."""
    if world() == 1:
        return local
    w = world()
    sizes = [shard_bounds(n_total, r, w) for r in range(w)]
    max_rows = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((max_rows,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    out = local.new_empty((w * max_rows,) + tuple(local.shape[1:]))
    if local.is_cuda:
        dist.all_gather_into_tensor(out, pad)
    else:
        dist.all_gather([out[r * max_rows:(r + 1) * max_rows] for r in range(w)], pad)
    if rank() != dst:
        return None
    return torch.cat([out[r * max_rows:r * max_rows + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)


def render_sharded(render_fn, rays_flat, **kwargs):
    """Renders rank-local rows of rays_flat [N, C] with render_fn(rows, **kwargs) -> dict of [n, ...] tensors and
    gathers rgb_map / disp_map / acc_map / depth_map on rank 0 as one packed [N, 6] tensor (24 B/ray)."""
    n = rays_flat.shape[0]
    ret = render_fn(shard_rows(rays_flat), **kwargs)
    packed = torch.cat([ret["rgb_map"], ret["disp_map"][:, None], ret["acc_map"][:, None], ret["depth_map"][:, None]], -1)
    full = gather_rows(packed.contiguous(), n)
    if full is None:
        return None
    return {"rgb_map": full[:, 0:3], "disp_map": full[:, 3], "acc_map": full[:, 4], "depth_map": full[:, 5]}


def _flat_view(grads):
    """If the gradient tensors are consecutive views of ONE contiguous buffer (what ops.mlp_backward returns: 24 views of
    one flat fp32 tensor per network), returns that buffer as a 1-D tensor sharing their memory; otherwise None."""
    g0 = grads[0]
    if not g0.is_contiguous():
        return None
    st = g0.untyped_storage()
    off = g0.storage_offset()
    total = 0
    for g in grads:
        if (g.dtype != g0.dtype or g.device != g0.device or not g.is_contiguous() or
                g.untyped_storage().data_ptr() != st.data_ptr() or g.storage_offset() != off + total):
            return None
        total += g.numel()
    return torch.empty(0, dtype=g0.dtype, device=g0.device).set_(st, off, (total,))


class GradAllReducer:
    """Sum-allreduce of the gradients of a list of parameters as ONE flat fp32 bucket (async).  When the gradients already
    live back to back in one buffer (the MLP backward writes them that way) the collective runs IN PLACE on it: no
    concatenation before and no 24 copy kernels after."""

    def __init__(self, params):
        self.params = [p for p in params]
        self.handle = None
        self.flat = None
        self.in_place = False

    def start(self):
        if world() == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        view = _flat_view(grads) if all(p.grad is not None for p in self.params) else None
        self.in_place = view is not None
        self.flat = view if self.in_place else torch.cat([g.reshape(-1) for g in grads])
        self.handle = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True)

    def finish(self):
        if self.handle is None:
            return
        self.handle.wait()
        if not self.in_place:
            off = 0
            for p in self.params:
                n = p.numel()
                g = self.flat[off:off + n].view_as(p)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
        self.handle, self.flat = None, None


def allreduce_grads(param_groups):
    """param_groups: iterable of parameter lists (e.g. [coarse params, fine params]); one bucket each."""
    reducers = [GradAllReducer(ps) for ps in param_groups]
    for r in reducers:
        r.start()
    for r in reducers:
        r.finish()
