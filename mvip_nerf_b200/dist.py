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
    return _gather_uneven(local, [hi - lo for lo, hi in (shard_bounds(n_total, r, w) for r in range(w))], dst)


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


def view_row_bands(n_views, H, r=None, w=None):
    """Rows of a stack of n_views images of H rows, split evenly over the ranks: rank r gets the contiguous range
    shard_bounds(n_views * H) of the global row index; returned as [(view, first_row, n_rows)] (a range may straddle views)."""
    lo, hi = shard_bounds(n_views * H, r, w)
    bands = []
    while lo < hi:
        v, i0 = divmod(lo, H)
        n = min(H - i0, hi - lo)
        bands.append((v, i0, n))
        lo += n
    return bands


def render_views_sharded(render_fn, poses, H, W, focal, near, far, normal_k=31, with_normals=True, **kwargs):
    """The guidance render batch (BASELINE cfg 5; render_path_4view + the normal-map branch of train(), DS_NeRF/run.py:948-982):
    V views of H x W, rgb + disp + acc + depth per pixel and, on rank 0, the depth-derived normal map of each view.

    Rows of the V*H image rows are sharded over the ranks (row bands -> mvip_rays_from_pose windows), every rank renders its
    bands with render_fn(H, W, focal, c2w=pose, patch=(i0, 0, n, W), near=, far=, **kwargs) -> [rgb, disp, acc, depth, extras],
    ONE gather of 24 B/pixel brings them to rank 0, which runs the normal-map kernels (a 31 x 31 stencil over the whole
    image; 24 B/pixel of traffic).  Returns on rank 0 a dict of [V,H,W,...] tensors (+ 'normal' [V,3,H,W], mapped to
    (n + 1) / 2 as in run.py:965); None on the other ranks."""
    from .run_nerf_helpers import depth2normal
    V = len(poses)
    pieces = []
    for v, i0, n in view_row_bands(V, H):
        rgb, disp, acc, depth, _ = render_fn(H, W, focal, c2w=poses[v][:3, :4], patch=(i0, 0, n, W), near=near, far=far, **kwargs)
        pieces.append(torch.cat([rgb.reshape(-1, 3), disp.reshape(-1, 1), acc.reshape(-1, 1), depth.reshape(-1, 1)], -1))
    some = poses[0]
    dev = some.device if torch.is_tensor(some) and some.is_cuda else (
        torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
    packed = torch.cat(pieces, 0) if pieces else torch.empty((0, 6), device=dev)
    # row bands are whole rows, so the row-sharded gather of pixels == gather_rows with W pixels per row
    full = _gather_uneven(packed.contiguous(), [(b[1] - b[0]) * W for b in (shard_bounds(V * H, r, world()) for r in range(world()))])
    if full is None:
        return None
    full = full.view(V, H, W, 6)
    out = {"rgb_map": full[..., 0:3], "disp_map": full[..., 3], "acc_map": full[..., 4], "depth_map": full[..., 5]}
    if with_normals:
        K = [[focal, 0., W / 2], [0., focal, H / 2], [0., 0., 1.]]
        out["normal"] = torch.cat([(depth2normal(out["depth_map"][v].contiguous(), K, normal_k) + 1) / 2 for v in range(V)], 0)
    return out


def _gather_uneven(local, sizes, dst=0):
    """all ranks contribute `sizes[r]` rows; rank dst gets the concatenation (one padded all_gather), others None."""
    if world() == 1:
        return local
    w = world()
    max_rows = max(sizes)
    pad = local.new_zeros((max_rows,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    out = local.new_empty((w * max_rows,) + tuple(local.shape[1:]))
    if local.is_cuda:
        dist.all_gather_into_tensor(out, pad)
    else:
        dist.all_gather([out[r * max_rows:(r + 1) * max_rows] for r in range(w)], pad)
    if rank() != dst:
        return None
    return torch.cat([out[r * max_rows:r * max_rows + sizes[r]] for r in range(w)], 0)


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
