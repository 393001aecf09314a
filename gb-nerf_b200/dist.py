"""Multi-GPU plumbing: one process per GPU, rays sharded by contiguous blocks, weights replicated.

Replaces the reference's ``nn.DataParallel`` wrapping (run.py:2020,2056), which scatters every 65,536-point
MLP call, re-broadcasts all parameters on every forward and gathers to GPU 0.  Here:

  * forward needs no collective: rank g renders rays [g*R/G, (g+1)*R/G) of the frame / batch;
  * training has ONE exchange: an all-reduce of one flat fp32 gradient bucket (2 x 595,844 floats) per step;
  * inference ends with ONE gather of the 24 B/ray image outputs.

Everything here is written against ``torch.distributed`` only, so the host logic runs under the ``gloo``
backend on CPU tensors in the tests and under NCCL (NVLink 5 / NVSwitch) on the GPU box.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) — (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world_size):
    """Contiguous block [lo, hi) of ``n`` items owned by ``rank``; sizes differ by at most one, in rank order."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(int(n), world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays_flat, rank=None, world_size=None):
    """This rank's rows of a flat [R, C] ray batch (a view, no copy)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    lo, hi = shard_bounds(rays_flat.shape[0], rank, world_size)
    return rays_flat[lo:hi]


class GradBucket:
    """One flat fp32 buffer aliasing the gradients of every parameter, all-reduced with a single collective.

    ``p.grad`` of each parameter becomes a view into the bucket, so backward kernels and autograd write straight
    into it and no flatten/unflatten copy happens per step.
    """

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def rebind(self):
        """Re-attach the views if something replaced ``p.grad`` (e.g. ``zero_grad(set_to_none=True)``): a replaced
        gradient is copied into the bucket, a missing one (parameter unused this step) counts as zero."""
        off = 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                view.zero_()
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view
            off += p.numel()

    def all_reduce(self, average=False, async_op=False):
        """SUM over ranks (the per-rank losses are already scaled by 1/R_global).

        Safe after ``optimizer.zero_grad()`` (which sets grads to None, so that autograd allocates fresh ones outside
        the bucket): gradients that no longer alias the bucket are copied in and re-attached first."""
        self.rebind()
        rank, ws = world()
        if ws == 1:
            return None
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if average:
            if async_op:
                work.wait()
            self.flat.div_(ws)
        return work


def gather_rows(local, total_rows, dst=0):
    """Concatenate per-rank row blocks (sharded by ``shard_bounds``) on rank ``dst``; returns None elsewhere.

    One ``gather`` to ``dst``: only that rank allocates and receives the whole [total_rows, ...] result, every other
    rank sends its block once (24 B/ray for the image outputs).  ``dst=None`` gathers on every rank
    (``all_gather``).  Blocks may differ by one row, so they travel padded to the largest.
    """
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_bounds(total_rows, r, ws) for r in range(ws)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < maxn:
        pad = torch.cat([local, local.new_zeros((maxn - local.shape[0],) + tuple(local.shape[1:]))], 0)
    pad = pad.contiguous()
    even = all(hi - lo == maxn for lo, hi in sizes)
    if dst is not None and rank != dst:
        dist.gather(pad, None, dst=dst)
        return None
    buf = pad.new_empty((ws, maxn) + tuple(pad.shape[1:]))
    if dst is None:
        dist.all_gather(list(buf.unbind(0)), pad)
    else:
        dist.gather(pad, list(buf.unbind(0)), dst=dst)
    if even:
        return buf.view((ws * maxn,) + tuple(pad.shape[1:]))
    return torch.cat([buf[r, :hi - lo] for r, (lo, hi) in enumerate(sizes)], 0)


def render_sharded(render_fn, rays_flat, gather_keys=("rgb_map", "disp_map", "acc_map", "depth_map"), dst=0, **kw):
    """Render this rank's block of ``rays_flat`` with ``render_fn`` (= run.batchify_rays) and gather the
    image outputs packed as one [R, 6] tensor (24 B/ray: rgb(3), disp, acc, depth) on ``dst``."""
    rank, ws = world()
    mine = shard_rays(rays_flat, rank, ws)
    ret = render_fn(mine, **kw)
    cols = [ret[k].reshape(mine.shape[0], -1) for k in gather_keys]
    widths = [c.shape[1] for c in cols]
    packed = gather_rows(torch.cat(cols, 1), rays_flat.shape[0], dst=dst)
    if packed is None:
        return None, ret
    out, off = {}, 0
    for k, w in zip(gather_keys, widths):
        out[k] = packed[:, off:off + w] if w > 1 else packed[:, off]
        off += w
    return out, ret
