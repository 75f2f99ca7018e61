"""
Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  HMC/SGLD chains are independent (no data-path collective, only a final gather of the collected
samples); SVGD all-gathers particles and gradients once per epoch and all-reduces the 2048-bin select
histograms of the exact median; BBB all-reduces its (D,2) gradient.  With no process group everything is a
no-op on one device.
"""

import torch
import torch.distributed as dist


def active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank():
    return dist.get_rank() if active() else 0


def world():
    return dist.get_world_size() if active() else 1


def shard(n, r=None, w=None):
    """Contiguous share [lo, hi) of n units for rank r of w (first n % w ranks get one extra)."""
    r = rank() if r is None else r
    w = world() if w is None else w
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def broadcast_(t, src=0):
    """Rank `src`'s copy of `t` on every rank (replicated state must not depend on how the user seeded the ranks)."""
    if active():
        if not t.is_cuda and dist.get_backend() == "nccl":      # NCCL moves device memory only
            d = t.cuda()
            dist.broadcast(d, src=src)
            t.copy_(d)
        else:
            dist.broadcast(t, src=src)
    return t


_COMM = None


def comm():
    """The library's own communicator (ocbnn_comm, NCCL inside libocbnn_b200.so) for the one-call SVGD / BBB steps; None on one
    rank.  torch.distributed only carries the 128-byte id from rank 0 to the others."""
    global _COMM
    if not active():
        return None
    if _COMM is None or _COMM.world != world() or _COMM.rank != rank():
        from . import engine as E

        def exchange(ident):
            box = [ident]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        _COMM = E.Comm(rank(), world(), exchange)
    return _COMM


def all_sum_(t):
    if active():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_sum(t):
    t = t.clone()
    all_sum_(t)
    return t.item()


def gather_cat(t, dim=0):
    """Concatenate the ranks' shards of `t` along `dim` (shards may differ by one row)."""
    if not active():
        return t
    w = world()
    t = t.movedim(dim, 0).contiguous()
    sizes = torch.zeros(w, dtype=torch.int64, device=t.device)
    sizes[rank()] = t.shape[0]
    dist.all_reduce(sizes)
    sizes = [int(s) for s in sizes.cpu()]
    mx = max(sizes)
    if t.shape[0] < mx:
        pad = torch.zeros((mx - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], 0)
    out = torch.empty((w * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t) if t.is_cuda else dist.all_gather(list(out.view((w, mx) + tuple(t.shape[1:])).unbind(0)), t)
    parts = [out[r * mx:r * mx + sizes[r]] for r in range(w)]
    return torch.cat(parts, 0).movedim(0, dim).contiguous()


def gather_rows(t, n_total, out=None):
    """All-gather of row shards laid out by shard(n_total): (rows_r, ...) on rank r -> (n_total, ...) everywhere.  The shard
    sizes follow from n_total alone, so nothing is exchanged and nothing synchronises with the host (gather_cat asks the
    other ranks for their sizes).  `out`: optional preallocated result."""
    if not active():
        return t
    w = world()
    sizes = [shard(n_total, r, w)[1] - shard(n_total, r, w)[0] for r in range(w)]
    t = t.contiguous()
    if out is None:
        out = torch.empty((n_total,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    if all(sz == sizes[0] for sz in sizes):
        if t.is_cuda:
            dist.all_gather_into_tensor(out, t)
        else:
            dist.all_gather(list(out.split(sizes, 0)), t)
        return out
    mx = max(sizes)                             # ragged shards: pad to the largest, gather, drop the padding
    if t.shape[0] < mx:
        t = torch.cat([t, torch.zeros((mx - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)], 0)
    tmp = torch.empty((w * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    if t.is_cuda:
        dist.all_gather_into_tensor(tmp, t)
    else:
        dist.all_gather(list(tmp.split(mx, 0)), t)
    torch.cat([tmp[r * mx:r * mx + sizes[r]] for r in range(w)], 0, out=out)
    return out

