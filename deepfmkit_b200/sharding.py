"""Sharding of fit units (buffers, channels, realisations) over the GPUs of one box.

The readout path has no exchange step: every buffer / channel is independent.  A record
is cut into contiguous time slabs aligned to buffer boundaries, one per rank; the only shared datum is
the 4-double seed from buffer 0 (fitters.py:404-417): rank 0 fits that one buffer and broadcasts the result.  Result rows
are gathered on rank 0 in slab order.  One process per GPU; ``torch.distributed`` (nccl on the GPU box,
gloo in the CPU tests) is used only for that broadcast and the final gather of the small row table.
"""
from __future__ import annotations

import numpy as np


def slab_bounds(n_units: int, world_size: int, rank: int):
    """Contiguous range [lo, hi) of units owned by ``rank``; sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    base, extra = divmod(int(n_units), world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def all_slab_bounds(n_units: int, world_size: int):
    return [slab_bounds(n_units, world_size, r) for r in range(world_size)]


def broadcast_seed(seed_row, src: int = 0, group=None):
    """Broadcast the 4-double seed (host tensor for gloo, device tensor for nccl) from ``src``."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(seed_row, dtype=np.float64).copy())
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=src, group=group)
    return t.cpu().numpy()


def to_host(t):
    """Device tensor -> numpy through a page-locked buffer (torch caches it): a DMA at the PCIe rate instead of the
    driver's staged copy into pageable memory, which is what ``.cpu()`` does and costs 5-10x as long for large tables."""
    import torch
    if not t.is_cuda:
        return t.numpy()
    out = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return out.numpy()


def gather_rows(local_rows, n_units: int, dst: int = 0, group=None, return_tensor: bool = False):
    """Gather per-rank row blocks [hi-lo, W] into one [n_units, W] table on ``dst`` (None elsewhere).

    ``local_rows`` may be a numpy array or a torch tensor.  Under nccl a CUDA tensor is gathered where it lies --
    GPU to GPU over NVLink, no host staging -- and the table comes back to the host in one copy (or stays on the
    device with ``return_tensor``); under gloo everything is host memory."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bounds = all_slab_bounds(n_units, world)
    on_gpu = dist.get_backend(group) == "nccl"
    t = local_rows if isinstance(local_rows, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_rows, dtype=np.float64))
    if t.dim() != 2:
        t = t.reshape(-1, 8)
    t = t.cuda() if on_gpu else t.cpu()
    width = t.shape[1]
    longest = max(hi - lo for lo, hi in bounds)
    if t.shape[0] != longest:  # ragged split: pad to the longest block
        pad = torch.zeros((longest, width), dtype=torch.float64, device=t.device)
        pad[: t.shape[0]] = t
        t = pad
    t = t.contiguous()
    out = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, out, dst=dst, group=group)
    if rank != dst:
        return None
    table = torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, bounds)], dim=0)
    return table if return_tensor else to_host(table)


def nls_fit_sharded(x_slab, n_buffers_total, R, ndata, w0, init, device=None, group=None, tunables_from=None,
                    chunks_per_rank=1, gather=True, return_tensor=False, first_buffer=None):
    """NLS readout of one long record sharded over the ranks of ``group`` as contiguous buffer-aligned slabs.

    Every rank calls this with *its* slab: ``x_slab`` holds buffers ``slab_bounds(n_buffers_total, world, rank)``
    of the record -- a CUDA float64 tensor (used in place) or a host numpy array (streamed through the library's
    staged host path, slab copies overlapping the kernels).  Rank 0 fits buffer 0 *alone* from ``init``
    (fitters.py:404-405: one demodulation + one cold fit, well under a millisecond) and the fitted
    [amp, m, phi, psi] is broadcast (32 bytes); then all ranks, rank 0 included, fit their slabs concurrently as
    ``chunks_per_rank`` chains started from that seed (fitters.py:407-417 -- a rank boundary is a chunk boundary,
    i.e. the reference's pool schedule with ``n_cores = world * chunks_per_rank``).  No other exchange: the kernels
    of different ranks never communicate.  ``first_buffer``: the R samples of the record's buffer 0, given to every
    rank (160 kB for config 2) -- each rank then fits it itself and nothing is exchanged before the slab kernels.
    Returns the full ``[n_buffers_total, 8]`` row table on rank 0 (None elsewhere), or with ``gather=False`` this rank's
    own rows.
    """
    import torch
    import torch.distributed as dist
    from . import _lib
    from . import fit as fit_tunables

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = slab_bounds(n_buffers_total, world, rank)
    nb = hi - lo
    if device is None:
        device = torch.cuda.current_device()
    dev = torch.device("cuda", device)
    ctx = _lib.get_context(device)
    opts = fit_tunables.current_lm_opts(tunables_from)
    on_device = isinstance(x_slab, torch.Tensor) and x_slab.is_cuda
    first = 1 if rank == 0 else 0  # rank 0's slab opens with buffer 0, which is fitted cold
    if on_device:
        xt = x_slab.contiguous().view(-1)
        if xt.numel() < nb * R:
            raise ValueError(f"rank {rank}: slab holds {xt.numel()} samples, needs {nb * R}")
        rows = torch.zeros((nb, _lib.ROW_STRIDE), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            ctx.use_torch_stream()
            try:
                seed_local = np.zeros(4)
                if first_buffer is not None:  # every rank fits buffer 0 redundantly: no exchange before the kernels
                    fb = first_buffer if isinstance(first_buffer, torch.Tensor) else torch.from_numpy(
                        np.ascontiguousarray(first_buffer, dtype=np.float64))
                    fb = fb.to(dev).contiguous().view(-1)
                    row0 = rows[:1] if (rank == 0 and nb > 0) else torch.empty((1, _lib.ROW_STRIDE), dtype=torch.float64, device=dev)
                    ctx.nls_fit_dev(fb.data_ptr(), 1, R, ndata, w0, init, _lib.SCHED_EACH, opts, row0.data_ptr())
                    seed = row0[0, :4].cpu().numpy()
                else:
                    if rank == 0 and nb > 0:
                        ctx.nls_fit_dev(xt.data_ptr(), 1, R, ndata, w0, init, _lib.SCHED_EACH, opts, rows.data_ptr())
                        torch.cuda.current_stream(dev).synchronize()
                        seed_local = rows[0, :4].cpu().numpy()
                    seed = broadcast_seed(seed_local, src=0, group=group)
                if nb - first > 0:
                    ctx.nls_fit_seeded_dev(xt.data_ptr() + first * R * 8, nb - first, R, ndata, w0, seed, opts,
                                           rows.data_ptr() + first * _lib.ROW_STRIDE * 8, chunks=chunks_per_rank)
                torch.cuda.current_stream(dev).synchronize()
            finally:
                ctx.use_default_stream()
        local = rows  # stays on the device: the gather below is GPU to GPU
    else:
        xs = np.ascontiguousarray(np.asarray(x_slab, dtype=np.float64)).reshape(-1)
        if xs.size < nb * R:
            raise ValueError(f"rank {rank}: slab holds {xs.size} samples, needs {nb * R}")
        local = np.zeros((nb, _lib.ROW_STRIDE))
        seed_local = np.zeros(4)
        with torch.cuda.device(dev):
            if rank == 0 and nb > 0:
                local[:1] = ctx.nls_fit_host(xs[:R], R, ndata, w0, init, seeded=True, opts=opts)
                seed_local = local[0, :4].copy()
            seed = broadcast_seed(seed_local, src=0, group=group)
            if nb - first > 0:
                local[first:] = ctx.nls_fit_seeded_host(xs[first * R: nb * R], R, ndata, w0, seed, chunks=chunks_per_rank,
                                                        opts=opts)
    if not gather:
        return local if (return_tensor or not isinstance(local, torch.Tensor)) else to_host(local)
    return gather_rows(local, n_buffers_total, dst=0, group=group, return_tensor=return_tensor)


def ekf_fit_sharded(z_channels, n_channels_total, f_samp, f_mod, n, device=None, group=None, **ekf_kwargs):
    """EKF over many channels split over the ranks by channel (time is sequential, channels are independent).

    Every rank passes the records of *its* channels ``slab_bounds(n_channels_total, world, rank)`` as ``[C_local, T]``;
    rank 0 gets the ``[n_channels_total, nbuf, 8]`` table, the others None.  No exchange but the final gather.
    """
    import torch
    import torch.distributed as dist
    from .fitters import ekf_fit_batch

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = slab_bounds(n_channels_total, world, rank)
    if device is None:
        device = torch.cuda.current_device()
    z = np.atleast_2d(np.asarray(z_channels, dtype=np.float64)) if not isinstance(z_channels, torch.Tensor) else z_channels
    if z.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: got {z.shape[0]} channels, owns {hi - lo}")
    rows = ekf_fit_batch(z, f_samp, f_mod, n, device=device, return_tensor=True, **ekf_kwargs) if hi > lo else np.zeros((0, 0, 8))
    nbuf = rows.shape[1] if hi > lo else 0
    sizes = [None] * world
    dist.all_gather_object(sizes, nbuf, group=group)
    nbuf = max(sizes)
    flat = rows.reshape(hi - lo, nbuf * 8) if hi > lo else np.zeros((0, nbuf * 8))
    table = gather_rows(flat, n_channels_total, dst=0, group=group)
    return None if table is None else table.reshape(n_channels_total, nbuf, 8)
