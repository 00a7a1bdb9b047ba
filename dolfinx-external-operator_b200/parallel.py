"""Multi-GPU plumbing: one process per GPU, quadrature points partitioned along the
existing cell partition (reference: each rank evaluates its own local+ghost cells,
external_operator.py:368-370, and `scatter_forward`s afterwards, :445).

For Quadrature spaces every DOF is cell-interior and ghosts are recomputed locally, so the
hot path needs NO data exchange.  The one collective is the all-reduce of the statistics
record (plastic / non-converged counts, Newton-iteration histogram: SUM; maxima: MAX) that
replaces the rank-local prints of demo_plasticity_mohr_coulomb.py:584-591: one all-gather of
the 1.7 KB record + a local combine, into a record SEPARATE from the one the kernels
accumulate into (`allreduce_stats`, C ABI `eo_allreduce_stats`).
"""

from __future__ import annotations

import numpy as np

from ._lib import EO_NITER_BINS

N_SUM = 4 + EO_NITER_BINS  # int64 entries of eo_stats' SUM block
N_MAX = 4  # f64 entries of the MAX block


def partition(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition of `n` quadrature points (or cells): [start, stop) of `rank`.
    The first n % world ranks get one extra item - the same rule DOLFINx's index maps use for
    an unpartitioned range."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, rem = divmod(int(n), world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def local_submesh(dofmap: np.ndarray, x_dofmap: np.ndarray, x: np.ndarray, rank: int, world: int) -> dict:
    """The cell block of `rank` as a self-contained mesh for `Tabulator` / `QuadratureForms`: cells [start, stop) of the
    contiguous partition, dofs and geometry nodes renumbered locally in order of first use.  Returns dict(cells (start,
    stop), dofmap, x_dofmap, x, n_dofs, dof_l2g, node_l2g).  Every local cell is an OWNED cell (cell integrals run over
    owned cells only, like DOLFINx's assemblers); dofs on the block interface appear on both ranks and their contributions
    are summed afterwards - `b.ghostUpdate(ADD, REVERSE)` in the reference's callbacks (demo_vm:512), `sum_shared` here."""
    start, stop = partition(dofmap.shape[0], rank, world)
    dm, xd = np.asarray(dofmap)[start:stop], np.asarray(x_dofmap)[start:stop]
    dof_l2g, dm_loc = np.unique(dm.reshape(-1), return_inverse=True)
    node_l2g, xd_loc = np.unique(xd.reshape(-1), return_inverse=True)
    return {"cells": (start, stop), "dofmap": dm_loc.reshape(dm.shape).astype(np.int32),
            "x_dofmap": xd_loc.reshape(xd.shape).astype(np.int32), "x": np.ascontiguousarray(np.asarray(x)[node_l2g]),
            "n_dofs": int(dof_l2g.size), "dof_l2g": dof_l2g, "node_l2g": node_l2g}


def sum_shared(b_local: np.ndarray, dof_l2g: np.ndarray, n_dofs_global: int, bs: int = 1, group=None) -> np.ndarray:
    """Global vector (bs * n_dofs_global) = sum over ranks of the local cell-integral vectors (host tensors; gloo or
    nccl-with-host-staging).  A dense all-reduce: fine for tests and moderate sizes; production codes exchange only the
    interface dofs with their neighbours (what PETSc's ghostUpdate does)."""
    import torch
    import torch.distributed as dist

    g = np.zeros((n_dofs_global, bs))
    g[dof_l2g] = np.asarray(b_local, dtype=np.float64).reshape(-1, bs)
    t = torch.from_numpy(g)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy().reshape(-1)


def stats_to_record(stats: dict) -> np.ndarray:
    """The statistics dict packed like `eo_stats` (include/eo_b200.h): N_SUM int64 then N_MAX float64, as one int64
    vector (the float64 block bit-cast) - the payload of the collective on either backend."""
    rec = np.zeros(N_SUM + N_MAX, dtype=np.int64)
    rec[0], rec[1], rec[2], rec[3] = stats["n_points"], stats["n_plastic"], stats["n_nonconverged"], stats["n_nonfinite"]
    rec[4:N_SUM] = stats["niter_hist"]
    rec[N_SUM:].view(np.float64)[:] = [stats["niter_max"], stats["f_max"], stats["res_max"], 0.0]
    return rec


def combine_records(records: np.ndarray) -> dict:
    """records (world, N_SUM + N_MAX) int64 -> the global statistics dict: SUM over the int64 block, MAX over the float64
    block (what `eo_stats_combine_kernel` does on the device)."""
    records = np.ascontiguousarray(records, dtype=np.int64).reshape(-1, N_SUM + N_MAX)
    s = records[:, :N_SUM].sum(axis=0)
    m = np.ascontiguousarray(records[:, N_SUM:]).view(np.float64).reshape(-1, N_MAX)
    m = np.where(np.isnan(m).all(axis=0), np.nan, np.nanmax(np.where(np.isnan(m), -np.inf, m), axis=0))
    return {
        "n_points": int(s[0]), "n_plastic": int(s[1]), "n_nonconverged": int(s[2]), "n_nonfinite": int(s[3]),
        "niter_hist": np.asarray(s[4:], dtype=np.int64).copy(),
        "niter_max": float(m[0]), "f_max": float(m[1]), "res_max": float(m[2]),
    }


def allreduce_stats_host(stats: dict, group=None) -> dict:
    """Global statistics from every rank's dict through torch.distributed with host tensors (gloo): ONE all-gather of
    the packed record, combined locally.  The input is not modified."""
    import torch
    import torch.distributed as dist

    rec = torch.from_numpy(stats_to_record(stats))
    out = [torch.empty_like(rec) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, rec, group=group)
    return combine_records(torch.stack(out).numpy())


class _DevView:
    """`__cuda_array_interface__` window on device memory owned by the library."""

    def __init__(self, ptr: int, n: int, typestr: str = "<i8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


def allreduce_stats_device(ctx, group=None, stats: dict | None = None) -> None:
    """The one collective of the hot path on a torch.distributed NCCL group: ctx's LOCAL statistics record (or `stats`,
    a dict read earlier) of every rank -> ctx's GLOBAL record (`ctx.stats_global()`).  One `all_gather_into_tensor` of
    the 1.7 KB record on the ctx's collective stream plus the library's combine kernel; asynchronous - it is ordered
    after the work queued on the compute stream so far and overlaps whatever is queued next.  The local record is not
    modified, so this may follow every evaluation of a record that keeps accumulating."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from ._lib import Stats

    world = dist.get_world_size(group)
    send, recv, stream = C.c_void_p(), C.c_void_p(), C.c_void_p()
    host = None
    if stats is not None:
        host = Stats.from_buffer_copy(stats_to_record(stats).tobytes())
    ctx.check(ctx.lib.eo_stats_collective_begin(ctx.handle, world, C.byref(host) if host is not None else None,
                                                C.byref(send), C.byref(recv), C.byref(stream)))
    nrec = N_SUM + N_MAX
    dev = f"cuda:{ctx.device}"
    t_send = torch.as_tensor(_DevView(send.value, nrec), device=dev)
    t_recv = torch.as_tensor(_DevView(recv.value, nrec * world), device=dev)
    with torch.cuda.stream(torch.cuda.ExternalStream(stream.value, device=dev)):
        dist.all_gather_into_tensor(t_recv, t_send, group=group)
    ctx.check(ctx.lib.eo_stats_collective_end(ctx.handle, world))


def allreduce_stats(ctx, stats: dict | None = None, group=None) -> dict:
    """Global statistics of all ranks of `group` (default: the world group), whatever its backend: NCCL -> on the
    device (`allreduce_stats_device`, then one 1.7 KB read), anything else -> host tensors.  `stats` defaults to the
    ctx's current local record.  Every rank of the group must call it."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("allreduce_stats needs an initialised torch.distributed process group (one rank per GPU)")
    if "nccl" in str(dist.get_backend(group)):
        allreduce_stats_device(ctx, group, stats)
        return ctx.stats_global()
    return allreduce_stats_host(stats if stats is not None else ctx.stats(), group)


def bind_to_gpu_numa(device: int) -> list[int] | None:
    """Pin the calling process to the CPU cores that are local to CUDA device `device` (NVML's ideal CPU affinity, the
    GPU being identified by its PCI bus id), so that the
    page-locked result buffers it allocates afterwards - and the threads that touch them - live on the GPU's own
    NUMA node.  One rank per GPU (external_operator.py:368-370 under MPI) without this lands every rank's pinned
    memory wherever the launcher started it, and the D2H of the tangent then crosses the socket interconnect.
    Returns the core list, or None when NVML / sched_setaffinity are unavailable (nothing is changed then)."""
    import os

    try:
        import ctypes as C

        import pynvml

        from . import _lib

        pynvml.nvmlInit()
        buf = C.create_string_buffer(32)
        if _lib.load().eo_device_pci_bus_id(int(device), buf, 32) != 0:  # `device` is a CUDA ordinal of this process
            return None
        h = pynvml.nvmlDeviceGetHandleByPciBusId(buf.value)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cores = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cores = [c for c in cores if c in allowed]
        if not cores:
            return None
        os.sched_setaffinity(0, cores)
        return cores
    except Exception:
        return None
