"""Worker of tests/test_collective_gpu.py::test_stats_collective_over_nccl_ranks - run under torchrun, one rank per GPU.

Every rank evaluates a DIFFERENT seeded batch (different size, different plastic fraction), then the statistics
collective is exercised three ways and each result is compared with the sum / max of the local records, which are
gathered independently as Python objects:
  1. `eo_allreduce_stats` through the C ABI over a raw `ncclComm_t` created here with ncclCommInitRank;
  2. `parallel.allreduce_stats_device` on the torch.distributed NCCL group, issued after EVERY one of several
     accumulating evaluations (the pattern that multiplied the counts when the reduction was in place);
  3. the public `MohrCoulomb.summary(reduce=True)` / `VonMises.global_stats()`.
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dolfinx_external_operator_b200 as eo  # noqa: E402
from dolfinx_external_operator_b200 import parallel as par  # noqa: E402
from dolfinx_external_operator_b200 import synthetic as syn  # noqa: E402


def gather_locals(st):
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in st.items()})
    return out


def check(glob, locs, what):
    for k in ("n_points", "n_plastic", "n_nonconverged", "n_nonfinite"):
        assert glob[k] == sum(l[k] for l in locs), (what, k, glob[k], [l[k] for l in locs])
    assert np.array_equal(glob["niter_hist"], np.sum([np.asarray(l["niter_hist"]) for l in locs], axis=0)), what
    for k in ("niter_max", "f_max", "res_max"):
        v = [l[k] for l in locs if l[k] == l[k]]
        assert glob[k] == (max(v) if v else glob[k]), (what, k, glob[k], [l[k] for l in locs])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ctx = eo.Context(local)

    # ---- per-rank batches
    n = 40_000 + 7_001 * rank
    deps, sn, p = syn.vm_batch(n, seed=rank)
    deps *= 0.5 + rank  # different plastic fractions per rank
    vm = eo.VonMises(ctx=ctx)
    vm.set_history(sn, p)
    _, _, dp = vm((1,))(deps)
    assert vm.local_stats()["n_points"] == n and vm.local_stats()["n_plastic"] == int((np.asarray(dp) > 0).sum())

    # 1. raw ncclComm_t through the C ABI
    nccl = C.CDLL("libnccl.so.2")
    uid = (C.c_char * 128)()
    if rank == 0:
        assert nccl.ncclGetUniqueId(uid) == 0
    box = [bytes(uid)]
    dist.broadcast_object_list(box, src=0)

    class UniqueId(C.Structure):
        _fields_ = [("internal", C.c_char * 128)]

    u = UniqueId()
    C.memmove(C.byref(u), box[0], 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), world, u, rank) == 0
    ctx.stats_reset()
    vm.eval_device(ctx.to_device(deps.reshape(-1)), ctx.empty((16 * n,)))
    for rep in range(3):  # repeated calls on an unchanged local record give the same global record
        ctx.check(ctx.lib.eo_allreduce_stats(ctx.handle, comm))
        g = ctx.stats_global()
        loc = ctx.stats()
        assert loc["n_points"] == n, "the collective must not modify the local record"
        check(g, gather_locals(loc), f"C ABI rep {rep}")
    assert g["n_points"] == sum(40_000 + 7_001 * r for r in range(world))
    nccl.ncclCommDestroy(comm)

    # 2. torch.distributed group, after every one of K accumulating evaluations
    d_deps, d_Ct = ctx.to_device(deps.reshape(-1)), ctx.empty((16 * n,))
    ctx.stats_reset()
    K = 6
    for k in range(K):
        vm.eval_device(d_deps, d_Ct)
        par.allreduce_stats_device(ctx)
    g = ctx.stats_global()
    locs = gather_locals(ctx.stats())
    check(g, locs, "torch group, accumulating")
    assert g["n_points"] == K * sum(40_000 + 7_001 * r for r in range(world)), g["n_points"]
    frac = g["n_plastic"] / g["n_points"]
    assert 0.0 < frac < 1.0

    # 3. public API: Mohr-Coulomb summary over all ranks
    mc = eo.MohrCoulomb(ctx=ctx)
    nm = 3_000 + 500 * rank
    dm, sm = syn.mc_batch(nm, seed=rank, stepper=mc.stress_update)
    mc.set_history(sm)
    mc((1,))(dm.reshape(-1, 1, 4))
    s_loc = mc.summary()
    s_glob = mc.summary(reduce=True)
    locs = gather_locals(mc.local_stats())
    assert s_glob["n_points"] == sum(l["n_points"] for l in locs) == sum(3_000 + 500 * r for r in range(world))
    assert s_glob["n_plastic"] == sum(l["n_plastic"] for l in locs)
    assert s_glob["max_f"] == max(l["f_max"] for l in locs) and s_glob["max_residual"] == max(l["res_max"] for l in locs)
    assert int(np.sum(s_glob["counts"])) == s_glob["n_points"] and s_loc["n_points"] == nm
    gv = vm.global_stats()
    assert gv["n_points"] == sum(40_000 + 7_001 * r for r in range(world))
    ctx.sync()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok: global plastic fraction {frac:.4f}", flush=True)


if __name__ == "__main__":
    main()
