"""The one collective of the path through the C ABI: eo_allreduce_stats over a caller-owned ncclComm_t.
Single-GPU check with a one-rank communicator (ncclCommInitAll); the N > 1 path is exercised by bench.py --gpus N
(parallel.allreduce_stats_device) and, for the host logic, by the world-size-2 gloo test in test_api_cpu.py."""

import ctypes as C

import numpy as np
import pytest

import dolfinx_external_operator_b200 as eo
from oracle import inputs

pytestmark = pytest.mark.gpu


def _nccl():
    import torch  # noqa: F401  - brings libnccl.so.2 into the process

    for name in ("libnccl.so.2", "libnccl.so"):
        try:
            return C.CDLL(name)
        except OSError:
            continue
    pytest.skip("no libnccl in this environment")


def test_allreduce_stats_over_a_one_rank_communicator(ctx):
    nccl = _nccl()
    comm = C.c_void_p()
    devs = (C.c_int * 1)(ctx.device)
    assert nccl.ncclCommInitAll(C.byref(comm), 1, devs) == 0
    try:
        n = 50_000
        deps, sn, p = inputs.vm_batch(n, seed=3)
        vm = eo.VonMises(ctx=ctx)
        vm.set_history(sn, p)
        ctx.stats_reset()
        _, _, dp = vm((1,))(deps)
        before = ctx.stats()
        ctx.check(ctx.lib.eo_allreduce_stats(ctx.handle, comm))
        ctx.sync()
        after = ctx.stats()
        assert after["n_points"] == before["n_points"] == n
        assert after["n_plastic"] == before["n_plastic"] == int((np.asarray(dp) > 0).sum())
        assert np.array_equal(after["niter_hist"], before["niter_hist"])
        assert after["f_max"] == before["f_max"] and after["res_max"] == before["res_max"]
    finally:
        nccl.ncclCommDestroy(comm)
    with pytest.raises(eo.EOError):
        ctx.check(ctx.lib.eo_allreduce_stats(ctx.handle, None))
