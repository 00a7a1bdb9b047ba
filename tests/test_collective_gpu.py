"""The one collective of the path: eo_allreduce_stats over a caller-owned ncclComm_t (C ABI), and
parallel.allreduce_stats_device / the models' global_stats() over a torch.distributed NCCL group.
Single-GPU check with a one-rank communicator (ncclCommInitAll); the N > 1 check spawns one rank per GPU with torchrun
(tests/collective_worker.py; skipped on a box with fewer than 2 GPUs); host logic: world-size-2 gloo test in
test_api_cpu.py.  bench.py --gpus N asserts the same identity (global == sum of locals) on every run."""

import ctypes as C
import os
import subprocess
import sys

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
        for _ in range(3):  # the local record is never modified: repeating the collective changes nothing
            ctx.check(ctx.lib.eo_allreduce_stats(ctx.handle, comm))
        ctx.sync()
        assert ctx.stats()["n_points"] == n
        after = ctx.stats_global()
        assert after["n_points"] == before["n_points"] == n
        assert after["n_plastic"] == before["n_plastic"] == int((np.asarray(dp) > 0).sum())
        assert np.array_equal(after["niter_hist"], before["niter_hist"])
        assert after["f_max"] == before["f_max"] and after["res_max"] == before["res_max"]
    finally:
        nccl.ncclCommDestroy(comm)
    with pytest.raises(eo.EOError):
        ctx.check(ctx.lib.eo_allreduce_stats(ctx.handle, None))


def test_stats_collective_over_nccl_ranks():
    """N >= 2 GPUs: reduced counts == sum of the ranks' local counts, through the C ABI (raw ncclComm_t), through the
    torch.distributed group after every one of several accumulating evaluations, and through the public model API."""
    import torch

    n_gpu = torch.cuda.device_count()
    if n_gpu < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    world = min(n_gpu, 4)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), os.path.join(root, "tests", "collective_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    assert res.stdout.count(" ok") == world
