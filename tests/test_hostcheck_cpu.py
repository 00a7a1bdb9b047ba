"""CPU: the host/device "core" headers of the CUDA kernels (csrc/*_core.cuh), compiled with g++ by
tests/hostcheck/Makefile, against the oracle.  This checks the kernels' ALGEBRA (e.g. the Taylor-jet
restatement of the reference's nested forward-mode AD for Mohr-Coulomb) on a machine without a GPU; the
GPU parity tests proper are tests/test_*_gpu.py.  The harness library is test-only: the product package
never loads it."""

import ctypes as C
import dataclasses
import os
import subprocess

import numpy as np
import pytest

from oracle import constitutive as oc
from oracle import inputs, native
from mc_util import check_mc, check_mc_exact

HERE = os.path.dirname(os.path.abspath(__file__))
MC_PRM = oc.MohrCoulombParams()


@pytest.fixture(scope="module")
def hc():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "hostcheck")], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(HERE, "hostcheck", "libhostcheck.so"))


def _mc(hc, deps, sn, prm):
    n = deps.shape[0]
    deps, sn = np.ascontiguousarray(deps), np.ascontiguousarray(sn)
    q = native._mc_prm(prm)
    out = {"C_tang": np.empty((n, 4, 4)), "sigma": np.empty((n, 4)), "niter": np.empty(n, dtype=np.int32),
           "yielding": np.empty(n), "norm_res": np.empty(n), "dlambda": np.empty(n)}
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    hc.hostcheck_mc(C.byref(q), p(deps), p(sn), p(out["C_tang"]), p(out["sigma"]), p(out["niter"]),
                    p(out["yielding"]), p(out["norm_res"]), p(out["dlambda"]), C.c_int64(n))
    return out


@pytest.mark.parametrize("name", ["mc_path_10x9.npz", "mc_rand_seed0_n96.npz", "mc_rand_seed1_n2048.npz"])
def test_mc_core_against_reference_golden(hc, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    o = _mc(hc, g["deps"], g["sigma_n"], MC_PRM)
    check_mc(o, g, g["deps"], g["sigma_n"], MC_PRM)
    check_mc_exact(o, g["deps"], g["sigma_n"], MC_PRM)  # flat 1e-10 at every point, corners included


@pytest.mark.parametrize("psi_deg", [30, 10])
def test_mc_core_against_oracle(hc, psi_deg):
    prm = dataclasses.replace(MC_PRM, psi=psi_deg * np.pi / 180)
    step = lambda d, s: native.mc_stress(d, s, prm, parallel=True)[0]  # noqa: E731
    d, s = inputs.mc_batch(20_000, seed=3, stepper=step)
    o = _mc(hc, d, s, prm)
    check_mc(o, native.mc_return_mapping(d, s, prm, parallel=True), d, s, prm)
    check_mc_exact(o, d, s, prm)


def test_mc_core_edge_semantics(hc):
    d = np.array([[0.0, 0, 0, 0], [0.0, 0, 0, 0], [1e-3, -1e-3, 0, 0]])
    s = np.array([[-1.0, -1.2, -0.9, 0.1], [-1.0, -1, -1, 0], [-1.0, -1, -1, 0]])
    o, r = _mc(hc, d, s, MC_PRM), native.mc_return_mapping(d, s, MC_PRM)
    assert list(o["niter"]) == [0, 0, 0] == list(r["niter"])
    assert np.array_equal(o["sigma"], r["sigma"])
    assert np.array_equal(np.isnan(o["C_tang"]), np.isnan(r["C_tang"]))
