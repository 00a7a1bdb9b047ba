"""CPU: rounding audit of the Mohr-Coulomb parity tolerance (tests/mc_util.py).

The reference's float64 program is evaluated once more in x87 extended precision (oracle/csrc/mc_oracle.cpp built with
real = long double) and that value is taken as exact.  Checked here, on the 2048-point golden made by the reference's
own source and on a seeded batch:
  * the extended evaluation reproduces the reference's iteration counts and flags;
  * the reference's own float64 results (golden = torch-executed source, oracle = C++ restatement) are within
    C_REF eps / w2 of exact, and DO exceed the flat 1e-10 at hexagon corners - the tolerance rule of check_mc is the
    reference's noise, not slack for the kernel;
  * the kernel arithmetic (mc_core.cuh built for the host) is within 1e-10 of exact at every point."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mc_util import C_REF, EPS, check_mc_exact, lode_w2
from oracle import constitutive as oc
from oracle import inputs, native

HERE = os.path.dirname(os.path.abspath(__file__))
PRM = oc.MohrCoulombParams()


@pytest.fixture(scope="module")
def hc():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "hostcheck")], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(HERE, "hostcheck", "libhostcheck.so"))


def _rel(a, b, key, n):
    A, B = np.asarray(a[key]).reshape(n, -1), np.asarray(b[key]).reshape(n, -1)
    return np.nanmax(np.abs(A - B) / (np.abs(B) + np.abs(B[np.isfinite(B)]).max()), axis=1)


def test_reference_float64_noise_follows_eps_over_w2(golden_dir):
    g = np.load(os.path.join(golden_dir, "mc_rand_seed1_n2048.npz"))
    d, s = g["deps"], g["sigma_n"]
    ex = native.mc_return_mapping(d, s, PRM, parallel=True, extended=True)
    o64 = native.mc_return_mapping(d, s, PRM, parallel=True)
    assert np.array_equal(ex["niter"], g["niter"]) and np.array_equal(ex["niter"], o64["niter"])
    w2 = lode_w2(ex["sigma"], d, s, PRM)
    pl = ex["yielding"] > 0
    assert pl.sum() > 600
    for ref in (g, o64):
        for key in ("C_tang", "sigma", "dlambda"):
            e = np.where(pl, _rel(ref, ex, key, d.shape[0]), 0.0)
            assert (e <= 1e-13 + C_REF * EPS / w2).all(), (key, (e * w2 / EPS).max())


def test_reference_exceeds_flat_tolerance_at_corners_kernel_does_not(hc):
    from test_hostcheck_cpu import _mc

    step = lambda d_, s_: native.mc_stress(d_, s_, PRM, parallel=True)[0]  # noqa: E731
    d, s = inputs.mc_batch(200_000, seed=7, stepper=step)
    ex = native.mc_return_mapping(d, s, PRM, parallel=True, extended=True)
    o64 = native.mc_return_mapping(d, s, PRM, parallel=True)
    assert np.array_equal(ex["niter"], o64["niter"])
    w2 = lode_w2(ex["sigma"], d, s, PRM)
    pl = ex["yielding"] > 0
    e_ref = np.where(pl, _rel(o64, ex, "C_tang", d.shape[0]), 0.0)
    assert e_ref.max() > 1e-10                               # the reference's own float64 noise at the corners
    assert (e_ref <= 1e-13 + C_REF * EPS / w2).all()         # ... and its measured bound
    assert ((C_REF * EPS / w2 > 1e-10) & pl).mean() < 6e-3   # how many points that concerns
    worst = check_mc_exact(_mc(hc, d, s, PRM), d, s, PRM)    # the kernel arithmetic: flat 1e-10, no exception
    assert max(worst.values()) < 2e-11, worst
