"""CPU: the Isihara kernel core (isihara_core.cuh, host build) against the golden produced by the reference's
own torch code and weights (oracle/gen_golden.py gen_isihara), and against the reference executed live when
/root/reference is present.  For this model the oracle IS the reference code: there is no separate restatement."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from dolfinx_external_operator_b200.isihara import preprocess_state_dict
from isi_util import check, load_golden
from oracle import inputs, ref_exec

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hc():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "hostcheck")], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(HERE, "hostcheck", "libhostcheck.so"))


def _eval(hc, w, F):
    F = np.ascontiguousarray(F, dtype=np.float64).reshape(-1, 4)
    n = F.shape[0]
    dP, P = np.empty((n, 4, 4)), np.empty((n, 4))
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    hc.hostcheck_isihara(C.byref(w), p(F), p(dP), p(P), C.c_int64(n))
    return dP, P


def test_core_against_reference_golden(hc):
    g, sd = load_golden()
    w = preprocess_state_dict(sd)
    for i in range(4):
        w.H[i] = g["H_flat"][i]
    dP, P = _eval(hc, w, g["F"])
    check(dP, P, g)
    # known answers: P(F = I) = 0 up to the float32 correction (demo_hyperelasticity.py:362-381); the tangent
    # is a Hessian, hence symmetric (the reference's is only to 1e-8)
    assert np.abs(P[0]).max() < 1e-6
    assert np.abs(dP - dP.transpose(0, 2, 1)).max() < 1e-12 * np.abs(dP).max()


def test_preprocess_rejects_other_architectures():
    _, sd = load_golden()
    bad = dict(sd)
    bad["layers.2.weights"] = np.zeros((32, 64), dtype=np.float32)
    with pytest.raises(ValueError):
        preprocess_state_dict(bad)


@pytest.mark.skipif(not ref_exec.reference_available(), reason="/root/reference absent (GPU box)")
def test_core_against_live_reference(hc):
    import torch

    ns = ref_exec.load_isihara()
    w = preprocess_state_dict(torch.load(os.path.join(ref_exec.DEMO_DIR, "Isihara_noise=high.pth")))
    for i in range(4):
        w.H[i] = float(ns["H_flat"][i])
    F = inputs.isihara_batch(257, seed=5)
    rdP, rP = ns["dP_dF_impl"](F)
    dP, P = _eval(hc, w, F)
    check(dP, P, {"dP": rdP.reshape(-1, 4, 4), "P": rP.reshape(-1, 4)})
