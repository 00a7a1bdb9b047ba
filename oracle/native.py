"""TEST INFRASTRUCTURE - ctypes access to oracle/liboracle.so (the C/C++ restatements).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


class OracleVmParams(C.Structure):
    _fields_ = [("lmbda", C.c_double), ("mu", C.c_double), ("H", C.c_double), ("sigma_0", C.c_double)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def num_threads() -> int:
    return int(load().oracle_num_threads())


def use_all_cores() -> int:
    """Let the OpenMP legs use every host core this process may run on (torchrun sets OMP_NUM_THREADS=1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    load().oracle_set_num_threads(C.c_int(n))
    return num_threads()


def vm_return_mapping(deps, sigma_n, p, prm, parallel: bool = False):
    """C restatement of demo_vm:298-332; arrays in the reference's layout."""
    lib = load()
    deps = np.ascontiguousarray(deps, dtype=np.float64).reshape(-1, 4)
    sigma_n = np.ascontiguousarray(sigma_n, dtype=np.float64).reshape(-1, 4)
    p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1)
    n = p.size
    Ct = np.empty((n, 4, 4))
    sig = np.empty((n, 4))
    dp = np.empty(n)
    q = OracleVmParams(prm.lmbda, prm.mu, prm.H, prm.sigma_0)
    lib.oracle_vm_return_mapping(C.byref(q), _p(deps), _p(sigma_n), _p(p), _p(Ct), _p(sig), _p(dp), C.c_int64(n),
                                 C.c_int(int(parallel)))
    return Ct, sig, dp


_HEAT = {"k": (0, 1), "dk": (1, 1), "q": (2, 2), "dqdT": (3, 2), "dqdsigma": (4, 4)}


def heat(which: str, T, sigma=None, A=1.0, B=1.0, parallel: bool = False):
    lib = load()
    code, width = _HEAT[which]
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(-1)
    s = None if sigma is None else np.ascontiguousarray(sigma, dtype=np.float64).reshape(-1)
    out = np.empty(T.size * width)
    lib.oracle_heat(C.c_int(code), C.c_double(A), C.c_double(B), _p(T), _p(s), _p(out), C.c_int64(T.size),
                    C.c_int(int(parallel)))
    return out


class OracleMcParams(C.Structure):
    _fields_ = [("E", C.c_double), ("nu", C.c_double), ("c", C.c_double), ("phi", C.c_double), ("psi", C.c_double),
                ("theta_T", C.c_double), ("a", C.c_double), ("tol", C.c_double), ("Nitermax", C.c_int32)]


def _mc_prm(prm):
    return OracleMcParams(prm.E, prm.nu, prm.c, prm.phi, prm.psi, prm.theta_T, prm.a, prm.tol, prm.Nitermax)


def mc_return_mapping(deps, sigma_n, prm, parallel: bool = False, extended: bool = False):
    """C++ dual-number restatement of demo_mc:474-555.  Returns dict(C_tang, sigma, niter, yielding, norm_res, dlambda).
    extended=True: the same program evaluated in x87 extended precision (eps 1.1e-19) and rounded to double at the end -
    the "exact" value for rounding audits."""
    lib = load()
    deps = np.ascontiguousarray(deps, dtype=np.float64).reshape(-1, 4)
    sigma_n = np.ascontiguousarray(sigma_n, dtype=np.float64).reshape(-1, 4)
    n = deps.shape[0]
    out = {"C_tang": np.empty((n, 4, 4)), "sigma": np.empty((n, 4)), "niter": np.empty(n, dtype=np.int32),
           "yielding": np.empty(n), "norm_res": np.empty(n), "dlambda": np.empty(n)}
    q = _mc_prm(prm)
    fn = lib.oracle_mc_return_mapping_ld if extended else lib.oracle_mc_return_mapping
    fn(C.byref(q), _p(deps), _p(sigma_n), _p(out["C_tang"]), _p(out["sigma"]), _p(out["niter"]), _p(out["yielding"]),
       _p(out["norm_res"]), _p(out["dlambda"]), C.c_int64(n), C.c_int(int(parallel)))
    return out


def mc_stress(deps, sigma_n, prm, parallel: bool = False):
    lib = load()
    deps = np.ascontiguousarray(deps, dtype=np.float64).reshape(-1, 4)
    sigma_n = np.ascontiguousarray(sigma_n, dtype=np.float64).reshape(-1, 4)
    n = deps.shape[0]
    sig, nit, yl = np.empty((n, 4)), np.empty(n, dtype=np.int32), np.empty(n)
    q = _mc_prm(prm)
    lib.oracle_mc_stress(C.byref(q), _p(deps), _p(sigma_n), _p(sig), _p(nit), _p(yl), C.c_int64(n),
                         C.c_int(int(parallel)))
    return sig, nit, yl


def mc_yield(sigma, prm):
    lib = load()
    sigma = np.ascontiguousarray(sigma, dtype=np.float64).reshape(-1, 4)
    f = np.empty(sigma.shape[0])
    q = _mc_prm(prm)
    lib.oracle_mc_yield(C.byref(q), _p(sigma), _p(f), C.c_int64(sigma.shape[0]))
    return f


def forms_p2_cells(mode: str, m: dict, weights, u, prm=None, sigma_n=None, p=None, C_tang=None):
    """OpenMP C restatement of the von Mises demo's cell loop on a 2-d vector Lagrange mesh dict `m` (dofmap, x_dofmap, x,
    dphi (2, nq, nb), dpsi): mode 'tab' -> strain (n_cells, nq, 4); 'fused' -> (C_tang, sigma, dp); 'step' -> (b, C_tang,
    sigma, dp); 'action' -> y = A u with the given C_tang.  CPU baseline of the corresponding bench legs."""
    lib = load()
    code = {"tab": 0, "fused": 1, "step": 2, "action": 3}[mode]
    dm = np.ascontiguousarray(m["dofmap"], dtype=np.int32)
    xd = np.ascontiguousarray(m["x_dofmap"], dtype=np.int32)
    x = np.ascontiguousarray(m["x"], dtype=np.float64)
    dphi = np.ascontiguousarray(m["dphi"], dtype=np.float64)
    dpsi = np.ascontiguousarray(m["dpsi"], dtype=np.float64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
    nc, nb = dm.shape
    nq = dphi.shape[1]
    n = nc * nq
    if nb > 10 or nq > 16 or dphi.shape[0] != 2:
        raise ValueError("forms_p2_cells: 2-d elements with nb <= 10, nq <= 16")
    strain = np.empty((nc, nq, 4)) if code == 0 else None
    if code in (1, 2):
        C_tang, sigma, dp = np.empty((n, 4, 4)), np.empty((n, 4)), np.empty(n)
        sigma_n = np.ascontiguousarray(sigma_n, dtype=np.float64).reshape(-1, 4)
        p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1)
        q = OracleVmParams(prm.lmbda, prm.mu, prm.H, prm.sigma_0)
    else:
        sigma = dp = None
        q = OracleVmParams(0.0, 0.0, 0.0, 0.0)
        if code == 3:
            C_tang = np.ascontiguousarray(C_tang, dtype=np.float64).reshape(-1, 16)
    vec = np.zeros(2 * m["n_dofs"]) if code >= 2 else None
    lib.oracle_forms_p2_cells(C.c_int(code), C.c_int(nb), C.c_int(nq), _p(dphi), _p(dpsi), _p(w), _p(dm), _p(xd), _p(x), _p(u),
                              C.byref(q), None if sigma_n is None else _p(sigma_n), None if p is None else _p(p),
                              None if strain is None else _p(strain), None if C_tang is None else _p(C_tang),
                              None if sigma is None else _p(sigma), None if dp is None else _p(dp),
                              None if vec is None else _p(vec), C.c_int64(nc))
    return {"tab": strain, "fused": (C_tang, sigma, dp), "step": (vec, C_tang, sigma, dp), "action": vec}[mode]
