"""TEST-ONLY: builds a JitModel's CUDA C++ source with g++ (the dual-number header include/eo_dual.h is
host-compilable), so that the model text and the AD algebra can be checked against the golden vectors on a
machine without a GPU.  The driver below restates what eo_jit_device.cuh does per point (seed, call, extract)."""

from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_DRIVER = r"""
#include <algorithm>
#include <cmath>
#include <cstdint>
#include "eo_dual.h"
using std::max;  // CUDA declares ::max / ::min / ::abs for double; the host build borrows the std ones
using std::min;
using std::abs;
#define __device__
#define __forceinline__ inline
%(source)s

static const int OPS[] = {%(ops)s};
static const int STS[] = {%(sts)s 0};
static const int AUXS[] = {%(auxs)s 0};
enum { NOPS = %(nops)d, NSTS = %(nsts)d, NAUXS = %(nauxs)d, NIN = %(nin)d, NST = %(nst)d, NOUT = %(nout)d, NAUX = %(naux)d,
       DA = %(da)d, DB = %(db)d, NA = %(na)d, NB = %(nb)d, OA = %(oa)d, OB = %(ob)d };

template <class T> struct seed;
template <> struct seed<double> {
  static void set(double& x, double v, int) { x = v; }
  static void out(const double& y, double* D) { D[0] = y; }
  enum { W = 1 };
};
template <> struct seed<eo::dual<NA>> {
  typedef eo::dual<NA> T;
  static void set(T& x, double v, int k) { x.v = v; for (int j = 0; j < NA; ++j) x.d[j] = (k == OA + j) ? 1.0 : 0.0; }
  static void out(const T& y, double* D) { for (int j = 0; j < NA; ++j) D[j] = y.d[j]; }
  enum { W = NA };
};
template <> struct seed<eo::dual<NA, eo::dual<NB>>> {
  typedef eo::dual<NB> V;
  typedef eo::dual<NA, V> T;
  static void set(T& x, double v, int k) {
    x.v.v = v;
    for (int j = 0; j < NB; ++j) x.v.d[j] = (k == OB + j) ? 1.0 : 0.0;
    for (int j = 0; j < NA; ++j) x.d[j] = V((k == OA + j) ? 1.0 : 0.0);
  }
  static void out(const T& y, double* D) { for (int a = 0; a < NA; ++a) for (int b = 0; b < NB; ++b) D[a * NB + b] = y.d[a].d[b]; }
  enum { W = NA * NB };
};

template <class T>
static void run(const double* const* operands, const double* const* state, const double* prm, double* out, double* value,
                double* const* aux, int64_t n) {
  typedef seed<T> S;
  for (int64_t i = 0; i < n; ++i) {
    T x[NIN], y[NOUT], w[NAUX + 1];
    double s[NST + 1];
    int k = 0;
    for (int o = 0; o < NOPS; ++o) for (int c = 0; c < OPS[o]; ++c, ++k) S::set(x[k], operands[o][i * OPS[o] + c], k);
    k = 0;
    for (int o = 0; o < NSTS; ++o) for (int c = 0; c < STS[o]; ++c, ++k) s[k] = state[o][i * STS[o] + c];
    %(entry)s<T>(x, s, prm, y, w);
    for (int o = 0; o < NOUT; ++o) S::out(y[o], out + (i * NOUT + o) * S::W);
    if (value) for (int o = 0; o < NOUT; ++o) value[i * NOUT + o] = eo::value(y[o]);
    k = 0;
    for (int o = 0; o < NAUXS; ++o) for (int c = 0; c < AUXS[o]; ++c, ++k) if (aux[o]) aux[o][i * AUXS[o] + c] = eo::value(w[k]);
  }
}

extern "C" void host_eval(const double* const* operands, const double* const* state, const double* prm, double* out,
                          double* value, double* const* aux, int64_t n) {
  run<%(T)s>(operands, state, prm, out, value, aux, n);
}
"""


def host_eval(model, derivatives, operands, state=()):
    """Evaluate `model` (a JitModel, possibly compile_only) on the HOST for `derivatives`.
    Returns (out, value, [aux...]) as flat float64 arrays."""
    derivatives = tuple(derivatives)
    ops, sts, auxs = model.operand_sizes, model.state_sizes, model.aux_sizes
    order = sum(derivatives)
    idx = [i for i, d in enumerate(derivatives) for _ in range(d)]
    da = idx[0] if order >= 1 else 0
    db = idx[1] if order >= 2 else 0
    na, nb = ops[da], ops[db]
    oa, ob = sum(ops[:da]), sum(ops[:db])
    T = {0: "double", 1: "eo::dual<NA>", 2: "eo::dual<NA, eo::dual<NB>>"}[order]
    text = _DRIVER % dict(
        source=model._src.decode(), entry=model._entry.decode(), ops=",".join(map(str, ops)),
        sts="".join(f"{s}," for s in sts), auxs="".join(f"{s}," for s in auxs), nops=len(ops), nsts=len(sts),
        nauxs=len(auxs), nin=sum(ops), nst=sum(sts), nout=model.out_size, naux=sum(auxs), da=da, db=db, na=na, nb=nb,
        oa=oa, ob=ob, T=T)
    with open(os.path.join(ROOT, "include", "eo_dual.h"), "rb") as f:
        key = hashlib.sha1(text.encode() + f.read()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), "eo_jit_hostcheck")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, f"m{key}.so")
    if not os.path.exists(so):
        src = os.path.join(d, f"m{key}.cpp")
        with open(src, "w") as f:
            f.write(text)
        subprocess.check_call(["g++", "-O2", "-fPIC", "-std=c++17", "-ffp-contract=off", "-shared", "-I",
                               os.path.join(ROOT, "include"), "-o", so, src, "-lm"])
    lib = C.CDLL(so)
    operands = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in operands]
    state = [np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in state]
    n = operands[0].size // ops[0]
    width = model.out_size * (na if order >= 1 else 1) * (nb if order >= 2 else 1)
    out = np.empty(n * width)
    value = np.empty(n * model.out_size)
    aux = [np.empty(n * s) for s in auxs]
    pp = lambda arrs: (C.c_void_p * max(1, len(arrs)))(*[a.ctypes.data for a in arrs])  # noqa: E731
    prm = model.params
    lib.host_eval.restype = None
    lib.host_eval(pp(operands), pp(state), C.c_void_p(prm.ctypes.data if prm.size else None), C.c_void_p(out.ctypes.data),
                  C.c_void_p(value.ctypes.data), pp(aux), C.c_int64(n))
    return out, value, aux
