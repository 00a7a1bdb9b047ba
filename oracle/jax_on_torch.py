"""TEST INFRASTRUCTURE - the handful of JAX entry points the reference's
Mohr-Coulomb demo uses (`doc/demo/demo_plasticity_mohr_coulomb.py:282-555`),
implemented over float64 `torch.func`, so that the reference's OWN source can be
executed in a container where JAX is not installable (no network).

Covered: `jax.jacfwd(fn, has_aux=)`, `jax.lax.cond`, `jax.lax.while_loop`,
`jnp.{vdot,sqrt,clip,arcsin,sin,cos,abs,concatenate,logical_and,c_}`,
`jnp.linalg.{norm,solve}`.

Semantics kept from JAX:
  * `jacfwd` = one forward-mode JVP per input basis vector, stacked on the LAST
    axis of the output (jax `jacfwd` convention), aux returned un-differentiated.
  * `lax.while_loop` under forward mode: the predicate sees primal values only,
    tangents are carried through the body -> differentiation THROUGH the loop
    (the behaviour the demo relies on at :555).
  * `lax.cond` with a scalar predicate: only the taken branch contributes value
    and tangent.

Forward mode is done by looping `torch.func.jvp` (no vmap), so data-dependent
Python control flow on primal values stays legal at every nesting level.
"""

from __future__ import annotations

import types

import numpy as np
import torch
from torch.func import jvp as _jvp

_F64 = torch.float64


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=_F64)


def _primal_bool(x) -> bool:
    if isinstance(x, torch.Tensor):
        return bool(x)
    return bool(x)


# ------------------------------------------------------------------ jax.*
def jacfwd(fn, has_aux: bool = False):
    def wrapped(x, *rest):
        x = _t(x)
        n = x.numel()
        cols = []
        aux_out = None
        primal_out = None
        for i in range(n):
            e = torch.zeros(n, dtype=_F64)
            e[i] = 1.0
            e = e.reshape(x.shape)
            if has_aux:

                def with_tensor_aux(z):
                    o, aux = fn(z, *rest)
                    return o, tuple(a if isinstance(a, torch.Tensor) else torch.as_tensor(a) for a in aux)

                out, tang, aux = _jvp(with_tensor_aux, (x,), (e,), has_aux=True)
                aux_out = aux
            else:
                out, tang = _jvp(lambda z: fn(z, *rest), (x,), (e,))
            primal_out = out
            cols.append(tang)
        jac = torch.stack(cols, dim=-1).reshape(tuple(primal_out.shape) + tuple(x.shape))
        if has_aux:
            return jac, aux_out
        return jac

    return wrapped


def _cond(pred, true_fn, false_fn, *operands):
    taken = true_fn if _primal_bool(pred) else false_fn
    out = taken(*operands)
    return out


def _while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while _primal_bool(cond_fun(val)):
        val = body_fun(val)
    return val


lax = types.SimpleNamespace(cond=_cond, while_loop=_while_loop)
jax = types.SimpleNamespace(jacfwd=jacfwd, lax=lax)


# ------------------------------------------------------------------ jax.numpy.*
class _CClass:
    def __getitem__(self, key):
        # only the form used at demo_mc:462: jnp.c_["0,1,-1", vec, scalar]
        spec, *items = key
        assert spec == "0,1,-1"
        return torch.cat([_t(i).reshape(-1) for i in items])


def _vdot(a, b):
    return torch.dot(_t(a).reshape(-1), _t(b).reshape(-1))


def _concatenate(items):
    return torch.cat([_t(i).reshape(-1) for i in items])


def _logical_and(a, b):
    return _primal_bool(a) and _primal_bool(b)


jnp = types.SimpleNamespace(
    vdot=_vdot,
    sqrt=lambda x: torch.sqrt(_t(x)),
    clip=lambda x, lo, hi: torch.clamp(_t(x), lo, hi),
    arcsin=lambda x: torch.arcsin(_t(x)),
    sin=lambda x: torch.sin(_t(x)),
    cos=lambda x: torch.cos(_t(x)),
    abs=lambda x: torch.abs(_t(x)),
    concatenate=_concatenate,
    logical_and=_logical_and,
    c_=_CClass(),
    linalg=types.SimpleNamespace(
        norm=lambda x: torch.linalg.vector_norm(_t(x)),
        solve=lambda a, b: torch.linalg.solve(_t(a), _t(b)),
    ),
)


def numpy_constants_to_torch(ns: dict, names) -> None:
    """The demo mixes NumPy constant matrices with JAX arrays (`dev @ sigma`);
    JAX accepts that, torch does not -> re-bind those module constants as
    float64 tensors after they have been defined by the reference source."""
    for nm in names:
        ns[nm] = _t(ns[nm])
