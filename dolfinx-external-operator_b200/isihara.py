"""Isihara ICNN hyperelastic model: `external_function` factory and torch custom op.

replaces: `compute_stress_local`, `vectorized_stress_and_tangent`, `dP_dF_impl`, `P_external`,
doc/demo/demo_hyperelasticity.py:429-502 (network :242-307, weights `Isihara_noise=high.pth`, corrections
:362-381).  The callable returns `(dP.reshape(-1), P.reshape(-1))` like :451-456.

Two entry points, as north_star prescribes for this (PyTorch-based) demo:
  * `Isihara((1,))(F)` - the NumPy protocol of the other models (ctypes C ABI);
  * `torch.ops.eo.isihara_dP_dF(F) -> (dP, P)` - a torch custom op on CUDA tensors (zero copy: the kernel reads
    and writes the tensors' storage; any DLPack / `__cuda_array_interface__` producer works through
    `torch.from_dlpack`).  Registered by `register_torch_op(model)`.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import IsiharaWeights
from .constitutive import _as_input, _ModelBase
from .context import Context, DeviceArray, _ptr


def _softplus64(w):
    w = np.asarray(w, dtype=np.float64)
    return np.where(w > 20.0, w, np.log1p(np.exp(np.minimum(w, 20.0))))  # torch softplus, threshold 20


def preprocess_state_dict(sd) -> IsiharaWeights:
    """State dict of the reference's ICNN (n_hidden = [64, 64, 64], :304) -> eo_isihara_weights.
    softplus of the convex weights (:238) is applied once, in float64; layer 1 is collapsed onto the affine
    layer 0 (z0 = L0 x + b0 has no activation, :289); everything is then rounded to float32."""
    g = lambda k: np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k], dtype=np.float64)  # noqa: E731
    L0w, L0b = g("layers.0.weight"), g("layers.0.bias")
    W1, W2, W3 = _softplus64(g("layers.1.weights")), _softplus64(g("layers.2.weights")), _softplus64(g("layers.3.weights"))
    S1w, S1b = g("skip_layers.1.weight"), g("skip_layers.1.bias")
    S2w, S2b = g("skip_layers.2.weight"), g("skip_layers.2.bias")
    S3 = _softplus64(g("skip_layers.3.weights"))
    if L0w.shape != (64, 3) or W1.shape != (64, 64) or W2.shape != (64, 64) or W3.shape != (1, 64):
        raise ValueError("this kernel is built for the reference's ICNN: 3 -> 64 -> 64 -> 64 -> 1")
    A1 = np.concatenate([W1 @ L0w + S1w, (W1 @ L0b + S1b)[:, None]], axis=1)
    S2 = np.concatenate([S2w, S2b[:, None]], axis=1)
    w = IsiharaWeights()
    def put(field, a):
        a = np.ascontiguousarray(a, dtype=np.float32)  # kept alive until the copy is done
        assert a.nbytes == C.sizeof(field)
        C.memmove(field, a.ctypes.data, a.nbytes)

    put(w.A1, A1)
    put(w.S2, S2)
    put(w.W2, W2)
    put(w.W2T, W2.T)
    put(w.w3, W3[0])
    put(w.s3, np.concatenate([S3[0], [0.0]]))
    for i in range(4):
        w.H[i] = 0.0
    return w


class Isihara(_ModelBase):
    def __init__(self, state_dict, *, H_flat=None, ctx: Context | None = None):
        """`state_dict`: the reference's `torch.load("Isihara_noise=high.pth")` (or a dict of arrays with the
        same keys).  `H_flat`: the stress correction of :362-367; by default it is evaluated as the reference
        does, -P_NN(F = I), with this library."""
        super().__init__(ctx)
        self.weights = preprocess_state_dict(state_dict)
        h = C.c_void_p()
        c = self.ctx
        c.check(c.lib.eo_isihara_create(c.handle, C.byref(self.weights), C.byref(h)))
        self._h = h
        if H_flat is None:
            F0 = np.array([[1.0, 0.0, 0.0, 1.0]])
            dP0, P0 = np.empty(16), np.empty(4)
            c.check(c.lib.eo_isihara_eval(self._h, _ptr(F0), _ptr(dP0), _ptr(P0), 1))
            H_flat = -P0
        self.H_flat = np.ascontiguousarray(H_flat, dtype=np.float64)
        c.check(c.lib.eo_isihara_set_correction(self._h, self.H_flat.ctypes.data_as(C.POINTER(C.c_double))))

    def close(self):
        if self._h is not None and self.ctx.alive:
            self.ctx.lib.eo_isihara_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, derivatives):  # P_external, :495-499
        if derivatives == (1,):
            return self.dP_dF_impl
        raise NotImplementedError(f"No external function is defined for the requested derivative {derivatives}.")

    def dP_dF_impl(self, Fvals):
        """F (n_cells, n_pts, 2, 2) or any array of n*4 values -> (dP.reshape(-1), P.reshape(-1)), :451-456."""
        F = _as_input(Fvals)
        n = F.size // 4
        dP, P = self._out("dP", 16 * n), self._out("P", 4 * n)
        c = self.ctx
        c.check(c.lib.eo_isihara_eval(self._h, _ptr(F), _ptr(dP), _ptr(P), n))
        c.sync()
        return dP, P

    def eval_device(self, F: DeviceArray, dP: DeviceArray, P: DeviceArray):
        c = self.ctx
        c.check(c.lib.eo_isihara_eval(self._h, F.ptr, dP.ptr, P.ptr, F.size // 4))


_TORCH_MODEL = None


def register_torch_op(model: Isihara):
    """Define `torch.ops.eo.isihara_dP_dF(Tensor F) -> (Tensor dP, Tensor P)` for CUDA float64 tensors,
    backed by `model`: zero copy, launched on torch's CURRENT stream (eo_isihara_eval_on_stream) - asynchronous and
    ordered like any other torch op, no synchronisation on either side."""
    import torch

    global _TORCH_MODEL
    first = _TORCH_MODEL is None
    _TORCH_MODEL = model
    if not first:
        return torch.ops.eo.isihara_dP_dF

    @torch.library.custom_op("eo::isihara_dP_dF", mutates_args=(), schema="(Tensor F) -> (Tensor, Tensor)")
    def isihara_dP_dF(F):
        if not (F.is_cuda and F.dtype == torch.float64):
            raise TypeError("eo::isihara_dP_dF needs a CUDA float64 tensor")
        m = _TORCH_MODEL
        Fc = F.contiguous().reshape(-1, 4)
        n = Fc.shape[0]
        dP = torch.empty((n, 4, 4), dtype=torch.float64, device=F.device)
        P = torch.empty((n, 4), dtype=torch.float64, device=F.device)
        c = m.ctx
        if F.device.index != c.device:
            raise ValueError(f"eo::isihara_dP_dF: tensor on cuda:{F.device.index}, model context on cuda:{c.device}")
        stream = torch.cuda.current_stream(F.device).cuda_stream
        c.check(c.lib.eo_isihara_eval_on_stream(m._h, Fc.data_ptr(), dP.data_ptr(), P.data_ptr(), n, stream))
        return dP, P

    @isihara_dP_dF.register_fake
    def _(F):
        n = F.numel() // 4
        return F.new_empty((n, 4, 4)), F.new_empty((n, 4))

    return torch.ops.eo.isihara_dP_dF
