"""TEST INFRASTRUCTURE - CPU restatement (torch, functional form) of the reference's Isihara ICNN stress / tangent
evaluation, doc/demo/demo_hyperelasticity.py: network :242-307, stress correction :362-381, batched AD entry point
:429-456.  Only tests/ and bench.py's CPU-baseline legs may import this; the product never does.

Pinned: tests/test_isihara_cpu.py checks it against tests/golden/isihara_seed0_n2049.npz, which the reference's own
module produced (oracle/gen_golden.py gen_isihara) - same float32 network arithmetic, so agreement is at float32
rounding level."""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as tf


def _sd_tensors(state_dict) -> dict:
    return {k: torch.as_tensor(np.asarray(v), dtype=torch.float32) for k, v in state_dict.items()}


def energy(f: torch.Tensor, sd: dict) -> torch.Tensor:
    """W_NN(F) for ONE flat deformation gradient f = [F11, F12, F21, F22] (:263-300)."""
    F11, F12, F21, F22 = f[0], f[1], f[2], f[3]
    C11, C22 = F11**2 + F21**2, F12**2 + F22**2  # :269-272
    C12 = F11 * F12 + F21 * F22
    I1 = C11 + C22 + 1.0  # :275-277
    I2 = C11 + C22 - C12 * C12 + C11 * C22
    I3 = C11 * C22 - C12 * C12
    K1 = I1 * torch.pow(I3, -1 / 3) - 3.0  # :280-283
    K2 = I2 * torch.pow(I3, -2 / 3) - 3.0
    K3 = (torch.sqrt(I3) - 1) ** 2
    x = torch.stack((K1, K2, K3)).float()  # :286 the network runs in float32
    z = sd["layers.0.weight"] @ x + sd["layers.0.bias"]  # :288-289, no activation
    for layer in ("1", "2"):  # :290-295
        z = tf.softplus(sd[f"layers.{layer}.weights"]) @ z + sd[f"skip_layers.{layer}.weight"] @ x + sd[f"skip_layers.{layer}.bias"]
        z = tf.softplus(z)
        z = 1 / 12.0 * torch.square(z)
    return (tf.softplus(sd["layers.3.weights"]) @ z + tf.softplus(sd["skip_layers.3.weights"]) @ x).squeeze()  # :299


def stress_correction(sd: dict) -> torch.Tensor:
    """H_flat = -dW_NN/dF at F = I, evaluated in float32 like the reference (:362-367)."""
    F0 = torch.tensor([1.0, 0.0, 0.0, 1.0], dtype=torch.float32)
    return -torch.func.grad(lambda f: energy(f, sd))(F0).detach()


def dP_dF(Fvals: np.ndarray, state_dict, H_flat=None):
    """(dP [n,4,4], P [n,4]) float64 - `dP_dF_impl` (:451-456): vmap(jacfwd(P, has_aux)) with P = grad W_NN + F @ H."""
    sd = _sd_tensors(state_dict)
    h = stress_correction(sd) if H_flat is None else torch.as_tensor(np.asarray(H_flat), dtype=torch.float32)
    z = torch.zeros((), dtype=h.dtype)
    H = torch.stack([torch.stack([h[0], h[1], z, z]), torch.stack([h[2], h[3], z, z]), torch.stack([z, z, h[0], h[1]]),
                     torch.stack([z, z, h[2], h[3]])])  # :369-378

    def stress(f):
        P = torch.func.grad(lambda g: energy(g, sd))(f) + f @ H.to(f)  # :436-441
        return P, P

    F = torch.from_numpy(np.ascontiguousarray(Fvals, dtype=np.float64)).reshape(-1, 4)
    dP, P = torch.func.vmap(torch.func.jacfwd(stress, has_aux=True))(F)
    return dP.detach().numpy(), P.detach().numpy()
