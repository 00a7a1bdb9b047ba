#!/usr/bin/env python
"""The reference's Mohr-Coulomb demo (doc/demo/demo_plasticity_mohr_coulomb.py, slope stability) with the constitutive
update, the residual and the tangent matrix evaluated on the B200 (QuadratureForms); sparse LU on the host.  Needs a GPU.

    python examples/slope_stability.py [nx ny]
"""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np  # noqa: E402

import slope_driver as ss  # noqa: E402

if __name__ == "__main__":
    nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (25, 25)
    mesh = ss.slope_mesh(nx, ny)
    steps = np.concatenate([np.linspace(2, 22.9, 50), [22.96, 22.99], np.linspace(23.2, 27, 20)])
    res = ss.solve(mesh, ss.GpuBackend(mesh), load_steps=steps, verbose=True)
    k = res["n_converged"]
    print(f"Slope stability factor: {steps[k - 1] * ss.H / ss.C_COHESION:.3f} (limit analysis: {ss.L_LIM})")
