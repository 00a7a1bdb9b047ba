#!/usr/bin/env python
"""The reference's von Mises demo (doc/demo/demo_plasticity_von_mises.py) with the constitutive update, the residual and
the tangent matrix evaluated on the B200 (QuadratureForms); sparse LU on the host.  Needs a GPU - there is no fallback.

    python examples/thick_walled_cylinder.py [n_r n_theta]
"""

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import cylinder_driver as twc  # noqa: E402

if __name__ == "__main__":
    n_r, n_t = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (20, 64)
    mesh = twc.quarter_ring_mesh(n_r, n_t)
    res = twc.solve(mesh, twc.GpuBackend(mesh), n_steps=20, verbose=True)
    k = 3
    print(f"Lame check at q/q_lim = {res['load'][k]:.3f}: u_x(R_i,0) = {res['u_probe'][k]:.6e}, analytic "
          f"{twc.lame_inner_displacement(res['load'][k] * twc.Q_LIM):.6e}")
