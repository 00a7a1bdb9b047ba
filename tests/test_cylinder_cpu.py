"""End-to-end known answers of the von Mises demo problem (thick-walled cylinder, demo_plasticity_von_mises.py) on the
NumPy oracle chain tabulation -> return mapping -> residual / tangent forms: the anchors SURVEY.md 8c lists for a chain
whose DOLFINx arithmetic cannot run here - Lame solution in the elastic range, quadratic Newton convergence with the
consistent tangent, plastic zone reaching the outer radius at the analytic collapse load q_lim (demo_vm:542)."""

import numpy as np

from cylinder_util import OracleBackend
import cylinder_driver as twc


def test_boundary_data_of_the_quarter_ring():
    m = twc.quarter_ring_mesh(4, 12)
    # unit pressure on the inner arc: resultant = int_0^{pi/2} R_i (cos t, sin t) dt = (R_i, R_i) up to the chord error
    fx, fy = m["load"][0::2].sum(), m["load"][1::2].sum()
    assert abs(fx - twc.R_I) < 2e-3 and abs(fy - twc.R_I) < 2e-3 and abs(fx - fy) < 1e-12
    xy = m["dof_coords"]
    fixed_y = m["fixed"][m["fixed"] % 2 == 1] // 2
    fixed_x = m["fixed"][m["fixed"] % 2 == 0] // 2
    assert np.allclose(xy[fixed_y, 1], 0.0) and np.allclose(xy[fixed_x, 0], 0.0)
    assert fixed_y.size == 2 * 4 + 1 and fixed_x.size == 2 * 4 + 1
    assert np.allclose(xy[m["probe"] // 2], [twc.R_I, 0.0])


def test_load_displacement_curve_known_answers():
    m = twc.quarter_ring_mesh(5, 16)
    res = twc.solve(m, OracleBackend(m), n_steps=12)
    load, u, pf = res["load"], res["u_probe"], res["plastic_fraction"]
    # elastic range: Lame (P2 on straight-sided cells: discretisation error of a few 1e-4)
    k_el = np.nonzero((pf == 0.0) & (load > 0))[0]
    assert k_el.size >= 2
    for k in k_el:
        assert abs(u[k] - twc.lame_inner_displacement(load[k] * twc.Q_LIM)) < 2e-3 * u[k]  # chord error of 16 edges
        assert res["newton_iterations"][k] == 1  # linear problem: one Newton step
    # plastic range: consistent tangent => quadratic convergence (residual drops faster than squared, up to a constant)
    k_pl = int(np.nonzero((pf > 0.2) & (pf < 1.0))[0][-1])
    h = np.array(res["residual_histories"][k_pl])
    assert 2 <= len(h) - 1 <= 8
    assert h[-1] <= 1e-8 * h[0] and h[-1] / h[-2] < 1e-2 * (h[-2] / h[-3] if len(h) > 2 else 1.0) * 10
    # the plastic zone reaches the outer radius at the analytic collapse load (hardening E_t = E/100 shifts it slightly up)
    assert pf[load < 0.9].max() < 1.0
    assert pf[load > 1.02].min() == 1.0
    # and the response softens by an order of magnitude there
    slope_el = (u[k_el[-1]] - u[k_el[0]]) / (load[k_el[-1]] - load[k_el[0]])
    slope_end = (u[-1] - u[-2]) / (load[-1] - load[-2])
    assert slope_end > 8 * slope_el


def test_lame_solution_mesh_convergence():
    """Elastic step on two meshes: the error against the Lame solution is the O(h^2) chord error of the arcs."""
    err = []
    for n_r, n_t in ((4, 12), (8, 24)):
        m = twc.quarter_ring_mesh(n_r, n_t)
        res = twc.solve(m, OracleBackend(m), n_steps=2, max_load=0.05)
        assert res["plastic_fraction"][1] == 0.0
        exact = twc.lame_inner_displacement(res["load"][1] * twc.Q_LIM)
        err.append(abs(res["u_probe"][1] - exact) / exact)
    assert err[0] < 4e-3 and err[1] < err[0] / 3.5
