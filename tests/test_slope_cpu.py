"""End-to-end known answer of the Mohr-Coulomb demo problem (slope stability, demo_plasticity_mohr_coulomb.py:708-770)
on the oracle chain tabulation -> local-Newton return mapping -> residual / tangent forms: the load-displacement curve
of the crest point reaches its plateau at the slope stability factor l_lim = gamma_lim H / c = 6.69 of limit analysis
(Chen; demo_mc:764-765).  A coarse P2 mesh gives a slightly stiffer (upper-bound) answer."""

import numpy as np

import slope_driver as ss
from slope_util import OracleBackend

STEPS = np.concatenate([np.linspace(2, 22.9, 50), [22.96, 22.99], np.linspace(23.2, 27, 20)])  # demo_mc:708-710, extended


def check_collapse(res, lo=0.98, hi=1.05):
    k = res["n_converged"]
    assert 0 < k < len(STEPS), "the extended load path must end in a Newton failure (collapse)"
    l_num = STEPS[k - 1] * ss.H / ss.C_COHESION
    assert lo * ss.L_LIM <= l_num <= hi * ss.L_LIM, l_num
    u, pf = res["u_probe"][:k], res["plastic_fraction"][:k]
    assert np.all(np.diff(u) > 0) and np.all(np.diff(pf) >= -1e-12)
    # plateau: the compliance d u / d gamma at the end is far above the elastic one of the first steps
    slope0 = (u[1] - u[0]) / (STEPS[1] - STEPS[0])
    slope_end = (u[k - 1] - u[k - 2]) / (STEPS[k - 1] - STEPS[k - 2])
    assert slope_end > 20 * slope0
    return l_num


def test_slope_stability_factor_against_limit_analysis():
    m = ss.slope_mesh(12, 10)
    # clamped boundaries (demo_mc:129-145) and the self-weight resultant: sum of f_y = -|domain|
    b = OracleBackend(m)
    f = b.body_force()
    assert abs(f[1::2].sum() + ss.L * ss.H) < 1e-13 and np.abs(f[0::2]).max() == 0.0
    res = ss.solve(m, b, load_steps=STEPS)
    l_num = check_collapse(res)
    assert l_num >= ss.L_LIM  # displacement FE on a coarse mesh: upper bound
    assert res["newton_iterations"][0] <= 2 and res["plastic_fraction"][0] == 0.0  # gamma = 2: elastic
