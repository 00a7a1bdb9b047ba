"""CPU: the tabulation / forms oracle (and with it the kernels, which the GPU tier checks against the same oracle) against
golden vectors made by DOLFINx itself - tests/golden/tab_dolfinx_*.npz, produced by oracle/gen_golden_dolfinx.py on a
machine that has fenics-dolfinx 0.10 (the build container has not: DESIGN.md section 4).  Skipped while no such file is
committed; the day one is, SURVEY.md 8 rows A1 / f1 / f2 are pinned to `fem.Expression.eval` and the DOLFINx assemblers."""

import glob
import os

import numpy as np
import pytest

from oracle import forms as of
from oracle import tabulation as ot

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tab_dolfinx_*.npz")))
KIND = {"value": ot.VALUE, "grad": ot.GRAD, "mandel_strain": ot.MANDEL_STRAIN, "def_grad": ot.DEF_GRAD}


def test_generator_script_is_present_and_documented():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(here, "oracle", "gen_golden_dolfinx.py")).read()
    assert "fem.Expression" in src and "tab_dolfinx_" in src


@pytest.mark.skipif(not GOLDEN, reason="no tests/golden/tab_dolfinx_*.npz (needs fenics-dolfinx 0.10: oracle/gen_golden_dolfinx.py)")
@pytest.mark.parametrize("path", GOLDEN or [None])
def test_oracle_against_dolfinx(path):
    g = np.load(path, allow_pickle=True)
    bs, gdim = int(g["bs"]), int(g["gdim"])
    simplex = str(g["cell"]) in ("triangle", "tetrahedron") and g["x_dofmap"].shape[1] == gdim + 1
    phi, dphi, dgeo = g["phi"], g["dphi"], g["dgeo"]
    for kind, kid in KIND.items():
        if kind not in g.files:
            continue
        want = np.asarray(g[kind]).reshape(g["dofmap"].shape[0], phi.shape[0], -1)
        if simplex:
            got = ot.tabulate(kid, g["u"], g["dofmap"], bs, g["x"], g["x_dofmap"], phi, dphi, dgeo[:, 0, :])
        else:
            got = ot.tabulate_general(kid, g["u"], g["dofmap"], bs, g["x"], g["x_dofmap"], phi[None], dphi[None], dgeo[None])
        np.testing.assert_allclose(got.reshape(want.shape), want, rtol=1e-12, atol=1e-12 * np.abs(want).max())
    if simplex and "b_grad" in g.files:
        nc, nq = g["dofmap"].shape[0], phi.shape[0]
        geo = (g["x"], g["x_dofmap"], phi, dphi, dgeo[:, 0, :])
        b = of.assemble_vector(ot.GRAD, g["s"].reshape(nc, nq, -1), g["weights"], g["dofmap"], bs, int(g["n_dofs"]), *geo)
        np.testing.assert_allclose(b, g["b_grad"], rtol=0, atol=1e-12 * np.abs(g["b_grad"]).max())
        y = of.apply_action(ot.GRAD, ot.GRAD, g["D"].reshape(nc, nq, -1), g["xvec"], g["weights"], g["dofmap"], bs,
                            int(g["n_dofs"]), *geo)
        np.testing.assert_allclose(y, g["y_grad_grad"], rtol=0, atol=1e-11 * np.abs(g["y_grad_grad"]).max())
