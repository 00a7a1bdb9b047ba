"""GPU: the sm_100a tabulation kernels through `Tabulator` / `GeneralTabulator` against golden vectors made by DOLFINx
itself (tests/golden/tab_dolfinx_*.npz, see tests/test_dolfinx_golden_cpu.py).  Skipped while no such file is committed."""

import glob
import os

import numpy as np
import pytest

import dolfinx_external_operator_b200 as eo

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tab_dolfinx_*.npz")))


@pytest.mark.skipif(not GOLDEN, reason="no tests/golden/tab_dolfinx_*.npz (needs fenics-dolfinx 0.10: oracle/gen_golden_dolfinx.py)")
@pytest.mark.parametrize("path", GOLDEN or [None])
def test_kernels_against_dolfinx(ctx, path):
    g = np.load(path, allow_pickle=True)
    bs, gdim = int(g["bs"]), int(g["gdim"])
    simplex = str(g["cell"]) in ("triangle", "tetrahedron") and g["x_dofmap"].shape[1] == gdim + 1
    kw = dict(dofmap=g["dofmap"], x_dofmap=g["x_dofmap"], x=g["x"], phi=g["phi"], dphi=g["dphi"], bs=bs,
              n_dofs=int(g["n_dofs"]), ctx=ctx)
    tab = eo.Tabulator(**kw) if simplex and g["phi"].shape[1] in (3, 4, 6, 10) else eo.GeneralTabulator(dgeo=g["dgeo"], **kw)
    for kind in ("value", "grad", "mandel_strain", "def_grad"):
        if kind not in g.files:
            continue
        want = np.asarray(g[kind]).reshape(g["dofmap"].shape[0], g["phi"].shape[0], -1)
        got = np.asarray(tab.evaluate(kind, g["u"], output="host")).reshape(want.shape)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12 * np.abs(want).max())
