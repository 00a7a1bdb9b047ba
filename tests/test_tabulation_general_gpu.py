"""GPU parity of the general tabulation kernel (csrc/tabg.cu, through eo_gtab_* / GeneralTabulator) against the
NumPy oracle: quadrilaterals, hexahedra, simplices, cell lists and (cell, local facet) entities, every operand kind."""

import numpy as np
import pytest

import dolfinx_external_operator_b200 as eo
from oracle import tabulation as ot
from tab_util import general_case

pytestmark = pytest.mark.gpu
KINDS = {"value": ot.VALUE, "grad": ot.GRAD, "mandel_strain": ot.MANDEL_STRAIN, "def_grad": ot.DEF_GRAD}


def _tab(ctx, m):
    return eo.GeneralTabulator(dofmap=m["dofmap"], x_dofmap=m["x_dofmap"], x=m["x"], phi=m["phi"], dphi=m["dphi"],
                               dgeo=m["dgeo"], bs=m["bs"], n_dofs=m["n_dofs"], ctx=ctx)


def _kinds(m):
    k = ["value", "grad"]
    if m["bs"] == m["gdim"]:
        k.append("def_grad")
    if m["bs"] == 2 and m["gdim"] == 2:
        k.append("mandel_strain")
    return k


@pytest.mark.parametrize("cell,degree,bs", [("triangle", 2, 2), ("quadrilateral", 1, 1), ("quadrilateral", 2, 2),
                                            ("hexahedron", 1, 3), ("hexahedron", 2, 1), ("hexahedron", 2, 3),
                                            ("tetrahedron", 1, 3)])
def test_cells_against_oracle(ctx, cell, degree, bs):
    m = general_case(cell, degree, bs)
    t = _tab(ctx, m)
    rng = np.random.default_rng(2)
    u = rng.normal(size=bs * m["n_dofs"])
    n_cells = m["dofmap"].shape[0]
    some = rng.permutation(n_cells)[: max(1, n_cells // 3)].astype(np.int32)
    for kind in _kinds(m):
        for ent in (None, some):
            ref = ot.tabulate_general(KINDS[kind], u, m["dofmap"], bs, m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dgeo"],
                                      entities=ent)
            got = t.evaluate(kind, u, entities=ent, output="host")
            assert got.shape[:2] == ref.shape[:2]
            np.testing.assert_allclose(got.reshape(ref.shape), ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
            dev = t.evaluate(kind, ctx.to_device(u), entities=ent)  # device in, device out
            assert np.array_equal(dev.to_host().reshape(-1), got.reshape(-1))


@pytest.mark.parametrize("cell,degree,bs", [("triangle", 1, 1), ("triangle", 2, 2), ("quadrilateral", 2, 1),
                                            ("hexahedron", 1, 3), ("tetrahedron", 1, 1)])
def test_facet_entities_against_oracle(ctx, cell, degree, bs):
    m = general_case(cell, degree, bs, facets=True)
    t = _tab(ctx, m)
    rng = np.random.default_rng(3)
    u = rng.normal(size=bs * m["n_dofs"])
    n_f = m["phi"].shape[0]
    ent = np.stack([rng.integers(0, m["dofmap"].shape[0], 57), rng.integers(0, n_f, 57)], axis=1).astype(np.int32)
    for kind in _kinds(m):
        ref = ot.tabulate_general(KINDS[kind], u, m["dofmap"], bs, m["x"], m["x_dofmap"], m["phi"], m["dphi"], m["dgeo"],
                                  entities=ent)
        got = t.evaluate(kind, u, entities=ent, output="host")
        np.testing.assert_allclose(got.reshape(ref.shape), ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    # the reference test's own field u = x + y on boundary facets (test_codim_external_operator.py:72,98)
    if bs == 1:
        uf = m["dof_coords"][:, : m["gdim"]].sum(axis=1)
        val = t.evaluate("value", uf, entities=ent, output="host")
        np.testing.assert_allclose(val, m["xq"][ent[:, 1], ent[:, 0]].sum(axis=2), rtol=1e-12, atol=1e-12)


def test_evaluate_operands_with_facet_entities(ctx):
    """`evaluate_operands(ops, entities=parent_to_sub)` with (cell, local facet) pairs, as the reference calls it."""
    from types import SimpleNamespace

    m = general_case("triangle", 1, 1, facets=True)
    t = _tab(ctx, m)
    u = m["dof_coords"][:, :2].sum(axis=1)
    t.coefficient = u
    t.register("u", "value")
    op = SimpleNamespace(ufl_operands=["u"], b200_tabulator=t)
    ent = np.array([[0, 0], [3, 2], [5, 1]], dtype=np.int32)
    out = eo.evaluate_operands([op], entities=ent)["u"]
    assert out.shape == (3, 2)
    np.testing.assert_allclose(out.to_host(), m["xq"][ent[:, 1], ent[:, 0]].sum(axis=2), rtol=1e-13)


def test_bad_arguments(ctx):
    m = general_case("triangle", 1, 1, facets=True)
    t = _tab(ctx, m)
    u = np.zeros(m["n_dofs"])
    with pytest.raises(eo.EOError):  # facet tables need (cell, facet) pairs
        t.evaluate("value", u, entities=np.array([0, 1], dtype=np.int32))
    with pytest.raises(eo.EOError):
        t.evaluate("value", u, entities=np.array([[0, 3]], dtype=np.int32))  # local facet 3 of a triangle
    with pytest.raises(ValueError):
        t.evaluate("mandel_strain", u, entities=np.array([[0, 1]], dtype=np.int32))
    assert t.evaluate("value", u, entities=np.zeros((0, 2), dtype=np.int32), output="host").size == 0
