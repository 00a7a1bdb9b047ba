"""GPU parity tests for the Isihara ICNN kernel through the C ABI, the callable protocol and the torch custom
op, against the golden made by the reference's own torch code (float32-level tolerance, see isi_util.RTOL)."""

import numpy as np
import pytest

import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200 import synthetic as syn
from isi_util import check, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model(ctx):
    g, sd = load_golden()
    return eo.Isihara(sd, H_flat=g["H_flat"], ctx=ctx)


def test_against_reference_golden(model):
    g, _ = load_golden()
    with pytest.raises(NotImplementedError):
        model((0,))
    dP, P = model((1,))(g["F"].reshape(-1, 1, 2, 2))  # operand shape (n_cells, n_points, 2, 2)
    assert dP.shape == (16 * g["F"].shape[0],) and P.shape == (4 * g["F"].shape[0],)
    check(dP, P, g)
    assert np.abs(P[:4]).max() < 1e-6  # P(F = I) = 0 by construction


def test_default_correction_matches_the_references(ctx):
    g, sd = load_golden()
    m = eo.Isihara(sd, ctx=ctx)  # H_flat evaluated by this library: -P_NN(F = I)
    np.testing.assert_allclose(m.H_flat, g["H_flat"], atol=5e-7)  # the reference value is float32 rounding noise
    dP, P = m((1,))(g["F"])
    check(dP, P, g)


@pytest.mark.parametrize("n", [1, 127, 128, 129, 10_001])
def test_ragged_sizes_device_and_host_agree(ctx, model, n):
    F = syn.isihara_batch(n, seed=n)
    dP, P = model((1,))(F)
    dF, ddP, dPd = ctx.to_device(F.reshape(-1)), ctx.empty((16 * n,)), ctx.empty((4 * n,))
    model.eval_device(dF, ddP, dPd)
    ctx.sync()
    assert np.array_equal(ddP.to_host(), dP) and np.array_equal(dPd.to_host(), P)
    T = dP.reshape(n, 4, 4)
    assert np.abs(T - T.transpose(0, 2, 1)).max() <= 1e-12 * np.abs(T).max()
    assert np.isfinite(dP).all()


def test_torch_custom_op(model):
    torch = pytest.importorskip("torch")
    g, _ = load_golden()
    op = eo.register_torch_op(model)
    F = torch.from_numpy(g["F"]).cuda()
    dP, P = op(F)
    assert dP.is_cuda and dP.shape == (g["F"].shape[0], 4, 4) and P.shape == (g["F"].shape[0], 4)
    check(dP.cpu().numpy(), P.cpu().numpy(), g)
    dP2, P2 = torch.ops.eo.isihara_dP_dF(torch.from_dlpack(F))  # DLPack round trip
    assert torch.equal(dP2, dP) and torch.equal(P2, P)
    with pytest.raises(Exception):
        op(F.float())
