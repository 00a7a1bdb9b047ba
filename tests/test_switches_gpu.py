"""The per-cell geometry cache of the tabulation handle (`eo_tab_geometry`), the whole-line tangent stores of the fused
kernel (`eo_st_tangent_quad`) and the one-thread-per-cell mapping of the residual step (`form_vm_step_cell_kernel`) change
how values are fetched, stored and scheduled, never the values: processes run with the switches flipped (EO_GEOM_CACHE=0,
EO_QUAD_STORE=0, EO_STEP_CELL=0 - the round-1 code paths) must reproduce the default results BIT FOR BIT for every
per-point array; the scattered integrals (atomic adds in arbitrary order) agree to 1e-13 of their scale."""

import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(tmp_path, name, **env):
    out = str(tmp_path / (name + ".npz"))
    e = dict(os.environ)
    e.update(env)
    subprocess.run([sys.executable, os.path.join(HERE, "switch_worker.py"), out], check=True, env=e, timeout=600)
    return np.load(out)


def test_geometry_cache_and_quad_stores_do_not_change_results(tmp_path):
    ref = _run(tmp_path, "default")
    for name, env in (("nocache", {"EO_GEOM_CACHE": "0"}), ("noquad", {"EO_QUAD_STORE": "0"}),
                      ("perpoint", {"EO_STEP_CELL": "0"}),
                      ("round1", {"EO_GEOM_CACHE": "0", "EO_QUAD_STORE": "0", "EO_STEP_CELL": "0"})):
        got = _run(tmp_path, name, **env)
        assert sorted(got.files) == sorted(ref.files)
        for k in ref.files:
            a, b = ref[k], got[k]
            assert a.shape == b.shape and np.isfinite(a).all(), (name, k)
            if k in ("action", "vector") or k.startswith("step_b"):
                assert np.abs(a - b).max() <= 1e-13 * np.abs(a).max(), (name, k)
            else:
                assert np.array_equal(a, b), (name, k)
    assert 0.2 < (ref["fused_dp_0"] > 0).mean() < 0.8  # the comparison covers elastic and plastic points
