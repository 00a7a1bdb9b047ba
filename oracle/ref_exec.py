"""TEST INFRASTRUCTURE - executes the reference's OWN constitutive callables.

The reference package cannot be imported here (needs dolfinx/ufl/basix/petsc4py,
`external_operator.py:7-15`), and its demos are scripts that build meshes at
import time.  The constitutive callables inside the demos, however, depend only
on NumPy / Numba / JAX / torch.  This module pulls selected top-level
definitions out of a demo file with `ast` (the source stays where it lies under
/root/reference - nothing is copied into this repo) and executes them in a
namespace where the FEM-only names are stubbed:

  * `PETSc.ScalarType`          -> numpy.float64
  * `jax` / `jax.numpy`         -> `oracle/jax_on_torch.py`, a float64 torch.func
                                   shim (JAX itself is not installable here)

It is used ONLY by `oracle/gen_golden.py` (to write `tests/golden/*.npz`) and by
CPU tests that are skipped when /root/reference is absent (the GPU box).
"""

from __future__ import annotations

import ast
import os
import types

REFERENCE_ROOT = os.environ.get("EO_REFERENCE_ROOT", "/root/reference")
DEMO_DIR = os.path.join(REFERENCE_ROOT, "doc", "demo")


def reference_available() -> bool:
    return os.path.isdir(DEMO_DIR)


def _target_names(node: ast.AST) -> set[str]:
    names: set[str] = set()
    if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
        names.add(node.name)
    elif isinstance(node, ast.Assign):
        for t in node.targets:
            for sub in ast.walk(t):
                if isinstance(sub, ast.Name):
                    names.add(sub.id)
    elif isinstance(node, ast.AnnAssign) and isinstance(node.target, ast.Name):
        names.add(node.target.id)
    elif isinstance(node, ast.Expr):
        # e.g. `deviatoric[:3, :3] -= ...` is AugAssign, handled below
        pass
    elif isinstance(node, ast.AugAssign):
        for sub in ast.walk(node.target):
            if isinstance(sub, ast.Name):
                names.add(sub.id)
    return names


def exec_reference_symbols(demo_file: str, wanted: list[str], namespace: dict) -> dict:
    """Execute, in file order, every top-level statement of `demo_file` that
    defines (or augments) one of `wanted`.  Returns the namespace."""
    path = os.path.join(DEMO_DIR, demo_file)
    with open(path) as fh:
        src = fh.read()
    tree = ast.parse(src, filename=path)
    want = set(wanted)
    found: set[str] = set()
    for node in tree.body:
        names = _target_names(node)
        if names & want:
            mod = ast.Module(body=[node], type_ignores=[])
            code = compile(mod, path, "exec")
            exec(code, namespace)
            found |= names & want
    missing = want - found
    if missing:
        raise KeyError(f"{demo_file}: reference symbols not found: {sorted(missing)}")
    return namespace


def petsc_stub():
    import numpy as np

    mod = types.ModuleType("PETSc")  # a real module: numba can type its attributes
    mod.ScalarType = np.float64
    return mod


# --------------------------------------------------------------------------
# von Mises  (doc/demo/demo_plasticity_von_mises.py:183-204, 298-332)
# --------------------------------------------------------------------------
def load_von_mises(num_quadrature_points: int = 3):
    import numba
    import numpy as np

    ns = {"np": np, "numba": numba, "PETSc": petsc_stub(), "num_quadrature_points": num_quadrature_points}
    exec_reference_symbols(
        "demo_plasticity_von_mises.py",
        ["E", "E_tangent", "H", "sigma_0", "lmbda", "mu", "C_elas", "deviatoric", "return_mapping"],
        ns,
    )
    return ns


# --------------------------------------------------------------------------
# heat  (part1.py:247-272, part2.py:209-261)
# --------------------------------------------------------------------------
def load_heat_part1():
    import numpy as np

    ns = {"np": np}
    exec_reference_symbols("demo_nonlinear_heat_equation_part1.py", ["A", "B", "k_impl", "dkdT_impl"], ns)
    return ns


def load_heat_part2(gdim: int = 2):
    import numpy as np

    ns = {"np": np, "gdim": gdim}
    exec_reference_symbols(
        "demo_nonlinear_heat_equation_part2.py",
        ["A", "B", "Id", "k", "q_impl", "dqdT_impl", "dqdsigma_impl"],
        ns,
    )
    # the demo re-binds gdim from the mesh; keep the caller's
    ns["gdim"] = gdim
    return ns


# --------------------------------------------------------------------------
# Mohr-Coulomb  (demo_plasticity_mohr_coulomb.py:110-116, 282-533, 555)
# --------------------------------------------------------------------------
def load_mohr_coulomb():
    import numpy as np

    from . import jax_on_torch as jot

    ns = {
        "np": np,
        "jax": jot.jax,
        "jnp": jot.jnp,
        "PETSc": petsc_stub(),
        "stress_dim": 4,
    }
    exec_reference_symbols(
        "demo_plasticity_mohr_coulomb.py",
        [
            "E", "nu", "c", "phi", "psi", "theta_T", "a",
            "J3", "J2", "theta", "sign", "coeff1", "coeff2", "coeff3", "C", "B", "A", "K", "a_g",
            "dev", "tr", "surface", "f", "g", "dgdsigma",
            "lmbda", "mu", "C_elas", "S_elas", "ZERO_VECTOR",
            "deps_p", "r_g", "r_f", "r", "drdy",
            "Nitermax", "tol", "ZERO_SCALAR", "return_mapping", "dsigma_ddeps",
        ],
        ns,
    )
    jot.numpy_constants_to_torch(ns, ["dev", "tr", "C_elas", "ZERO_VECTOR"])
    return ns


# --------------------------------------------------------------------------
# Isihara ICNN  (demo_hyperelasticity.py:221-315, 362-381, 429-456)
# --------------------------------------------------------------------------
def load_isihara():
    import numpy as np
    import torch

    ns = {"np": np, "torch": torch}
    cwd = os.getcwd()
    os.chdir(DEMO_DIR)  # the demo loads the .pth by relative path (:314)
    try:
        exec_reference_symbols(
            "demo_hyperelasticity.py",
            ["convexLinear", "ICNN", "n_input", "n_output", "n_hidden", "dropout"],
            ns,
        )
        # `model = ICNN(...)` then `model.load_state_dict(...)`/`model.eval()` are
        # expression statements (:314-315) and `model = torch.compile(model)`
        # (:416) must be skipped, so bind the model by hand exactly as :310-315.
        ns["model"] = ns["ICNN"](
            n_input=ns["n_input"], n_hidden=ns["n_hidden"], n_output=ns["n_output"], dropout=ns["dropout"]
        )
        ns["model"].load_state_dict(torch.load("Isihara_noise=high.pth"))
        ns["model"].eval()
        exec_reference_symbols(
            "demo_hyperelasticity.py",
            ["F_0", "W_NN_0", "P_NN_0", "H_flat", "H", "compute_stress_local",
             "vectorized_stress_and_tangent", "dP_dF_impl"],
            ns,
        )
    finally:
        os.chdir(cwd)
    return ns
