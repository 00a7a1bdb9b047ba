"""TEST INFRASTRUCTURE - writes tests/golden/*.npz by executing the REFERENCE's own
constitutive source (see oracle/ref_exec.py).  Runs only in the build container, where
/root/reference exists; the GPU box uses the committed .npz files.

    python -m oracle.gen_golden [vm] [heat] [mc] [isihara]
"""

from __future__ import annotations

import os
import sys

import numpy as np

import sys as _sys, os as _os
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from . import inputs, ref_exec  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def gen_vm():
    ns = ref_exec.load_von_mises(num_quadrature_points=3)
    out = {}
    for kind in ("mixed", "elastic", "plastic"):
        n = 1026  # 342 cells x 3 points
        deps, sigma_n, p = inputs.vm_batch(n, seed=0, kind=kind)
        Ct, sig, dp = ns["return_mapping"](deps.reshape(-1, 3, 4), sigma_n.reshape(-1, 3, 4), p.reshape(-1, 3))
        out.update({f"{kind}_deps": deps, f"{kind}_sigma_n": sigma_n, f"{kind}_p": p,
                    f"{kind}_C_tang": Ct.reshape(n, 4, 4), f"{kind}_sigma": sig.reshape(n, 4), f"{kind}_dp": dp.reshape(n)})
    out["params"] = np.array([ns["E"], 0.3, ns["E_tangent"], ns["sigma_0"], ns["H"], ns["lmbda"], ns["mu"]])
    np.savez_compressed(os.path.join(GOLDEN, "vm_seed0_n1026.npz"), **out)
    print("vm golden written; plastic fraction (mixed):", float((out["mixed_dp"] > 0).mean()))


def gen_heat():
    h1 = ref_exec.load_heat_part1()
    h2 = ref_exec.load_heat_part2()
    n_cells, n_pts = 1366, 3
    T, sigma = inputs.heat_batch(n_cells * n_pts, seed=0)
    T2 = T.reshape(n_cells, n_pts)
    s2 = sigma.reshape(n_cells, n_pts * 2)
    np.savez_compressed(
        os.path.join(GOLDEN, "heat_seed0_n4098.npz"),
        T=T, sigma=sigma,
        k=h1["k_impl"](T2), dk=h1["dkdT_impl"](T2),
        q=h2["q_impl"](T2, s2), dqdT=h2["dqdT_impl"](T2, s2), dqdsigma=h2["dqdsigma_impl"](T2, s2),
    )
    print("heat golden written")


def gen_isihara():
    """Isihara ICNN golden: the reference's own torch code (demo_hyperelasticity.py:221-315, 362-381, 429-456)
    with the shipped Isihara_noise=high.pth, eager (no torch.compile).  The state dict travels with the golden so
    that the GPU box (no /root/reference) can build the model."""
    import torch

    ns = ref_exec.load_isihara()
    n = 2049
    F = inputs.isihara_batch(n, seed=0)
    F[0] = [1.0, 0.0, 0.0, 1.0]  # the undeformed state: P = 0 by construction (:362-381)
    dP, P = ns["dP_dF_impl"](F)
    sd = {k: v.numpy() for k, v in ns["model"].state_dict().items()}
    np.savez_compressed(os.path.join(GOLDEN, "isihara_seed0_n2049.npz"), F=F, dP=dP.reshape(n, 4, 4), P=P.reshape(n, 4),
                        H_flat=ns["H_flat"].numpy().astype(np.float64), **{"sd/" + k: v for k, v in sd.items()})
    print("isihara golden written; |P(F=I)| =", float(np.abs(P.reshape(n, 4)[0]).max()))


_MC_NS = None


def _mc_init():
    global _MC_NS
    import torch

    torch.set_num_threads(1)
    _MC_NS = ref_exec.load_mohr_coulomb()


def _mc_one(args):
    """One quadrature point through the reference's own `dsigma_ddeps` (demo_mc:555)."""
    import torch

    de, sn = args
    Ct, aux = _MC_NS["dsigma_ddeps"](torch.as_tensor(de), torch.as_tensor(sn))
    sig, niter, yielding, norm_res, dlambda = aux
    return (Ct.numpy(), sig.numpy(), int(niter), float(yielding), float(norm_res), float(dlambda))


def _mc_run(deps, sigma_n, procs=8):
    import multiprocessing as mp

    with mp.get_context("spawn").Pool(procs, initializer=_mc_init) as pool:
        res = pool.map(_mc_one, [(deps[i], sigma_n[i]) for i in range(deps.shape[0])], chunksize=1)
    return {
        "C_tang": np.stack([r[0] for r in res]), "sigma": np.stack([r[1] for r in res]),
        "niter": np.array([r[2] for r in res], dtype=np.int32), "yielding": np.array([r[3] for r in res]),
        "norm_res": np.array([r[4] for r in res]), "dlambda": np.array([r[5] for r in res]),
    }


def gen_mc():
    """Mohr-Coulomb goldens: the reference's source (demo_mc:282-555) executed over the torch.func
    JAX shim.  ~30 s per plastic point -> small sets.  The stress paths (sigma_n) are walked with the
    C++ oracle's stress update; every golden point is then evaluated by the reference itself."""
    from . import constitutive as oc
    from . import native

    prm = oc.MohrCoulombParams()
    step = lambda d, s: native.mc_stress(d, s, prm, parallel=True)[0]  # noqa: E731
    d, s = inputs.mc_demo_path(10, 9, stepper=step)
    out = _mc_run(d, s)
    np.savez_compressed(os.path.join(GOLDEN, "mc_path_10x9.npz"), deps=d, sigma_n=s, **out)
    print("mc path golden written; niter histogram", np.unique(out["niter"], return_counts=True))
    d, s = inputs.mc_batch(96, seed=0, stepper=step)
    out = _mc_run(d, s)
    np.savez_compressed(os.path.join(GOLDEN, "mc_rand_seed0_n96.npz"), deps=d, sigma_n=s, **out)
    print("mc random golden written; niter histogram", np.unique(out["niter"], return_counts=True))


def gen_mc_big(n: int = 2048, seed: int = 1, procs: int | None = None):
    """A larger random Mohr-Coulomb golden (>= 2000 points so that the Lode-corner set is populated):
    same recipe as gen_mc, seed 1.  About 30 s of CPU per plastic point: run it in the background
    (`nice python -m oracle.gen_golden mc_big`)."""
    from . import constitutive as oc
    from . import native

    prm = oc.MohrCoulombParams()
    step = lambda d, s: native.mc_stress(d, s, prm, parallel=True)[0]  # noqa: E731
    d, s = inputs.mc_batch(n, seed=seed, stepper=step)
    out = _mc_run(d, s, procs or int(os.environ.get("GEN_PROCS", "6")))
    np.savez_compressed(os.path.join(GOLDEN, f"mc_rand_seed{seed}_n{n}.npz"), deps=d, sigma_n=s, **out)
    print("mc big golden written; niter histogram", np.unique(out["niter"], return_counts=True))


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    what = sys.argv[1:] or ["vm", "heat"]
    for w in what:
        globals()[f"gen_{w}"]()
