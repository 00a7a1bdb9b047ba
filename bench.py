#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 quadrature-point engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model vm|heat] [--n QP_PER_GPU]
    python bench.py --impl reference ...          # the CPU arm (reference algorithm on host cores)

Metric (BASELINE.json): quadrature points per second for stress + consistent tangent +
internal state, on the von Mises configuration (configs[1]) evaluated on a synthetic batch of
--n points per GPU (default 1e8 = SURVEY.md section 8d's per-GPU target size).

  value      device-resident: one kernel launch per step over all --n points; inputs (strain
             increment, committed history) already in HBM; timed with CUDA events on the
             library's compute stream; max over ranks.
  e2e        through the public API `VonMises.C_tang_impl` (the callable that
             `evaluate_external_operators` invokes): strain increment in pinned HOST memory,
             tangent/stress/dp delivered into pinned HOST arrays, H2D and D2H inside the timed
             region (history stays resident in HBM, as north_star prescribes).
  roofline   algorithmic bytes per QP (240 for von Mises: SURVEY.md section 8d) x n / mean kernel time,
             against MEASURED_PEAKS.json's hbm_gbs.
  cpu_baseline  the C restatement of the reference's Numba kernel (oracle/, "port") on the
             box's host cores, bounded sample.
One JSON line on stdout (rank 0).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_QP = {"vm": 240, "jitvm": 240, "jitfused": 235, "jitvm3d": 448, "heat": 88, "mc": 252, "tab": 176.0 / 3.0, "fused": 235, "isihara": 192,
                # device-side consumers (csrc/form.cu): + read-modify-write of the DOF vector (2 nodes x 2 x 8 B per cell, twice)
                "step": 235 + 64.0 / 3.0, "action": 128 + 80.0 / 3.0 + 64.0 / 3.0}
# per P2-triangle cell: dofmap 24 + x_dofmap 12 + 6 gathered dofs x 16 + 3 gathered vertices x 16 = 180 B (vs 80 B unique)
_CELL_GATHER = (24 + 12 + 96 + 48) / 3.0
GATHERED_BYTES_PER_QP = {"tab": _CELL_GATHER + 32, "fused": _CELL_GATHER + 40 + 168, "jitfused": _CELL_GATHER + 40 + 168,
                         "step": _CELL_GATHER + 40 + 168 + 12 * 16 / 3.0, "action": _CELL_GATHER + 128 + 12 * 16 / 3.0}
METRIC = "quadrature points per second (stress + consistent tangent + internal state)"


def _measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------- CPU arm
def cpu_port_rate(model: str, sample_n: int, min_seconds: float, parallel: bool = True):
    """Time the oracle's C restatement on host cores over a bounded sample; returns (QP/s, cores, passes)."""
    from oracle import constitutive as oc
    from oracle import inputs, native

    native.build()
    native.use_all_cores()
    cores = native.num_threads() if parallel else 1
    if model == "vm":
        deps, sigma_n, p = inputs.vm_batch(sample_n, seed=0)
        prm = oc.VonMisesParams()
        fn = lambda: native.vm_return_mapping(deps, sigma_n, p, prm, parallel=parallel)  # noqa: E731
    elif model == "mc":
        mprm = oc.MohrCoulombParams()
        deps, sigma_n = inputs.mc_batch(sample_n, seed=0,
                                        stepper=lambda d, s: native.mc_stress(d, s, mprm, parallel=True)[0])
        fn = lambda: native.mc_return_mapping(deps, sigma_n, mprm, parallel=parallel)  # noqa: E731
    else:
        T, sigma = inputs.heat_batch(sample_n, seed=0)

        def fn():
            for w in ("q", "dqdT", "dqdsigma"):
                native.heat(w, T, sigma, parallel=parallel)
    fn()  # warm-up (page faults, thread pool)
    best, passes, t_all = float("inf"), 0, time.perf_counter()
    while True:
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        best = min(best, dt)
        passes += 1
        if time.perf_counter() - t_all >= min_seconds and passes >= 3:
            break
    return sample_n / best, cores, passes


def cpu_tab_rate(model: str, min_seconds: float):
    """CPU baseline for the cell-loop legs (tab / fused / step / action): the OpenMP C restatement of the von Mises demo's
    cell loop (oracle/csrc/forms_oracle.c: tabulation of the Mandel strain, radial return, residual scatter, tangent
    action) on all host cores, on a bounded mesh (1.28 M cells = 3.84 M quadrature points)."""
    from dolfinx_external_operator_b200 import elements as el
    from dolfinx_external_operator_b200 import synthetic as syn
    from oracle import constitutive as oc
    from oracle import native

    cores = native.use_all_cores()
    m = syn.triangle_mesh(800, 800, 2, jitter=0.2, seed=0)
    m["phi"], m["dphi"] = el.lagrange_triangle(2, el.triangle_quadrature(2))
    m["dpsi"] = el.p1_geometry_derivatives(2)
    u = syn.smooth_displacement(m["dof_coords"], scale=1.5e-3, seed=0).reshape(-1)
    nq = 3 * m["dofmap"].shape[0]
    _, sn, p = syn.vm_batch(nq, seed=0)
    W3 = el.triangle_quadrature_weights(2)
    prm = oc.VonMisesParams()
    mode = {"tab": "tab", "fused": "fused", "jitfused": "fused", "step": "step", "action": "action"}[model]
    Ct0 = native.forms_p2_cells("fused", m, W3, u, prm, sn, p)[0] if mode == "action" else None

    def fn():
        if mode == "action":
            native.forms_p2_cells("action", m, W3, u, C_tang=Ct0)
        else:
            native.forms_p2_cells(mode, m, W3, u, prm, sn, p)

    fn()
    best, passes, t_all = float("inf"), 0, time.perf_counter()
    while True:
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
        passes += 1
        if time.perf_counter() - t_all >= min_seconds and passes >= 3:
            break
    what = {"tab": "tabulation of the Mandel strain", "fused": "tabulation + von Mises radial return",
            "step": "tabulation + von Mises radial return + residual scatter",
            "action": "tangent action (tabulate x, contract with the stored tangent, scatter)"}[mode]
    return {"value": nq / best, "unit": "QP/s", "cores": cores, "kind": "port",
            "sample": f"{nq} QPs x {passes} passes (best pass, includes allocating the result arrays); OpenMP C restatement "
                      f"of the demo's cell loop ({what}; the reference runs it serially per MPI rank through DOLFINx/FFCx "
                      "and Numba)"}


def e2e_device_consumers(ctx, eo, inputs, ne, rank, world, max_over_ranks, barrier):
    """Supplementary end-to-end figure of the default (von Mises) line: the same constitutive update consumed ON the
    device - host displacement vector in, host residual vector out, tangent resident for the
    matrix-free action - instead of shipping 168 B per point back to DOLFINx's assembler."""
    from dolfinx_external_operator_b200 import elements as el

    nxy = max(2, int(round((ne / 6.0) ** 0.5)))
    mesh = inputs.triangle_mesh(nxy, nxy, 2, jitter=0.2, seed=rank)
    nq = 3 * mesh["dofmap"].shape[0]
    phi, dphi = el.lagrange_triangle(2, el.triangle_quadrature(2))
    tab = eo.Tabulator(dofmap=mesh["dofmap"], x_dofmap=mesh["x_dofmap"], x=mesh["x"], phi=phi, dphi=dphi, bs=2,
                       n_dofs=mesh["n_dofs"], ctx=ctx)
    forms = eo.QuadratureForms(tab, el.triangle_quadrature_weights(2))
    vm = eo.VonMises(n_qp=nq, ctx=ctx)
    _, sn_t, p_t = inputs.vm_batch(min(nq, 1 << 22), seed=rank)
    _tile_to_device(ctx, vm.sigma_n_dev, sn_t, nq, 4)
    _tile_to_device(ctx, vm.p_dev, p_t, nq, 1)
    nd = 2 * tab.n_dofs
    u_h, b_h, y_h = ctx.pinned_empty((nd,)), ctx.pinned_empty((nd,)), ctx.pinned_empty((nd,))
    u_h[:] = inputs.smooth_displacement(mesh["dof_coords"], scale=1.5e-3, seed=rank).reshape(-1)
    del mesh
    out = {}
    for name, call in (("residual", lambda: forms.vm_residual(vm, u_h, out=b_h)),
                       ("tangent_action", lambda: forms.action("mandel_strain", "mandel_strain", forms.C_tang, u_h, out=y_h))):
        for _ in range(2):
            call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            call()
        ctx.sync()
        dt = max_over_ranks(time.perf_counter() - t0)
        out[name] = {"value": world * nq * 5 / dt, "unit": "QP/s", "ms_per_step": 1e3 * dt / 5}
    out.update(h2d_bytes_per_step=8 * nd, d2h_bytes_per_step=8 * nd, qp_per_step_per_gpu=nq,
               api="QuadratureForms.vm_residual(vm, u_host, out=b_host) / .action(..., x_host, out=y_host): eo_form_vm_step, "
                   "eo_form_action; P2 vector triangles, 3 points per cell; boundary terms, lifting and the solve stay with "
                   "the caller")
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = int(args.cpu_sample)
    from oracle import constitutive as oc
    from oracle import inputs, native

    native.build()
    native.use_all_cores()
    cores = native.num_threads()
    if args.model in ("tab", "fused", "step", "action"):
        r = cpu_tab_rate(args.model, 5.0)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "QP/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOADS[args.model]}, "cpu_baseline": r,
                "e2e": {"value": r["value"], "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return
    if args.model == "mc":
        sample = min(sample, 200_000)
        mprm = oc.MohrCoulombParams()
        deps, sigma_n = inputs.mc_batch(sample, seed=0,
                                        stepper=lambda d, s: native.mc_stress(d, s, mprm, parallel=True)[0])
        ref_step = lambda: native.mc_return_mapping(deps, sigma_n, mprm, parallel=True)  # noqa: E731
        what = ("Mohr-Coulomb return mapping + AD-through-the-loop tangent (demo_plasticity_mohr_coulomb.py:474-555), "
                "demo stress-path family, ~34% plastic points")
        port = ("C++ nested-dual-number restatement of the reference's JAX program (JAX is not installable "
                "offline), OpenMP")
    elif args.model == "heat":
        T, sig = inputs.heat_batch(sample, seed=0)

        def ref_step():
            for w in ("q", "dqdT", "dqdsigma"):
                native.heat(w, T, sig, parallel=True)
        what = "nonlinear heat flux q, dq/dT, dq/dsigma (demo_nonlinear_heat_equation_part2.py:219-261)"
        port = "C restatement of the reference's NumPy functions, OpenMP"
    else:
        deps, sigma_n, p = inputs.vm_batch(sample, seed=0)
        prm = oc.VonMisesParams()
        ref_step = lambda: native.vm_return_mapping(deps, sigma_n, p, prm, parallel=True)  # noqa: E731
        what = ("von Mises return mapping (demo_plasticity_von_mises.py:298-332), plane-strain Mandel 4-vectors, "
                "~53% plastic points")
        port = "C restatement of the reference's Numba kernel (the reference kernel itself is serial @numba.njit), OpenMP"
    for _ in range(args.warmup):
        ref_step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref_step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "QP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": what, "qp_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "QP/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} QPs per step; {port} over {cores} threads"},
        "e2e": {"value": value, "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
WORKLOADS = {
    "vm": "von Mises return mapping, plane-strain Mandel 4-vectors (BASELINE configs[1] callable at configs[4] "
          "batch size)",
    "heat": "nonlinear heat flux q, dq/dT, dq/dsigma fused (BASELINE configs[0] callables at configs[4] batch size)",
    "mc": "Mohr-Coulomb return mapping with apex smoothing, local Newton + tangent through the iterations "
          "(BASELINE configs[2] callable at configs[4] batch size), demo stress-path family",
    "isihara": "Isihara ICNN hyperelasticity (BASELINE configs[3] callable at configs[4] batch size): stress P and "
               "tangent dP/dF from the 3-64-64-64-1 float32 network, float64 invariants",
    "jitvm": "von Mises return mapping written as a user model for the run-time compiled (NVRTC) generic path: "
             "stress + tangent by forward-mode dual numbers + plastic multiplier, plane-strain Mandel 4-vectors",
    "jitvm3d": "EXTENSION (BASELINE configs[4] '3D'): von Mises return mapping for 6-component Mandel vectors as a "
               "run-time compiled user model, 6x6 tangent by dual numbers (no reference implementation: the demos are "
               "plane strain)",
    "jitfused": "operand tabulation fused into the run-time compiled (NVRTC) von Mises user model: P2 vector field, 3 points "
                "per triangle, strain never stored, tangent by dual numbers",
    "tab": "operand tabulation: Mandel strain of a P2 vector field at 3 quadrature points per triangle "
           "(evaluate_operands for the von Mises / Mohr-Coulomb demos), structured jittered mesh",
    "fused": "operand tabulation fused with the von Mises return mapping (strain never stored), P2 vector field, "
             "3 quadrature points per triangle",
    "step": "one Newton residual evaluation on the device (SURVEY.md 8f rank 1): Mandel strain of a P2 vector field -> "
            "von Mises return mapping (tangent / stress / dp stored in HBM) -> b = int sigma . epsilon(v) dx, ONE kernel; "
            "3 quadrature points per triangle",
    "action": "matrix-free tangent action y = J x with J = int (C_tang epsilon(u_hat)) . epsilon(v) dx, tangent resident "
              "in HBM (what a Krylov method needs from assemble_matrix); P2 vector field, 3 points per triangle",
}


def _nanmax(vals):
    v = [x for x in vals if x == x]
    return max(v) if v else float("nan")


def _same(a, b):
    return a == b or (a != a and b != b)


def _tile_to_device(ctx, dst, tile: np.ndarray, n: int, width: int, itemsize: int = 8):
    """Repeat a seeded host tile into a device array of n rows."""
    tile_n = tile.shape[0]
    d_tile = ctx.to_device(np.ascontiguousarray(tile).reshape(-1))
    for r in range(0, n, tile_n):
        m = min(tile_n, n - r)
        ctx.copy(dst.ptr + r * width * itemsize, d_tile, m * width * itemsize)
    ctx.sync()
    d_tile.free()


def run_gpu_arm(args):
    import ctypes as C

    import dolfinx_external_operator_b200 as eo
    from dolfinx_external_operator_b200 import synthetic as inputs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))

    def barrier():
        if dist is not None:
            import torch

            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    numa_cores = None
    if world > 1 and not args.no_numa_bind:
        # one rank per GPU: keep this rank's threads and pinned buffers on the GPU's own NUMA node
        from dolfinx_external_operator_b200.parallel import bind_to_gpu_numa

        numa_cores = bind_to_gpu_numa(local_rank)
    ctx = eo.Context(local_rank)
    n = int(args.n)
    model = args.model
    K, W = args.steps, args.warmup
    tile_n = min(n, 1 << 22)
    extra_cfg = {}

    # ---- synthetic inputs: a seeded tile (seed = rank) repeated to n points, resident in HBM
    if model == "vm":
        vm = eo.VonMises(n_qp=n, ctx=ctx, state_layout=args.state_layout)
        deps_t, sn_t, p_t = inputs.vm_batch(tile_n, seed=rank)
        d_deps = ctx.empty((n * 4,))
        _tile_to_device(ctx, d_deps, deps_t, n, 4)
        if args.state_layout == "soa":
            for c in range(4):
                col = ctx.to_device(np.ascontiguousarray(sn_t[:, c]))
                for r in range(0, n, tile_n):
                    m = min(tile_n, n - r)
                    ctx.copy(vm.sigma_n_dev.ptr + (c * n + r) * 8, col, m * 8)
                ctx.sync()
                col.free()
        else:
            _tile_to_device(ctx, vm.sigma_n_dev, sn_t, n, 4)
        _tile_to_device(ctx, vm.p_dev, p_t, n, 1)
        d_Ct = ctx.empty((n * 16,))

        def step():
            vm.eval_device(d_deps, d_Ct)
    elif model == "jitvm3d":
        from dolfinx_external_operator_b200 import jit_models as jm

        jv = jm.von_mises_3d(ctx=ctx)
        rng = np.random.default_rng(rank)
        deps_t, sn_t = rng.normal(0.0, 2e-3, (tile_n, 6)), rng.normal(0.0, 100.0, (tile_n, 6))
        p_t = np.abs(rng.normal(0.0, 1e-3, tile_n))
        d_deps = ctx.empty((n * 6,))
        _tile_to_device(ctx, d_deps, deps_t, n, 6)
        jv.state = [ctx.empty((n * 6,)), ctx.empty((n,))]
        _tile_to_device(ctx, jv.state[0], sn_t, n, 6)
        _tile_to_device(ctx, jv.state[1], p_t, n, 1)
        d_Ct, d_sig, d_dp = ctx.empty((n * 36,)), ctx.empty((n * 6,)), ctx.empty((n,))
        jv.compile((1,))

        def step():
            jv.eval_device((1,), [d_deps], d_Ct, d_sig, [d_dp])
    elif model == "jitvm":
        from dolfinx_external_operator_b200 import jit_models as jm

        jv = jm.von_mises(ctx=ctx)
        deps_t, sn_t, p_t = inputs.vm_batch(tile_n, seed=rank)
        d_deps = ctx.empty((n * 4,))
        _tile_to_device(ctx, d_deps, deps_t, n, 4)
        jv.state = [ctx.empty((n * 4,)), ctx.empty((n,))]
        _tile_to_device(ctx, jv.state[0], sn_t, n, 4)
        _tile_to_device(ctx, jv.state[1], p_t, n, 1)
        d_Ct, d_sig, d_dp = ctx.empty((n * 16,)), ctx.empty((n * 4,)), ctx.empty((n,))
        jv.compile((1,))

        def step():
            jv.eval_device((1,), [d_deps], d_Ct, d_sig, [d_dp])
    elif model == "mc":
        from dolfinx_external_operator_b200._lib import McParams

        mc = eo.MohrCoulomb(ctx=ctx, history=None)
        tile_n = min(n, 1 << 20)
        # the stress paths are walked with the GPU kernel itself as the stress update
        deps_t, sn_t = inputs.mc_batch(tile_n, seed=rank, stepper=mc.stress_update)
        d_deps, d_sn = ctx.empty((n * 4,)), ctx.empty((n * 4,))
        _tile_to_device(ctx, d_deps, deps_t, n, 4)
        _tile_to_device(ctx, d_sn, sn_t, n, 4)
        d_Ct, d_sig = ctx.empty((n * 16,)), ctx.empty((n * 4,))
        d_it = ctx.empty((n,), np.int32)
        d_yl, d_nr, d_dl = ctx.empty((n,)), ctx.empty((n,)), ctx.empty((n,))
        prm = McParams(mc.E, mc.nu, mc.c, mc.phi, mc.psi, mc.theta_T, mc.a, mc.tol, mc.Nitermax)
        scheme = {"queue": 0, "simple": 1, "queue-noaffinity": 2, "queue-onepass": 3}[args.mc_scheme]
        extra_cfg["mc_scheme"] = args.mc_scheme

        def step():
            ctx.check(ctx.lib.eo_mc_eval_scheme(ctx.handle, C.byref(prm), d_deps.ptr, d_sn.ptr, d_Ct.ptr, d_sig.ptr,
                                                d_it.ptr, d_yl.ptr, d_nr.ptr, d_dl.ptr, n, scheme))
    elif model == "isihara":
        gpath = os.path.join(ROOT, "tests", "golden", "isihara_seed0_n2049.npz")  # carries the reference's state dict
        g = np.load(gpath)
        isi = eo.Isihara({k[3:]: g[k] for k in g.files if k.startswith("sd/")}, ctx=ctx)
        F_t = inputs.isihara_batch(tile_n, seed=rank)
        d_F = ctx.empty((n * 4,))
        _tile_to_device(ctx, d_F, F_t, n, 4)
        d_dP, d_P = ctx.empty((n * 16,)), ctx.empty((n * 4,))

        def step():
            isi.eval_device(d_F, d_dP, d_P)
    elif model in ("tab", "fused", "jitfused", "step", "action"):
        from dolfinx_external_operator_b200 import elements as el

        nxy = max(2, int(round((n / 6.0) ** 0.5)))
        mesh = inputs.triangle_mesh(nxy, nxy, 2, jitter=0.2, seed=rank)
        n_cells = mesh["dofmap"].shape[0]
        n = 3 * n_cells  # quadrature points actually processed
        phi, dphi = el.lagrange_triangle(2, el.triangle_quadrature(2))
        tab = eo.Tabulator(dofmap=mesh["dofmap"], x_dofmap=mesh["x_dofmap"], x=mesh["x"], phi=phi, dphi=dphi, bs=2,
                           n_dofs=mesh["n_dofs"], ctx=ctx)
        d_u = ctx.to_device(inputs.smooth_displacement(mesh["dof_coords"], scale=1.5e-3, seed=rank).reshape(-1))
        extra_cfg.update(n_cells=n_cells, n_dofs=mesh["n_dofs"], element="P2 vector triangle, 3-point rule")
        del mesh
        if model == "tab":
            d_out = ctx.empty((n_cells, 3, 4))

            def step():
                tab.evaluate("mandel_strain", d_u, out=d_out)
        elif model in ("step", "action"):
            forms = eo.QuadratureForms(tab, el.triangle_quadrature_weights(2))
            vm = eo.VonMises(n_qp=n, ctx=ctx)
            _, sn_t, p_t = inputs.vm_batch(tile_n, seed=rank)
            _tile_to_device(ctx, vm.sigma_n_dev, sn_t, n, 4)
            _tile_to_device(ctx, vm.p_dev, p_t, n, 1)
            d_b = ctx.empty((2 * tab.n_dofs,))
            extra_cfg["vm_arithmetic"] = "exact" if args.fused_exact else "fast (2 divisions, FMA; identical flags)"
            extra_cfg["scatter"] = "fp64 RED.ADD per element-vector entry (12 per cell)"
            if model == "step":
                def step():
                    forms.vm_residual(vm, d_u, out=d_b, exact=args.fused_exact)
            else:
                forms.vm_residual(vm, d_u, out=d_b, exact=args.fused_exact)  # fills forms.C_tang
                d_y = ctx.empty((2 * tab.n_dofs,))

                def step():
                    forms.action("mandel_strain", "mandel_strain", forms.C_tang, d_u, out=d_y)
        elif model == "jitfused":
            from dolfinx_external_operator_b200 import jit_models as jm
            from dolfinx_external_operator_b200.tabulation import LazyOperand

            jv = jm.von_mises(ctx=ctx)
            _, sn_t, p_t = inputs.vm_batch(tile_n, seed=rank)
            jv.state = [ctx.empty((n * 4,)), ctx.empty((n,))]
            _tile_to_device(ctx, jv.state[0], sn_t, n, 4)
            _tile_to_device(ctx, jv.state[1], p_t, n, 1)
            dev = {"out": ctx.empty((n * 16,)), "value": ctx.empty((n * 4,)), "aux0": ctx.empty((n,))}
            lz = [LazyOperand(tab, 2, d_u)]
            dd = jv._deriv((1,))[1]

            def step():
                jv._evaluate_fused((1,), dd, 16, lz, dev)
        else:
            vm = eo.VonMises(n_qp=n, ctx=ctx)
            _, sn_t, p_t = inputs.vm_batch(tile_n, seed=rank)
            _tile_to_device(ctx, vm.sigma_n_dev, sn_t, n, 4)
            _tile_to_device(ctx, vm.p_dev, p_t, n, 1)
            d_Ct = ctx.empty((n * 16,))

            extra_cfg["vm_arithmetic"] = "exact" if args.fused_exact else "fast (2 divisions, FMA; identical flags)"

            def step():
                tab.vm_fused(vm, d_u, C_tang=d_Ct, exact=args.fused_exact)
    else:
        T_t, s_t = inputs.heat_batch(tile_n, seed=rank)
        d_T, d_s = ctx.empty((n,)), ctx.empty((n * 2,))
        _tile_to_device(ctx, d_T, T_t, n, 1)
        _tile_to_device(ctx, d_s, s_t, n, 2)
        d_q, d_dT, d_ds = ctx.empty((n * 2,)), ctx.empty((n * 2,)), ctx.empty((n * 4,))

        def step():
            ctx.check(ctx.lib.eo_heat_eval(ctx.handle, 1.0, 1.0, d_T.ptr, d_s.ptr, None, None, d_q.ptr, d_dT.ptr,
                                           d_ds.ptr, n))

    def collective():
        if dist is not None:
            from dolfinx_external_operator_b200.parallel import allreduce_stats_device

            allreduce_stats_device(ctx)

    # ---- device-resident timing
    for _ in range(W):
        step()
        collective()
    ctx.sync()
    ctx.stats_reset()
    ev = [ctx.event() for _ in range(2 * K + 2)]
    sampler = ClockSampler(local_rank)
    barrier()
    ctx.sync()
    sampler.start()
    launches0 = ctx.launch_count
    ctx.record(ev[0])
    for k in range(K):
        ctx.record(ev[2 + 2 * k])
        step()
        ctx.record(ev[3 + 2 * k])
        collective()
    ctx.record(ev[1])
    ctx.sync()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    total_ms = max_over_ranks(ctx.elapsed_ms(ev[0], ev[1]))
    kernel_ms = [ctx.elapsed_ms(ev[2 + 2 * k], ev[3 + 2 * k]) for k in range(K)]
    stats = ctx.stats()  # this rank's record over the K timed steps (the collective never modifies it)
    collective_check = None
    if dist is not None:
        # the K-th collective left the GLOBAL record of the K steps: it must equal the sum of the local records,
        # gathered here a second, independent way (host tensors through the store-backed object collective)
        import torch

        gstats = ctx.stats_global()
        loc = [None] * world
        dist.all_gather_object(loc, {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in stats.items()})
        want_points = sum(l["n_points"] for l in loc)
        want_plastic = sum(l["n_plastic"] for l in loc)
        want_hist = np.sum([np.asarray(l["niter_hist"], dtype=np.int64) for l in loc], axis=0)
        counted = model not in ("heat", "isihara", "tab", "action", "jitvm", "jitvm3d", "jitfused")
        ok = (gstats["n_points"] == want_points and gstats["n_plastic"] == want_plastic
              and np.array_equal(gstats["niter_hist"], want_hist)
              and _same(gstats["f_max"], _nanmax(l["f_max"] for l in loc))
              and _same(gstats["res_max"], _nanmax(l["res_max"] for l in loc))
              and (not counted or want_points == world * n * K))
        collective_check = {"ok": bool(ok), "n_points": gstats["n_points"], "expected_n_points": want_points,
                            "n_plastic": gstats["n_plastic"], "sum_of_local_n_plastic": want_plastic,
                            "what": "eo_stats global record after the last step == sum / max over the ranks' local records "
                                    "(one NCCL all-gather of the 1.7 KB record + combine kernel per step, on the "
                                    "collective stream, overlapping the next step)"}
        if not ok:
            print(f"[rank {rank}] statistics collective mismatch: {collective_check}", file=sys.stderr, flush=True)
            dist.destroy_process_group()
            sys.exit(3)
        stats = gstats
    value = world * n * K / (total_ms * 1e-3)

    # ---- end-to-end through the public callable (host buffers)
    e2e = None
    if model == "isihara" and args.e2e_n > 0:
        ne = int(min(args.e2e_n, n))
        F_h = ctx.pinned_empty((ne, 1, 2, 2))
        F_h.reshape(-1, 4)[:] = np.resize(F_t, (ne, 4))
        call = isi((1,))
        Ke = max(2, min(K, 5))
        for _ in range(2):
            out = call(F_h)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            out = call(F_h)
        ctx.sync()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * ne * Ke / dt, "unit": "QP/s", "h2d_bytes_per_step": 32 * ne, "d2h_bytes_per_step": 160 * ne,
               "qp_per_step_per_gpu": ne, "steps": Ke, "api": "Isihara((1,))(F) == external_function(derivatives)(*operands)"}
    if model in ("vm", "mc") and args.e2e_n > 0:
        ne = int(args.e2e_n)
        deps_h = ctx.pinned_empty((ne, 1, 4))  # (n_cells, n_points, 4), the operand shape of demo_vm:344
        flat = deps_h.reshape(-1, 4)
        if model == "vm":
            m_e = eo.VonMises(n_qp=ne, ctx=ctx)
            deps_t2, sn_t2, p_t2 = inputs.vm_batch(min(ne, 1 << 22), seed=rank)
            m_e.set_history(np.resize(sn_t2, (ne, 4)), np.resize(p_t2, ne))
            d2h = 168 * ne
            api = "VonMises((1,))(deps) == external_function(derivatives)(*operands), history resident in HBM"
        else:
            m_e = eo.MohrCoulomb(n_qp=ne, ctx=ctx)
            deps_t2, sn_t2 = deps_t, sn_t
            m_e.set_history(np.resize(sn_t2, (ne, 4)))
            d2h = (160 + 28) * ne
            api = "MohrCoulomb((1,))(deps) == external_function(derivatives)(*operands), history resident in HBM"
        for r in range(0, ne, deps_t2.shape[0]):
            m = min(deps_t2.shape[0], ne - r)
            flat[r:r + m] = deps_t2[:m]
        call = m_e((1,))
        Ke = max(2, min(K, 5))
        for _ in range(2):
            out = call(deps_h)
        barrier()
        t0 = time.perf_counter()
        checksum = 0.0
        for _ in range(Ke):
            out = call(deps_h)  # H2D + kernel + D2H, synchronous on return
            checksum += float(out[1][0])
        ctx.sync()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * ne * Ke / dt, "unit": "QP/s", "h2d_bytes_per_step": 32 * ne,
               "d2h_bytes_per_step": d2h, "qp_per_step_per_gpu": ne, "steps": Ke, "api": api}
        # host<->device transfer times on their own (CUDA events, pinned buffers): what bounds the end-to-end figure
        d_tmp = ctx.empty((16 * ne,))
        evs = [ctx.event() for _ in range(4)]
        ctx.copy(d_tmp, deps_h, 32 * ne)  # warm
        ctx.sync()
        ctx.record(evs[0])
        ctx.copy(d_tmp, deps_h, 32 * ne)
        ctx.record(evs[1])
        ctx.sync()
        ctx.record(evs[2])
        ctx.copy(out[0], d_tmp, 128 * ne)
        ctx.record(evs[3])
        ctx.sync()
        h2d_ms, d2h_ms = ctx.elapsed_ms(evs[0], evs[1]), ctx.elapsed_ms(evs[2], evs[3])
        e2e.update(h2d_ms_alone=h2d_ms, h2d_gbs_alone=32 * ne / h2d_ms / 1e6, d2h_tangent_ms_alone=d2h_ms,
                   d2h_gbs_alone=128 * ne / d2h_ms / 1e6, ms_per_step=1e3 * dt / Ke,
                   bound="PCIe device-to-host: the result is 168-188 B per point, the kernel needs < 10 % of the step")
        d_tmp.free()

    e2e_dc = None
    if model in ("step", "action"):
        # end to end through QuadratureForms with HOST vectors: only DOF vectors cross the link
        nd = 2 * tab.n_dofs
        u_h, b_h = ctx.pinned_empty((nd,)), ctx.pinned_empty((nd,))
        d_u.to_host(u_h)
        if model == "step":
            call = lambda: forms.vm_residual(vm, u_h, out=b_h, exact=args.fused_exact)  # noqa: E731
            api = "QuadratureForms.vm_residual(vm, u_host, out=b_host): eo_form_vm_step, history / tangent resident in HBM"
        else:
            call = lambda: forms.action("mandel_strain", "mandel_strain", forms.C_tang, u_h, out=b_h)  # noqa: E731
            api = "QuadratureForms.action(..., C_tang_resident, x_host, out=y_host): eo_form_action"
        Ke = max(2, min(K, 5))
        for _ in range(2):
            call()
        barrier()
        t0 = time.perf_counter()
        checksum = 0.0
        for _ in range(Ke):
            call()  # H2D + kernel + D2H, synchronous on return
            checksum += float(b_h[0])
        ctx.sync()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * n * Ke / dt, "unit": "QP/s", "h2d_bytes_per_step": 8 * nd, "d2h_bytes_per_step": 8 * nd,
               "qp_per_step_per_gpu": n, "steps": Ke, "api": api, "ms_per_step": 1e3 * dt / Ke}
    elif model == "vm" and args.e2e_n > 0 and not args.no_device_consumers:
        e2e_dc = e2e_device_consumers(ctx, eo, inputs, int(args.e2e_n), rank, world, max_over_ranks, barrier)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline for the dominant kernel (the only kernel in a step)
    k_ms = float(np.mean(kernel_ms))
    tj = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            tj = json.load(fh).get(model)
    traffic = tj.get("dram_bytes_per_launch") if (tj and tj.get("n") == n) else None
    hbm_peak, hbm_src = _measured_peaks()
    hbm_achieved = BYTES_PER_QP[model] * n / (k_ms * 1e-3) / 1e9
    if model == "mc":
        # FP64-pipe bound: flops per QP counted by ncu (DFMA = 2, DADD/DMUL = 1) for this input family,
        # peak = DFMA micro-benchmark measured now on this GPU
        fp64_peak = ctx.fp64_peak_tflops()
        flops_per_qp = tj.get("fp64_flops_per_qp") if tj else None
        ach = flops_per_qp * n / (k_ms * 1e-3) / 1e12 if flops_per_qp else None
        roofline = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": (ach / fp64_peak) if ach else None, "traffic": traffic,
                    "peak_source": "measured now: eo_fp64_peak DFMA micro-benchmark (8 chains/thread)",
                    "fp64_flops_per_qp": flops_per_qp, "kernel_ms": k_ms,
                    "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                            "bytes_per_qp": BYTES_PER_QP[model]}}
    elif model == "isihara":
        # float32 CUDA-core bound: ~21 k FMA per point in five 64x64 matrix-vector products (isihara_core.cuh)
        fma_per_qp = 5 * 64 * 64 + 64 * 16
        ach = 2.0 * fma_per_qp * n / (k_ms * 1e-3) / 1e12
        f32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        roofline = {"bound": "fp32", "achieved": ach, "peak": f32_peak, "unit": "TFLOP/s", "frac": ach / f32_peak,
                    "traffic": traffic, "peak_source": "nominal: 148 SM x 128 FMA/clk x 2 x 1.965 GHz (no measured f32 figure)",
                    "fma_per_qp": fma_per_qp, "kernel_ms": k_ms,
                    "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                            "bytes_per_qp": BYTES_PER_QP[model]}}
    else:
        roofline = {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_achieved / hbm_peak, "traffic": traffic, "peak_source": hbm_src,
                    "bytes_per_qp": BYTES_PER_QP[model], "kernel_ms": k_ms}
        if model in GATHERED_BYTES_PER_QP:
            # the gather-limited kernels report the unique-byte fraction (above: what must cross HBM once)
            # and the gathered-byte fraction (what the threads request: every cell's own copy of its dofs / vertices)
            gb = GATHERED_BYTES_PER_QP[model]
            ga = gb * n / (k_ms * 1e-3) / 1e9
            roofline["gathered"] = {"bytes_per_qp": gb, "achieved": ga, "unit": "GB/s", "frac_of_hbm_peak": ga / hbm_peak,
                                    "note": "requested bytes (per-cell copies of shared dofs / vertices served by L1/L2)"}

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if world > 1:
        cpu = None  # the CPU baseline is reported at N = 1 only (the reference arm is timed separately at every N)
    elif model == "isihara":
        cpu = None  # the CPU implementation is the reference's torch code, which needs /root/reference: timed in the
        #             build container only (69 k QP/s on 8 threads)
    elif args.cpu_seconds > 0 and model in ("tab", "fused", "jitfused", "step", "action"):
        cpu = cpu_tab_rate({"jitfused": "fused"}.get(model, model), args.cpu_seconds)
    elif args.cpu_seconds > 0:
        sample = int(args.cpu_sample) if model != "mc" else min(int(args.cpu_sample), 200_000)
        cm = "vm" if model in ("jitvm", "jitvm3d") else model
        rate, cores, passes = cpu_port_rate(cm, sample, args.cpu_seconds, parallel=True)
        rate1, _, _ = cpu_port_rate(cm, sample // 4, min(3.0, args.cpu_seconds), parallel=False)
        what = {"vm": "C restatement of the reference's Numba kernel (serial in the reference)",
                "jitvm": "C restatement of the reference's Numba kernel (serial in the reference)",
                "jitvm3d": "C restatement of the reference's plane-strain Numba kernel (no 3-D CPU implementation exists)",
                "heat": "C restatement of the reference's NumPy functions",
                "mc": "C++ nested-dual-number restatement of the reference's JAX program (JAX not installable offline)"}
        cpu = {"value": rate, "unit": "QP/s", "cores": cores, "kind": "port",
               "sample": f"{sample} QPs x {passes} passes (best pass); {what[model]}, OpenMP; single-thread rate "
                         f"{rate1:.3e} QP/s",
               "single_thread_value": rate1}

    hist = stats["niter_hist"]
    cfg = {
        "workload": WORKLOADS[model], "qp_per_gpu": n, "state_layout": args.state_layout,
        "plastic_fraction": stats["n_plastic"] / max(stats["n_points"], 1),
        "l2": f"inputs+outputs {BYTES_PER_QP[model] * n / 1e9:.1f} GB per step >> 126 MB L2 (no flush needed)",
        "partition": "contiguous block of QPs per rank, no halo; one statistics collective per step when n_gpus > 1 "
                     "(plastic_fraction / niter histogram are then the global figures)",
    }
    if model == "mc":
        tot = max(int(hist.sum()), 1)
        cfg["niter_histogram"] = {int(i): float(hist[i]) / tot for i in np.nonzero(hist)[0]}
        cfg["n_nonconverged"] = stats["n_nonconverged"]
    cfg.update(extra_cfg)
    if world > 1:
        cfg["cpu_binding"] = (f"rank pinned to the {len(numa_cores)} cores local to its GPU (NVML affinity)" if numa_cores
                              else "none")
    line = {
        "metric": METRIC, "value": value, "unit": "QP/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if e2e_dc is not None:
        line["e2e_device_consumers"] = e2e_dc
    if collective_check is not None:
        line["collective"] = collective_check
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="vm", choices=["vm", "heat", "mc", "tab", "fused", "isihara", "jitvm", "jitfused", "jitvm3d", "step", "action"])
    ap.add_argument("--fused-exact", action="store_true")
    ap.add_argument("--mc-scheme", default="queue", choices=["queue", "simple", "queue-noaffinity", "queue-onepass"])
    ap.add_argument("--n", "--qp-per-gpu", dest="n", type=float, default=1e8,
                    help="quadrature points per GPU (device-resident leg); under torchrun spell it --qp-per-gpu "
                         "(torchrun's own parser rejects --n as an ambiguous abbreviation)")
    ap.add_argument("--e2e-n", type=float, default=1.5e7, help="quadrature points per GPU for the end-to-end leg")
    ap.add_argument("--state-layout", default="aos", choices=["aos", "soa"])
    ap.add_argument("--cpu-sample", type=float, default=4e6)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-device-consumers", action="store_true",
                    help="vm: skip the supplementary end-to-end leg through the device-side consumers (QuadratureForms)")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-GPU: do not pin ranks to their GPU's NUMA node")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1 and args.impl == "b200":
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
