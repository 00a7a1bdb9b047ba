#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 quadrature-point engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model vm|heat|mc|...] [--n QP_PER_GPU]
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
             box's host cores, bounded sample; plus the serial Numba kernel itself (`numba_serial`).
  models     (default line only) every other kernel of the hot path timed the same way in the same run - Mohr-Coulomb,
             heat, operand tabulation, fused tabulate + von Mises, residual step, tangent action, Isihara - each with
             kernel_ms, its roofline fraction (HBM, or the FP64 / FP32 pipe peak measured in this run) and a short CPU
             leg; `peaks` carries the measured denominators.
  e2e_device_consumers  the same constitutive update consumed on the device (host DOF vector in, host residual out)
             with its own CPU baseline (the C cell loop: tabulate + radial return + residual scatter).
One JSON line on stdout (rank 0).
"""

from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_QP = {"vm": 240, "jitvm": 240, "jitfused": 235, "jitvm3d": 448, "heat": 88, "mc": 252, "tab": 176.0 / 3.0, "fused": 235, "isihara": 192,
                # device-side consumers (csrc/form.cu): + read-modify-write of the DOF vector (2 nodes x 2 x 8 B per cell, twice)
                "step": 235 + 64.0 / 3.0, "action": 128 + 80.0 / 3.0 + 64.0 / 3.0,
                # the same with the von Mises tangent kept as 6 numbers per point (48 instead of 128 B)
                "step6": 235 - 80 + 64.0 / 3.0, "action6": 48 + 80.0 / 3.0 + 64.0 / 3.0}
# per P2-triangle cell: dofmap 24 + x_dofmap 12 + 6 gathered dofs x 16 + 3 gathered vertices x 16 = 180 B (vs 80 B unique)
_CELL_GATHER = (24 + 12 + 96 + 48) / 3.0
GATHERED_BYTES_PER_QP = {"tab": _CELL_GATHER + 32, "fused": _CELL_GATHER + 40 + 168, "jitfused": _CELL_GATHER + 40 + 168,
                         "step": _CELL_GATHER + 40 + 168 + 12 * 16 / 3.0, "action": _CELL_GATHER + 128 + 12 * 16 / 3.0,
                         "step6": _CELL_GATHER + 40 + 88 + 12 * 16 / 3.0, "action6": _CELL_GATHER + 48 + 12 * 16 / 3.0}
# ALGORITHMIC FP64 work of the Mohr-Coulomb path per point for the demo stress-path family (34 % plastic points, 2-5
# Newton iterations): 2 x DFMA + DADD + DMUL thread instructions counted by ncu: yield test 245 +
# Newton / tangent recursion 1768.  Fixed: a faster kernel that needs fewer instructions does not shrink it.
MC_FLOPS_PER_QP = 2014.0
# FP32 FMAs of the Isihara network per point: five 64x64 matrix-vector products + the 3->64 / 64->1 layers (isihara_core.cuh)
ISIHARA_FMA_PER_QP = 5 * 64 * 64 + 64 * 16
MESH_MODELS = ("tab", "fused", "jitfused", "step", "action", "step6", "action6")
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


# ------------------------------------------------------------------------------- CPU legs (oracle/ = the checker, timed)
def _best_of(fn, min_seconds: float, min_passes: int = 3):
    fn()  # warm-up (page faults, thread pool, JIT)
    best, passes, t_all = float("inf"), 0, time.perf_counter()
    while True:
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
        passes += 1
        if time.perf_counter() - t_all >= min_seconds and passes >= min_passes:
            break
    return best, passes


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
    best, passes = _best_of(fn, min_seconds)
    return sample_n / best, cores, passes


def cpu_numba_rate(sample_n: int, min_seconds: float):
    """The reference's own CPU callable as it executes: serial @numba.njit loop (oracle/numba_vm.py, bit-identical to
    the golden made by the reference's kernel).  None when numba is not importable."""
    try:
        from oracle import inputs, numba_vm

        f = numba_vm.make_return_mapping()
    except Exception:  # numba absent / unusable: reported as unavailable, never fatal
        return None
    deps, sigma_n, p = inputs.vm_batch(sample_n, seed=0)
    best, passes = _best_of(lambda: f(deps, sigma_n, p), min_seconds, 2)
    return {"value": sample_n / best, "unit": "QP/s", "cores": 1, "kind": "port",
            "sample": f"{sample_n} QPs x {passes} passes (best pass): serial @numba.njit loop with per-point NumPy algebra and "
                      "three fresh result arrays per call - the execution model of the reference's return_mapping "
                      "(demo_plasticity_von_mises.py:298-332), restated in oracle/numba_vm.py and bit-identical to the golden made "
                      "by the reference's kernel"}


def cpu_isihara_rate(sample_n: int, min_seconds: float):
    """The reference's torch evaluation (vmap(jacfwd(grad)) through the ICNN), restated in oracle/isihara_torch.py and
    pinned to the golden made by the reference's module; eager ATen on torch.get_num_threads() threads."""
    import torch

    from oracle import inputs
    from oracle import isihara_torch as it

    try:
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, RuntimeError):
        pass
    g = np.load(os.path.join(ROOT, "tests", "golden", "isihara_seed0_n2049.npz"))
    sd = {k[3:]: g[k] for k in g.files if k.startswith("sd/")}
    F = inputs.isihara_batch(sample_n, seed=0)
    best, passes = _best_of(lambda: it.dP_dF(F, sd, g["H_flat"]), min_seconds, 2)
    return {"value": sample_n / best, "unit": "QP/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_n} QPs x {passes} passes (best pass): torch.func vmap(jacfwd(grad)) through the float32 ICNN "
                      "(demo_hyperelasticity.py:429-456), functional restatement pinned to the reference's golden"}


def cpu_tab_rate(model: str, min_seconds: float, nxy: int = 800):
    """CPU baseline for the cell-loop legs (tab / fused / step / action): the OpenMP C restatement of the von Mises demo's
    cell loop (oracle/csrc/forms_oracle.c: tabulation of the Mandel strain, radial return, residual scatter, tangent
    action) on all host cores, on a bounded mesh (nxy = 800: 1.28 M cells = 3.84 M quadrature points)."""
    from dolfinx_external_operator_b200 import elements as el
    from dolfinx_external_operator_b200 import synthetic as syn
    from oracle import constitutive as oc
    from oracle import native

    cores = native.use_all_cores()
    m = syn.triangle_mesh(nxy, nxy, 2, jitter=0.2, seed=0)
    m["phi"], m["dphi"] = el.lagrange_triangle(2, el.triangle_quadrature(2))
    m["dpsi"] = el.p1_geometry_derivatives(2)
    u = syn.smooth_displacement(m["dof_coords"], scale=1.5e-3, seed=0).reshape(-1)
    nq = 3 * m["dofmap"].shape[0]
    _, sn, p = syn.vm_batch(nq, seed=0)
    W3 = el.triangle_quadrature_weights(2)
    prm = oc.VonMisesParams()
    mode = {"tab": "tab", "fused": "fused", "jitfused": "fused", "step": "step", "action": "action", "step6": "step",
            "action6": "action"}[model]
    Ct0 = native.forms_p2_cells("fused", m, W3, u, prm, sn, p)[0] if mode == "action" else None

    def fn():
        if mode == "action":
            native.forms_p2_cells("action", m, W3, u, C_tang=Ct0)
        else:
            native.forms_p2_cells(mode, m, W3, u, prm, sn, p)

    best, passes = _best_of(fn, min_seconds)
    what = {"tab": "tabulation of the Mandel strain", "fused": "tabulation + von Mises radial return",
            "step": "tabulation + von Mises radial return + residual scatter",
            "action": "tangent action (tabulate x, contract with the stored tangent, scatter)"}[mode]
    return {"value": nq / best, "unit": "QP/s", "cores": cores, "kind": "port",
            "sample": f"{nq} QPs x {passes} passes (best pass, includes allocating the result arrays); OpenMP C restatement "
                      f"of the demo's cell loop ({what}; the reference runs it serially per MPI rank through DOLFINx/FFCx "
                      "and Numba)"}


def cpu_leg(model: str, seconds: float, sample: float | None = None):
    """The CPU baseline object of one bench leg (bounded sample, `seconds` of timed work at least)."""
    if model == "isihara":
        return cpu_isihara_rate(int(sample or 4096), seconds)
    if model in MESH_MODELS:
        return cpu_tab_rate(model, seconds, 800 if seconds >= 5 else 400)
    cm = "vm" if model in ("jitvm", "jitvm3d") else model
    n_s = int(sample or 4e6)
    if model == "mc":
        n_s = min(n_s, 200_000 if seconds >= 5 else 40_000)
    rate, cores, passes = cpu_port_rate(cm, n_s, seconds, parallel=True)
    what = {"vm": "C restatement of the reference's Numba kernel (serial in the reference)",
            "jitvm": "C restatement of the reference's Numba kernel (serial in the reference)",
            "jitvm3d": "C restatement of the reference's plane-strain Numba kernel (no 3-D CPU implementation exists)",
            "heat": "C restatement of the reference's NumPy functions",
            "mc": "C++ nested-dual-number restatement of the reference's JAX program (JAX not installable offline)"}
    return {"value": rate, "unit": "QP/s", "cores": cores, "kind": "port",
            "sample": f"{n_s} QPs x {passes} passes (best pass); {what[model]}, OpenMP"}


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
    if args.model in ("tab", "fused", "step", "action", "step6", "action6", "isihara"):
        r = cpu_tab_rate(args.model, 5.0) if args.model != "isihara" else cpu_isihara_rate(8192, 5.0)
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
    if args.model == "vm" and args.cpu_seconds > 0:
        nb = cpu_numba_rate(min(sample, 1_000_000), 3.0)
        if nb is not None:
            line["cpu_baseline"]["numba_serial"] = nb
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
           "(evaluate_operands for the von Mises / Mohr-Coulomb demos), jittered triangle mesh",
    "fused": "operand tabulation fused with the von Mises return mapping (strain never stored), P2 vector field, "
             "3 quadrature points per triangle",
    "step": "one Newton residual evaluation on the device (SURVEY.md 8f rank 1): Mandel strain of a P2 vector field -> "
            "von Mises return mapping (tangent / stress / dp stored in HBM) -> b = int sigma . epsilon(v) dx, ONE kernel; "
            "3 quadrature points per triangle",
    "action": "matrix-free tangent action y = J x with J = int (C_tang epsilon(u_hat)) . epsilon(v) dx, tangent resident "
              "in HBM (what a Krylov method needs from assemble_matrix); P2 vector field, 3 points per triangle",
    "step6": "the residual step with the von Mises tangent kept in FACTORED form for the device-side consumers (6 numbers per "
             "point: C_t = C_elas - cn v v^T - cd dev; 48 instead of 128 B per point written): eo_form_vm_step_factored; "
             "stress, dp, statistics and b identical to `step`",
    "action6": "the matrix-free tangent action with C_t rebuilt in registers from the factored tangent (48 instead of 128 B "
               "per point read): eo_form_action_vm_factored; agrees with `action` to rounding",
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


def build_mesh(ctx, eo, inputs, n, rank, order="structured"):
    """P2 vector triangle mesh with ~n quadrature points (3 per cell), resident tabulator + forms + displacement."""
    from dolfinx_external_operator_b200 import elements as el

    nxy = max(2, int(round((n / 6.0) ** 0.5)))
    mesh = inputs.triangle_mesh(nxy, nxy, 2, jitter=0.2, seed=rank)
    if order != "structured":
        mesh = inputs.renumber(mesh, order, seed=rank)
    n_cells = mesh["dofmap"].shape[0]
    phi, dphi = el.lagrange_triangle(2, el.triangle_quadrature(2))
    tab = eo.Tabulator(dofmap=mesh["dofmap"], x_dofmap=mesh["x_dofmap"], x=mesh["x"], phi=phi, dphi=dphi, bs=2,
                       n_dofs=mesh["n_dofs"], ctx=ctx)
    d_u = ctx.to_device(inputs.smooth_displacement(mesh["dof_coords"], scale=1.5e-3, seed=rank).reshape(-1))
    forms = eo.QuadratureForms(tab, el.triangle_quadrature_weights(2))
    cfg = {"n_cells": n_cells, "n_dofs": mesh["n_dofs"], "element": "P2 vector triangle, 3-point rule", "mesh_order": order}
    return SimpleNamespace(tab=tab, forms=forms, d_u=d_u, n_cells=n_cells, n=3 * n_cells, cfg=cfg, vm=None)


def build_workload(model, ctx, eo, inputs, n, rank, args, mesh=None):
    """Allocate the device-resident synthetic inputs of one leg (a seeded tile, seed = rank, repeated to n points) and
    return a namespace with `step()` (one pass of the hot path over the batch), the point count `n`, config extras and
    the objects the end-to-end legs reuse."""
    import ctypes as C

    w = SimpleNamespace(model=model, n=int(n), cfg={}, keep=[])
    tile_n = min(w.n, 1 << 22)
    if model == "vm":
        vm = eo.VonMises(n_qp=w.n, ctx=ctx, state_layout=args.state_layout)
        deps_t, sn_t, p_t = inputs.vm_batch(tile_n, seed=rank)
        d_deps = ctx.empty((w.n * 4,))
        _tile_to_device(ctx, d_deps, deps_t, w.n, 4)
        if args.state_layout == "soa":
            for c in range(4):
                col = ctx.to_device(np.ascontiguousarray(sn_t[:, c]))
                for r in range(0, w.n, tile_n):
                    m = min(tile_n, w.n - r)
                    ctx.copy(vm.sigma_n_dev.ptr + (c * w.n + r) * 8, col, m * 8)
                ctx.sync()
                col.free()
        else:
            _tile_to_device(ctx, vm.sigma_n_dev, sn_t, w.n, 4)
        _tile_to_device(ctx, vm.p_dev, p_t, w.n, 1)
        d_Ct = ctx.empty((w.n * 16,))
        w.keep += [vm, d_deps, d_Ct]
        w.step = lambda: vm.eval_device(d_deps, d_Ct)
    elif model in ("jitvm", "jitvm3d"):
        from dolfinx_external_operator_b200 import jit_models as jm

        nc = 6 if model == "jitvm3d" else 4
        jv = jm.von_mises_3d(ctx=ctx) if model == "jitvm3d" else jm.von_mises(ctx=ctx)
        if model == "jitvm3d":
            rng = np.random.default_rng(rank)
            deps_t, sn_t = rng.normal(0.0, 2e-3, (tile_n, 6)), rng.normal(0.0, 100.0, (tile_n, 6))
            p_t = np.abs(rng.normal(0.0, 1e-3, tile_n))
        else:
            deps_t, sn_t, p_t = inputs.vm_batch(tile_n, seed=rank)
        d_deps = ctx.empty((w.n * nc,))
        _tile_to_device(ctx, d_deps, deps_t, w.n, nc)
        jv.state = [ctx.empty((w.n * nc,)), ctx.empty((w.n,))]
        _tile_to_device(ctx, jv.state[0], sn_t, w.n, nc)
        _tile_to_device(ctx, jv.state[1], p_t, w.n, 1)
        d_Ct, d_sig, d_dp = ctx.empty((w.n * nc * nc,)), ctx.empty((w.n * nc,)), ctx.empty((w.n,))
        jv.compile((1,))
        w.keep += [jv, d_deps, d_Ct, d_sig, d_dp]
        w.step = lambda: jv.eval_device((1,), [d_deps], d_Ct, d_sig, [d_dp])
    elif model == "mc":
        from dolfinx_external_operator_b200._lib import McParams

        mc = eo.MohrCoulomb(ctx=ctx, history=None)
        tile_n = min(w.n, 1 << 20)
        # the stress paths are walked with the GPU kernel itself as the stress update
        deps_t, sn_t = inputs.mc_batch(tile_n, seed=rank, stepper=mc.stress_update)
        d_deps, d_sn = ctx.empty((w.n * 4,)), ctx.empty((w.n * 4,))
        _tile_to_device(ctx, d_deps, deps_t, w.n, 4)
        _tile_to_device(ctx, d_sn, sn_t, w.n, 4)
        d_Ct, d_sig = ctx.empty((w.n * 16,)), ctx.empty((w.n * 4,))
        d_it = ctx.empty((w.n,), np.int32)
        d_yl, d_nr, d_dl = ctx.empty((w.n,)), ctx.empty((w.n,)), ctx.empty((w.n,))
        prm = McParams(mc.E, mc.nu, mc.c, mc.phi, mc.psi, mc.theta_T, mc.a, mc.tol, mc.Nitermax)
        scheme = {"classes": 0, "simple": 1, "queue-noaffinity": 2, "queue-onepass": 3, "ring": 4}[args.mc_scheme]
        w.cfg["mc_scheme"] = args.mc_scheme
        w.deps_t, w.sn_t = deps_t, sn_t
        w.keep += [mc, d_deps, d_sn, d_Ct, d_sig, d_it, d_yl, d_nr, d_dl]
        n_ = w.n
        w.step = lambda: ctx.check(ctx.lib.eo_mc_eval_scheme(ctx.handle, C.byref(prm), d_deps.ptr, d_sn.ptr, d_Ct.ptr, d_sig.ptr,
                                                             d_it.ptr, d_yl.ptr, d_nr.ptr, d_dl.ptr, n_, scheme))
    elif model == "isihara":
        gpath = os.path.join(ROOT, "tests", "golden", "isihara_seed0_n2049.npz")  # carries the reference's state dict
        g = np.load(gpath)
        isi = eo.Isihara({k[3:]: g[k] for k in g.files if k.startswith("sd/")}, ctx=ctx)
        F_t = inputs.isihara_batch(tile_n, seed=rank)
        d_F = ctx.empty((w.n * 4,))
        _tile_to_device(ctx, d_F, F_t, w.n, 4)
        d_dP, d_P = ctx.empty((w.n * 16,)), ctx.empty((w.n * 4,))
        w.isi, w.F_t = isi, F_t
        w.keep += [d_F, d_dP, d_P]
        w.step = lambda: isi.eval_device(d_F, d_dP, d_P)
    elif model in MESH_MODELS:
        M = mesh if mesh is not None else build_mesh(ctx, eo, inputs, w.n, rank, args.mesh_order)
        w.mesh = M
        w.n = M.n
        w.cfg.update(M.cfg)
        tab, forms, d_u = M.tab, M.forms, M.d_u
        tile_n = min(w.n, 1 << 22)
        if model == "tab":
            d_out = ctx.empty((M.n_cells, 3, 4))
            w.keep.append(d_out)
            w.step = lambda: tab.evaluate("mandel_strain", d_u, out=d_out)
        elif model == "jitfused":
            from dolfinx_external_operator_b200 import jit_models as jm
            from dolfinx_external_operator_b200.tabulation import LazyOperand

            jv = jm.von_mises(ctx=ctx)
            _, sn_t, p_t = inputs.vm_batch(tile_n, seed=rank)
            jv.state = [ctx.empty((w.n * 4,)), ctx.empty((w.n,))]
            _tile_to_device(ctx, jv.state[0], sn_t, w.n, 4)
            _tile_to_device(ctx, jv.state[1], p_t, w.n, 1)
            dev = {"out": ctx.empty((w.n * 16,)), "value": ctx.empty((w.n * 4,)), "aux0": ctx.empty((w.n,))}
            lz = [LazyOperand(tab, 2, d_u)]
            dd = jv._deriv((1,))[1]
            w.keep += [jv, dev]
            w.step = lambda: jv._evaluate_fused((1,), dd, 16, lz, dev)
        else:
            vm = M.vm
            if vm is None:
                vm = M.vm = eo.VonMises(n_qp=w.n, ctx=ctx)
                _, sn_t, p_t = inputs.vm_batch(tile_n, seed=rank)
                _tile_to_device(ctx, vm.sigma_n_dev, sn_t, w.n, 4)
                _tile_to_device(ctx, vm.p_dev, p_t, w.n, 1)
            w.vm = vm
            w.cfg["vm_arithmetic"] = "exact" if args.fused_exact else "fast (2 divisions, FMA; identical flags)"
            if model == "fused":
                if forms.C_tang is None or forms.C_tang.size != 16 * w.n:
                    forms.C_tang = ctx.empty((16 * w.n,))
                w.step = lambda: tab.vm_fused(vm, d_u, C_tang=forms.C_tang, exact=args.fused_exact)
            else:
                w.cfg["scatter"] = "fp64 RED.ADD per element-vector entry (12 per cell)"
                d_b = ctx.empty((2 * tab.n_dofs,))
                w.keep.append(d_b)
                if model == "step":
                    w.step = lambda: forms.vm_residual(vm, d_u, out=d_b, exact=args.fused_exact)
                elif model == "step6":
                    w.step = lambda: forms.vm_residual(vm, d_u, out=d_b, exact=args.fused_exact, tangent="factored")
                elif model == "action6":
                    forms.vm_residual(vm, d_u, out=d_b, exact=args.fused_exact, tangent="factored")  # fills forms.T6
                    d_y = ctx.empty((2 * tab.n_dofs,))
                    w.keep.append(d_y)
                    w.step = lambda: forms.vm_action(d_u, out=d_y)
                else:
                    forms.vm_residual(vm, d_u, out=d_b, exact=args.fused_exact)  # fills forms.C_tang
                    d_y = ctx.empty((2 * tab.n_dofs,))
                    w.keep.append(d_y)
                    w.step = lambda: forms.action("mandel_strain", "mandel_strain", forms.C_tang, d_u, out=d_y)
    elif model == "heat":
        T_t, s_t = inputs.heat_batch(tile_n, seed=rank)
        d_T, d_s = ctx.empty((w.n,)), ctx.empty((w.n * 2,))
        _tile_to_device(ctx, d_T, T_t, w.n, 1)
        _tile_to_device(ctx, d_s, s_t, w.n, 2)
        d_q, d_dT, d_ds = ctx.empty((w.n * 2,)), ctx.empty((w.n * 2,)), ctx.empty((w.n * 4,))
        w.keep += [d_T, d_s, d_q, d_dT, d_ds]
        n_ = w.n
        w.step = lambda: ctx.check(ctx.lib.eo_heat_eval(ctx.handle, 1.0, 1.0, d_T.ptr, d_s.ptr, None, None, d_q.ptr, d_dT.ptr,
                                                        d_ds.ptr, n_))
    else:
        raise ValueError(model)
    return w


def time_steps(ctx, step, K, W, barrier, max_over_ranks, collective=None, sampler=None):
    """W untimed warm-up steps, then K steps bracketed by barrier + sync; CUDA events on the library's compute stream
    around the whole region and around every step's kernel(s).  Returns (total_ms max over ranks, [kernel_ms], launches)."""
    for _ in range(W):
        step()
        if collective:
            collective()
    ctx.sync()
    ctx.stats_reset()
    ev = [ctx.event() for _ in range(2 * K + 2)]
    barrier()
    ctx.sync()
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count
    ctx.record(ev[0])
    for k in range(K):
        ctx.record(ev[2 + 2 * k])
        step()
        ctx.record(ev[3 + 2 * k])
        if collective:
            collective()
    ctx.record(ev[1])
    ctx.sync()  # all streams of the context, the collective stream included
    barrier()
    total_ms = max_over_ranks(ctx.elapsed_ms(ev[0], ev[1]))
    kernel_ms = [ctx.elapsed_ms(ev[2 + 2 * k], ev[3 + 2 * k]) for k in range(K)]
    time_steps.kernel_ms_max = max_over_ranks(float(np.mean(kernel_ms)))  # slowest rank's mean kernel time
    for e in ev:
        ctx.lib.eo_event_destroy(ctx.handle, e)
    return total_ms, kernel_ms, ctx.launch_count - launches0


def roofline_for(model, n, k_ms, peaks, traffic=None):
    hbm_peak, hbm_src = peaks["hbm_gbs"], peaks["hbm_source"]
    hbm_achieved = BYTES_PER_QP[model] * n / (k_ms * 1e-3) / 1e9
    hbm = {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
           "bytes_per_qp": BYTES_PER_QP[model]}
    if model == "mc":
        ach = MC_FLOPS_PER_QP * n / (k_ms * 1e-3) / 1e12
        return {"bound": "fp64", "achieved": ach, "peak": peaks["fp64_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["fp64_tflops"], "traffic": traffic,
                "peak_source": "measured in this run: eo_fp64_peak DFMA micro-benchmark (8 chains/thread, all SMs)",
                "fp64_flops_per_qp": MC_FLOPS_PER_QP, "kernel_ms": k_ms, "hbm": hbm}
    if model == "isihara":
        ach = 2.0 * ISIHARA_FMA_PER_QP * n / (k_ms * 1e-3) / 1e12
        return {"bound": "fp32", "achieved": ach, "peak": peaks["fp32_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["fp32_tflops"], "traffic": traffic,
                "peak_source": "measured in this run: eo_fp32_peak FFMA micro-benchmark (8 chains/thread, all SMs)",
                "fma_per_qp": ISIHARA_FMA_PER_QP, "kernel_ms": k_ms, "hbm": hbm}
    r = {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
         "traffic": traffic, "peak_source": hbm_src, "bytes_per_qp": BYTES_PER_QP[model], "kernel_ms": k_ms}
    if model in GATHERED_BYTES_PER_QP:
        # the gather-limited kernels report the unique-byte fraction (above: what must cross HBM once)
        # and the gathered-byte fraction (what the threads request: every cell's own copy of its dofs / vertices)
        gb = GATHERED_BYTES_PER_QP[model]
        ga = gb * n / (k_ms * 1e-3) / 1e9
        r["gathered"] = {"bytes_per_qp": gb, "achieved": ga, "unit": "GB/s", "frac_of_hbm_peak": ga / hbm_peak,
                         "note": "requested bytes (per-cell copies of shared dofs / vertices served by L1/L2)"}
    return r


def _traffic(model, n):
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None
    with open(tpath) as fh:
        tj = json.load(fh).get(model)
    return tj.get("dram_bytes_per_launch") if (tj and tj.get("n") == n) else None


def _stats_cfg(model, stats):
    out = {"plastic_fraction": stats["n_plastic"] / max(stats["n_points"], 1)}
    if model == "mc":
        hist = stats["niter_hist"]
        tot = max(int(hist.sum()), 1)
        out["niter_histogram"] = {int(i): float(hist[i]) / tot for i in np.nonzero(hist)[0]}
        out["n_nonconverged"] = stats["n_nonconverged"]
    return out


def models_block(ctx, eo, inputs, args, rank, world, peaks, barrier, max_over_ranks, collective):
    """Every other kernel of the hot path, timed like the headline leg (W >= 3 warm-ups, K steps, CUDA events, max over
    ranks, the statistics collective after every step when n_gpus > 1), one after the other with its buffers freed in
    between.  Inputs + outputs of every leg are far larger than L2 (no flush needed)."""
    K, W = args.models_steps, 3
    out = {}
    names = [m for m in args.models.split(",") if m]
    n_for = {"isihara": min(args.n, int(args.models_isihara_n))}
    mesh_names = [m for m in names if m in MESH_MODELS]
    mesh = None
    for name in names:
        if name in MESH_MODELS and mesh is None:
            mesh = build_mesh(ctx, eo, inputs, int(args.n), rank, args.mesh_order)
        wl = build_workload(name, ctx, eo, inputs, n_for.get(name, args.n), rank, args, mesh=mesh)
        total_ms, k_ms, launches = time_steps(ctx, wl.step, K, W, barrier, max_over_ranks, collective)
        stats = ctx.stats()
        km = float(np.mean(k_ms))
        entry = {"workload": WORKLOADS[name], "qp_per_gpu": wl.n, "steps": K, "warmup": W, "kernel_ms": km,
                 "kernel_ms_max_over_ranks": time_steps.kernel_ms_max,
                 "ms_per_step": total_ms / K, "value": world * wl.n * K / (total_ms * 1e-3), "unit": "QP/s",
                 "gpu_launches": int(launches), "roofline": roofline_for(name, wl.n, km, peaks, _traffic(name, wl.n))}
        if name in ("mc", "fused", "step", "step6"):
            entry.update(_stats_cfg(name, stats))
        entry.update(wl.cfg)
        if rank == 0 and world == 1 and args.models_cpu_seconds > 0:
            entry["cpu_baseline"] = cpu_leg(name, args.models_cpu_seconds)
        out[name] = entry
        del wl
        if name in MESH_MODELS and name == mesh_names[-1]:
            mesh = None
        gc.collect()
    return out


def e2e_device_consumers(ctx, eo, inputs, ne, rank, world, max_over_ranks, barrier, cpu_seconds):
    """Supplementary end-to-end figure of the default (von Mises) line: the same constitutive update consumed ON the
    device - host displacement vector in, host residual vector out, tangent resident for the
    matrix-free action - instead of shipping 168 B per point back to DOLFINx's assembler."""
    M = build_mesh(ctx, eo, inputs, ne, rank)
    tab, forms, nq = M.tab, M.forms, M.n
    vm = eo.VonMises(n_qp=nq, ctx=ctx)
    _, sn_t, p_t = inputs.vm_batch(min(nq, 1 << 22), seed=rank)
    _tile_to_device(ctx, vm.sigma_n_dev, sn_t, nq, 4)
    _tile_to_device(ctx, vm.p_dev, p_t, nq, 1)
    nd = 2 * tab.n_dofs
    u_h, b_h, y_h = ctx.pinned_empty((nd,)), ctx.pinned_empty((nd,)), ctx.pinned_empty((nd,))
    M.d_u.to_host(u_h)
    out = {}
    for name, call in (("residual", lambda: forms.vm_residual(vm, u_h, out=b_h)),
                       ("tangent_action", lambda: forms.action("mandel_strain", "mandel_strain", forms.C_tang, u_h, out=y_h))):
        for _ in range(3):
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
    if rank == 0 and world == 1 and cpu_seconds > 0:
        # the reference-side counterpart of `residual`: host u in, host b out (+ tangent / stress / dp in host memory)
        out["residual"]["cpu_baseline"] = cpu_tab_rate("step", min(cpu_seconds, 5.0))
        out["tangent_action"]["cpu_baseline"] = cpu_tab_rate("action", min(cpu_seconds, 3.0), 400)
    return out


def run_gpu_arm(args):
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
    model = args.model
    K, W = args.steps, args.warmup

    hbm_peak, hbm_src = _measured_peaks()
    peaks = {"hbm_gbs": hbm_peak, "hbm_source": hbm_src, "fp64_tflops": ctx.fp64_peak_tflops(),
             "fp32_tflops": ctx.fp32_peak_tflops(),
             # by instruction form: all-register scalar FFMA and packed FFMA2 sit below the
             # uniform-operand form the roofline fraction is quoted against
             "fp32_tflops_ffma_3reg": ctx.fp32_peak_tflops(variant=1), "fp32_tflops_ffma2": ctx.fp32_peak_tflops(variant=2),
             "how": "FP64 / FP32: register-only DFMA / FFMA micro-benchmarks of this library (8 independent chains per thread, "
                    "all SMs, best of 3) run at the start of this bench"}

    def collective():
        if dist is not None:
            from dolfinx_external_operator_b200.parallel import allreduce_stats_device

            allreduce_stats_device(ctx)

    wl = build_workload(model, ctx, eo, inputs, int(args.n), rank, args)
    n = wl.n
    extra_cfg = wl.cfg

    # ---- device-resident timing
    sampler = ClockSampler(local_rank)
    total_ms, kernel_ms, launches = time_steps(ctx, wl.step, K, W, barrier, max_over_ranks, collective, sampler)
    clocks = sampler.stop()
    headline_kernel_ms_max = time_steps.kernel_ms_max  # ms_per_step - this = what the per-step collective costs
    stats = ctx.stats()  # this rank's record over the K timed steps (the collective never modifies it)
    collective_check = None
    if dist is not None:
        # the K-th collective left the GLOBAL record of the K steps: it must equal the sum of the local records,
        # gathered here a second, independent way (Python objects through the object collective)
        gstats = ctx.stats_global()
        loc = [None] * world
        dist.all_gather_object(loc, {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in stats.items()})
        want_points = sum(l["n_points"] for l in loc)
        want_plastic = sum(l["n_plastic"] for l in loc)
        want_hist = np.sum([np.asarray(l["niter_hist"], dtype=np.int64) for l in loc], axis=0)
        counted = model in ("vm", "mc", "fused", "step", "step6")
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
        F_h.reshape(-1, 4)[:] = np.resize(wl.F_t, (ne, 4))
        call = wl.isi((1,))
        Ke = max(2, min(K, 5))
        for _ in range(3):
            out = call(F_h)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            out = call(F_h)
        ctx.sync()
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * ne * Ke / dt, "unit": "QP/s", "h2d_bytes_per_step": 32 * ne, "d2h_bytes_per_step": 160 * ne,
               "qp_per_step_per_gpu": ne, "steps": Ke, "api": "Isihara((1,))(F) == external_function(derivatives)(*operands)"}
        del call, out, F_h
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
            deps_t2, sn_t2 = wl.deps_t, wl.sn_t
            m_e.set_history(np.resize(sn_t2, (ne, 4)))
            d2h = (160 + 28) * ne
            api = "MohrCoulomb((1,))(deps) == external_function(derivatives)(*operands), history resident in HBM"
        for r in range(0, ne, deps_t2.shape[0]):
            m = min(deps_t2.shape[0], ne - r)
            flat[r:r + m] = deps_t2[:m]
        call = m_e((1,))
        Ke = max(2, min(K, 5))
        for _ in range(3):
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
        del m_e, call, out, deps_h, flat
    if model in ("step", "action", "step6", "action6"):
        # end to end through QuadratureForms with HOST vectors: only DOF vectors cross the link
        tab, forms, vm = wl.mesh.tab, wl.mesh.forms, wl.vm
        nd = 2 * tab.n_dofs
        u_h, b_h = ctx.pinned_empty((nd,)), ctx.pinned_empty((nd,))
        wl.mesh.d_u.to_host(u_h)
        if model == "step":
            call = lambda: forms.vm_residual(vm, u_h, out=b_h, exact=args.fused_exact)  # noqa: E731
            api = "QuadratureForms.vm_residual(vm, u_host, out=b_host): eo_form_vm_step, history / tangent resident in HBM"
        elif model == "step6":
            call = lambda: forms.vm_residual(vm, u_h, out=b_h, exact=args.fused_exact, tangent="factored")  # noqa: E731
            api = "QuadratureForms.vm_residual(vm, u_host, out=b_host, tangent='factored'): eo_form_vm_step_factored"
        elif model == "action6":
            call = lambda: forms.vm_action(u_h, out=b_h)  # noqa: E731
            api = "QuadratureForms.vm_action(x_host, out=y_host): eo_form_action_vm_factored"
        else:
            call = lambda: forms.action("mandel_strain", "mandel_strain", forms.C_tang, u_h, out=b_h)  # noqa: E731
            api = "QuadratureForms.action(..., C_tang_resident, x_host, out=y_host): eo_form_action"
        Ke = max(2, min(K, 5))
        for _ in range(3):
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
        del call, tab, forms, vm

    # the headline leg's buffers are no longer needed: free them before the other legs allocate theirs
    wl_cfg_stats = _stats_cfg(model, stats)
    del wl
    gc.collect()

    e2e_dc = None
    if model == "vm" and args.e2e_n > 0 and not args.no_device_consumers:
        e2e_dc = e2e_device_consumers(ctx, eo, inputs, int(args.e2e_n), rank, world, max_over_ranks, barrier, args.cpu_seconds)
        gc.collect()

    models = None
    if model == "vm" and args.models:
        models = models_block(ctx, eo, inputs, args, rank, world, peaks, barrier, max_over_ranks, collective)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline for the dominant kernel (the only kernel in a step)
    k_ms = float(np.mean(kernel_ms))
    roofline = roofline_for(model, n, k_ms, peaks, _traffic(model, n))

    # ---- CPU baseline on this box's host cores (bounded sample)
    cpu = None
    if world == 1 and args.cpu_seconds > 0:  # reported at N = 1 only (the reference arm is timed separately at every N)
        cpu = cpu_leg(model, args.cpu_seconds, args.cpu_sample if model != "isihara" else None)
        if model in ("vm", "jitvm", "jitvm3d", "heat", "mc"):
            cm = "vm" if model in ("jitvm", "jitvm3d") else model
            n1 = int(args.cpu_sample) // 4 if model != "mc" else 20_000
            rate1, _, _ = cpu_port_rate(cm, n1, min(3.0, args.cpu_seconds), parallel=False)
            cpu["single_thread_value"] = rate1
            cpu["sample"] += f"; single-thread rate {rate1:.3e} QP/s"
        if model == "vm":
            nb = cpu_numba_rate(1_000_000, min(3.0, args.cpu_seconds))
            if nb is not None:
                cpu["numba_serial"] = nb

    cfg = {
        "workload": WORKLOADS[model], "qp_per_gpu": n, "state_layout": args.state_layout,
        "l2": f"inputs+outputs {BYTES_PER_QP[model] * n / 1e9:.1f} GB per step >> 126 MB L2 (no flush needed)",
        "partition": "contiguous block of QPs per rank, no halo; one statistics collective per step when n_gpus > 1 "
                     "(plastic_fraction / niter histogram are then the global figures)",
    }
    cfg.update(wl_cfg_stats)
    cfg.update(extra_cfg)
    if world > 1:
        cfg["cpu_binding"] = (f"rank pinned to the {len(numa_cores)} cores local to its GPU (NVML affinity)" if numa_cores
                              else "none")
    line = {
        "metric": METRIC, "value": value, "unit": "QP/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "kernel_ms_max_over_ranks": headline_kernel_ms_max, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks, "peaks": peaks,
    }
    if e2e_dc is not None:
        line["e2e_device_consumers"] = e2e_dc
    if collective_check is not None:
        line["collective"] = collective_check
    if models is not None:
        line["models"] = models
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="vm", choices=["vm", "heat", "mc", "tab", "fused", "isihara", "jitvm", "jitfused", "jitvm3d", "step", "action", "step6", "action6"])
    ap.add_argument("--fused-exact", action="store_true")
    ap.add_argument("--mc-scheme", default="classes", choices=["classes", "simple", "queue-noaffinity", "queue-onepass", "ring"])
    ap.add_argument("--n", "--qp-per-gpu", dest="n", type=float, default=1e8,
                    help="quadrature points per GPU (device-resident leg); under torchrun spell it --qp-per-gpu "
                         "(torchrun's own parser rejects --n as an ambiguous abbreviation)")
    ap.add_argument("--e2e-n", type=float, default=1.5e7, help="quadrature points per GPU for the end-to-end leg")
    ap.add_argument("--state-layout", default="aos", choices=["aos", "soa"])
    ap.add_argument("--mesh-order", default="structured", choices=["structured", "shuffled", "rcm", "morton"],
                    help="numbering of cells / dofs / nodes of the synthetic mesh (tab, fused, step, action): the row-major "
                         "structured numbering, a random permutation, reverse Cuthill-McKee of a shuffled mesh, or a Z-order "
                         "space-filling curve (cheap enough for 10^7 cells)")
    ap.add_argument("--cpu-sample", type=float, default=4e6)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--models", default="mc,heat,tab,fused,step,action,step6,action6,isihara",
                    help="vm (default) line: comma-separated legs of the `models` block ('' = none)")
    ap.add_argument("--models-steps", type=int, default=5)
    ap.add_argument("--models-cpu-seconds", type=float, default=2.0)
    ap.add_argument("--models-isihara-n", type=float, default=2e7)
    ap.add_argument("--no-device-consumers", action="store_true",
                    help="vm: skip the supplementary end-to-end leg through the device-side consumers (QuadratureForms)")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-GPU: do not pin ranks to their GPU's NUMA node")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    args.n = int(args.n)
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
