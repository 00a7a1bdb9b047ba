"""Small-problem latency of the public callables (the reference demos' sizes): time per call, host arrays in and out."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200 import synthetic as syn
from oracle import native, constitutive as oc

ctx = eo.Context(0)
native.build(); native.use_all_cores()
for n in (600, 3750, 10_000, 100_000, 1_000_000):
    deps, sn, p = syn.vm_batch(n, seed=0)
    vm = eo.VonMises(ctx=ctx); vm.set_history(sn, p)
    f = vm((1,)); d3 = deps.reshape(-1, 1, 4)
    for _ in range(5): f(d3)
    t0 = time.perf_counter(); K = 200 if n <= 100_000 else 20
    for _ in range(K): f(d3)
    t_gpu = (time.perf_counter() - t0) / K
    prm = oc.VonMisesParams()
    native.vm_return_mapping(deps, sn, p, prm, parallel=False)
    t0 = time.perf_counter(); Kc = 20
    for _ in range(Kc): native.vm_return_mapping(deps, sn, p, prm, parallel=False)
    t_cpu1 = (time.perf_counter() - t0) / Kc
    mc = eo.MohrCoulomb(ctx=ctx, history=None)
    dm, sm = syn.mc_batch(n, seed=0, stepper=mc.stress_update)
    m2 = eo.MohrCoulomb(ctx=ctx); m2.set_history(sm)
    g = m2((1,)); dm3 = dm.reshape(-1, 1, 4)
    for _ in range(5): g(dm3)
    t0 = time.perf_counter()
    for _ in range(K): g(dm3)
    t_mc = (time.perf_counter() - t0) / K
    print(f"n={n:8d}  vm callable {1e6*t_gpu:9.1f} us   cpu C port 1 thread {1e6*t_cpu1:9.1f} us   mc callable {1e6*t_mc:9.1f} us")
