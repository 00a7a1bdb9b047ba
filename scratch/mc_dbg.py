import sys, ctypes as C, time
sys.path.insert(0, '.')
import numpy as np
import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200 import synthetic as inputs
from dolfinx_external_operator_b200._lib import McParams
ctx = eo.Context(0)
mc = eo.MohrCoulomb(ctx=ctx, history=None)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
tile = min(n, 1 << 20)
d, s = inputs.mc_batch(tile, seed=0, stepper=mc.stress_update)
dd, ds = ctx.empty((n*4,)), ctx.empty((n*4,))
for r in range(0, n, tile):
    m = min(tile, n-r)
    ctx.copy(dd.ptr + r*32, np.ascontiguousarray(d[:m]), m*32); ctx.copy(ds.ptr + r*32, np.ascontiguousarray(s[:m]), m*32)
dC, dsig = ctx.empty((n*16,)), ctx.empty((n*4,))
prm = McParams(mc.E, mc.nu, mc.c, mc.phi, mc.psi, mc.theta_T, mc.a, mc.tol, mc.Nitermax)
schemes = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0]
for scheme in schemes * 3:
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    ctx.check(ctx.lib.eo_mc_eval_scheme(ctx.handle, C.byref(prm), dd.ptr, ds.ptr, dC.ptr, dsig.ptr, None, None, None, None, n, scheme))
    ctx.record(e1); ctx.sync()
    ms = ctx.elapsed_ms(e0, e1)
    cnt = np.zeros(64, dtype=np.uint32)
    ctx.check(ctx.lib.eo_debug_counters(ctx.handle, cnt.ctypes.data))
    c64 = cnt.view(np.uint64)
    c64 = c64[0:]
    tot = float(c64[8] + c64[9] + c64[10] + c64[11]) or 1.0
    print("  cycles share: T %.1f%% S0 %.1f%% U0 %.1f%% U %.1f%%  (sum %.3g warp-cycles; per SM-warp %.3g)" % (100*c64[8]/tot, 100*c64[9]/tot, 100*c64[10]/tot, 100*c64[11]/tot, tot, tot/148/12))
    print(f"scheme={scheme} n={n} {ms:.3f} ms  {n/ms/1e6:.3f} GQP/s  tiles={cnt[0]} T={cnt[1]}/{cnt[2]} S0={cnt[3]}/{cnt[4]} U0={cnt[5]}/{cnt[6]} U={cnt[7]}/{cnt[8]} wait={cnt[9]}")
