"""Streaming efficiency of the generic (NVRTC) path for various components-per-point counts."""
import sys
sys.path.insert(0, '.')
import numpy as np
import dolfinx_external_operator_b200 as eo
from dolfinx_external_operator_b200.jit import JitModel
ctx = eo.Context(0)
only = [int(a) for a in sys.argv[1:]]
rng = np.random.default_rng(0)

def rand_dev(n):
    """device array of n random doubles (NOT zeros: cleared memory is read without touching DRAM)"""
    t = rng.standard_normal(1 << 22)
    a = ctx.empty((n,)); dt = ctx.to_device(t)
    for r in range(0, n, t.size):
        ctx.copy(a.ptr + 8 * r, dt, 8 * min(t.size, n - r))
    ctx.sync(); dt.free()
    return a

for S in (only or (1, 2, 3, 4, 5, 6, 9)):
    src = "template <class T> __device__ void f(const T* x, const double*, const double* p, T* y, T*) { for (int i = 0; i < %d; ++i) y[i] = p[0] * x[i] * x[(i + 1) %% %d] + 1.0; }" % (S, S)
    m = JitModel(src, "f", [(S,)], (S,), params=[2.0], ctx=ctx)
    n = int(2.4e9 / (16 * S))
    x = rand_dev(n * S); y = ctx.empty((n * S,))
    for _ in range(3): m.eval_device((0,), [x], y)
    e0, e1 = ctx.event(), ctx.event()
    ctx.record(e0)
    for _ in range(10): m.eval_device((0,), [x], y)
    ctx.record(e1); ctx.sync()
    ms = ctx.elapsed_ms(e0, e1) / 10
    print(f"S={S}: n={n} {ms:.3f} ms  {16*S*n/ms/1e6:.0f} GB/s  ({16*S*n/ms/1e6/6536.4:.2f} of peak)")
    # first derivative: out S*S
    if S <= 6:
        n2 = int(2.4e9 / (8 * S + 8 * S * S)); x2 = rand_dev(n2 * S); y2 = ctx.empty((n2 * S * S,))
        for _ in range(3): m.eval_device((1,), [x2], y2)
        ctx.record(e0)
        for _ in range(10): m.eval_device((1,), [x2], y2)
        ctx.record(e1); ctx.sync()
        ms = ctx.elapsed_ms(e0, e1) / 10
        print(f"      d/dx: n={n2} {ms:.3f} ms  {(8*S+8*S*S)*n2/ms/1e6:.0f} GB/s  ({(8*S+8*S*S)*n2/ms/1e6/6536.4:.2f})")
        x2.free(); y2.free()
    x.free(); y.free()
