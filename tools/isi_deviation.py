"""tools/isi_deviation.py - worst deviation of the Isihara kernel from the reference golden, relative to the field
scale (the quantity tests/isi_util.check bounds by RTOL = 2e-6); run under gpurun."""
import sys
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import dolfinx_external_operator_b200 as eo
from isi_util import load_golden

g, sd = load_golden()
m = eo.Isihara(sd, H_flat=g["H_flat"], ctx=eo.Context(0))
dP, P = m((1,))(g["F"].reshape(-1, 1, 2, 2))
dP, P = dP.reshape(g["dP"].shape), P.reshape(g["P"].shape)
print("n %d  P %.3e  dP %.3e  (of the field maxima %.3e / %.3e)" % (
    g["F"].shape[0], np.abs(P - g["P"]).max() / np.abs(g["P"]).max(), np.abs(dP - g["dP"]).max() / np.abs(g["dP"]).max(),
    np.abs(g["P"]).max(), np.abs(g["dP"]).max()))
