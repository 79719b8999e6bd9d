import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "auv-real-time-interpolation_b200", "python")); sys.path.insert(0, ROOT)
import auvi
from oracle import binding as ob
n_lat, n_lon = int(sys.argv[1]), int(sys.argv[2])
z = ob.synth_grid(n_lat, n_lon)
bounds = (-180.0, -160.0, 20.0, 30.0)
g = auvi.Grid(z, *bounds)
pts, a, b = ob.lattice_queries(n_lat, n_lon, *bounds)
orc = ob.Oracle(z, *bounds)
for meth in (auvi.BILINEAR, auvi.CUBIC):
    got = g.lattice(meth, auvi.AXIS_EXPANDED, 2, 2)
    want = orc.batch(meth, pts).reshape(a, b)
    print(meth, "tma", g.uses_tma, "maxdiff", np.nanmax(np.abs(got - want)))
