"""Development aid: where the time to the first step goes (cfg2, C++ host model -> engine, both execution modes).
Run on the GPU box: python tools/flatten_time.py [size]"""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, scenes
from positionbaseddynamics_b200 import _capi
from positionbaseddynamics_b200.model import HostModel
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
t0 = time.perf_counter(); hm = HostModel()
hm.add_regular_triangle_model(n, n, t=(0, 1, 0), R=scenes.RX90, scale=(10, 10)); t1 = time.perf_counter()
hm.set_mass(0, 0.0); hm.set_mass(n - 1, 0.0)
hm.add_cloth_constraints(0, 4, dist_k=1e5); t2 = time.perf_counter()
hm.add_bending_constraints(0, 3, 100.0); t3 = time.perf_counter()
hm.init_groups(); t4 = time.perf_counter()
print("host model: mesh %.3f s | distance %.3f | bending %.3f | colouring %.3f | total %.3f" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0))
hm.set_params(dt=0.005, sub_steps=1, max_iter=20)
for mode, name in ((_capi.MODE_GRAPH, "graph"), (_capi.MODE_RESIDENT, "resident")):
    ts = hm.time_step(device=0); ts.set_mode(mode)
    t5 = time.perf_counter(); hm.step(1); ts.sync(); t6 = time.perf_counter(); hm.step(1); ts.sync(); t7 = time.perf_counter()
    print("%s: first step (upload + flatten + partition + capture) %.3f s, second step %.4f s" % (name, t6 - t5, t7 - t6))
