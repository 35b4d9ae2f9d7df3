"""Development aid: cfg2 + three static colliders, K steps -- the workload of bench.py's side.cfg2_contacts as a stand-alone script, so that
`ncu -k regex:k_contacts` can capture the contact kernel.  Usage: python tools/contacts_bench.py [size] [steps]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
from positionbaseddynamics_b200 import _capi

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
bench.WORKLOAD = "cfg2"
eng, inf = bench.make_engine(0, size, 20)
eng.set_mode(_capi.MODE_AUTO)
xc = eng.get_attr(_capi.ATTR_X); cx, cy, cz = [float(v) for v in xc.mean(axis=0)]
ident = [1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0]
def collider(shape, body, dim, centre, half):
    rc = _capi.RigidCollider(); rc.shape = shape; rc.body = body; rc.dim[:] = list(dim) + [0.0] * (3 - len(dim)); rc.restitution = 0.6; rc.friction = 0.2
    rc.R[:] = ident; rc.v2[:] = list(centre); rc.aabb_min[:] = [c - h - 0.05 for c, h in zip(centre, half)]; rc.aabb_max[:] = [c + h + 0.05 for c, h in zip(centre, half)]
    return rc
centres = [(cx, cy - 3.0, cz), (cx, cy - 1.7, cz), (cx + 3.0, cy - 0.3, cz + 2.0)]
eng.set_rigid_bodies([0.0] * 3, centres, [(1.0, 0, 0, 0)] * 3, [(1.0, 1.0, 1.0)] * 3)
eng.set_colliders([_capi.ParticleCollider(0, inf["n"], 0.5, 0.1)],
                  [collider(_capi.SHAPE_BOX, 0, (50.0, 0.5, 50.0), centres[0], (50.0, 0.5, 50.0)), collider(_capi.SHAPE_SPHERE, 1, (2.0,), centres[1], (2.0, 2.0, 2.0)),
                   collider(_capi.SHAPE_TORUS, 2, (1.5, 0.5), centres[2], (2.0, 0.5, 2.0))])
eng.set_contact_params(0.05, 100.0, 5); eng.record_contacts(1 << 20)
eng.step(steps); eng.sync()
_, found = eng.contacts(1)
print("cfg2 %dx%d + box/sphere/torus: %d steps, %d contacts in the last step, %.3f ms per step (last)" % (size, size, steps, found, eng.stats().last_step_ms / steps))
