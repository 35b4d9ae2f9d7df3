"""Development aid (CPU only): L2 sectors touched per warp-level particle gather for candidate particle layouts.

For every (colour, type) bucket the constraints are sorted by their lowest device slot (as the engine does), cut into warps
of 32 and, for every body position, the distinct 32-byte sectors (two float4 slots each) are counted.  1.0 = one sector per
thread (worst), 0.5 = perfectly dense."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from positionbaseddynamics_b200 import _capi
from positionbaseddynamics_b200.model import HostModel

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 200
hm = HostModel()
hm.add_regular_triangle_model(nx, nx, (0, 0, 0), np.eye(3), (10.0, 10.0))
hm.add_cloth_constraints(0, 4, 1e5)
hm.add_bending_constraints(0, 3, 100.0)
hm.init_groups()
types, bodies, params, nb = hm.constraints()
off, ids = hm.groups()
n = hm.num_particles()
print("particles", n, "constraints", len(types), "colours", len(off) - 1)
i = np.arange(n); row = i // nx; col = i % nx


def rank_by(key):
    order = np.argsort(key, kind="stable")
    slot = np.empty(n, dtype=np.int64); slot[order] = np.arange(n)
    return slot


layouts = {
    "linear": i,
    "deinterleave (i&1)": rank_by(i & 1),
    "col parity, row parity (4 classes)": rank_by((col & 1) * 2 + (row & 1)),
    "row parity major, col parity (4 classes)": rank_by((row & 1) * 2 + (col & 1)),
    "col mod 3": rank_by(col % 3),
    "col mod 4": rank_by(col % 4),
    "(col&1, row mod 3)": rank_by((col & 1) * 3 + row % 3),
    "(col mod 3, row mod 3)": rank_by((col % 3) * 3 + row % 3),
    "(col&3, row&1)": rank_by((col & 3) * 2 + (row & 1)),
    "(col&1,row&1) inside 64x64 blocks": rank_by(((row // 64) * ((nx + 63) // 64) + col // 64) * 4 + (col & 1) * 2 + (row & 1)),
}


def simulate(slot):
    tot_sectors = 0; tot_gathers = 0
    per_type = {}
    for g in range(len(off) - 1):
        cid = ids[off[g]:off[g + 1]]
        for t in np.unique(types[cid]):
            sel = cid[types[cid] == t]
            k = _capi.num_bodies(int(t))
            s = slot[bodies[sel][:, :k].astype(np.int64)]
            order = np.argsort(s.min(axis=1), kind="stable")
            s = s[order]
            pad = (-len(s)) % 32
            if pad:
                s = np.concatenate([s, np.repeat(s[-1:], pad, axis=0)])
            sec = (s // 2).reshape(-1, 32, k)
            cnt = 0
            for b in range(k):
                x = np.sort(sec[:, :, b], axis=1)
                cnt += (np.diff(x, axis=1) != 0).sum() + len(x)
            tot_sectors += cnt; tot_gathers += len(sel) * k
            a = per_type.setdefault(int(t), [0, 0]); a[0] += cnt; a[1] += len(sel) * k
    return tot_sectors / tot_gathers, {t: v[0] / v[1] for t, v in per_type.items()}


for name, key in layouts.items():
    tot, per = simulate(key)
    print("%-45s sectors/gather %.3f   %s" % (name, tot, {_capi.TYPE_NAMES[t]: round(v, 3) for t, v in per.items()}))
