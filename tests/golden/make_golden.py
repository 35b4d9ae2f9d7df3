#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  The reference ships no tests or golden vectors for this path, so
these files pin its behaviour: known answers of the stateless solver functions, mesh/constraint/colouring structure of
small scenes, and short trajectories.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Outputs (committed): kat_f32.npz, kat_f64.npz, structure.npz, trajectories.npz
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle  # noqa: E402
from oracle.pyoracle import (DISTANCE, DISTANCE_XPBD, DIHEDRAL, ISOBENDING, ISOBENDING_XPBD, FEMTRIANGLE, STRAINTRIANGLE,  # noqa: E402
                             VOLUME, VOLUME_XPBD, FEMTET, FEMTET_XPBD, STRAINTET, SHAPEMATCHING)
import scenes  # noqa: E402
from parity_util import perturb  # noqa: E402

SOLVE_TYPES = [DISTANCE, DISTANCE_XPBD, DIHEDRAL, ISOBENDING, ISOBENDING_XPBD, FEMTRIANGLE, STRAINTRIANGLE, VOLUME, VOLUME_XPBD,
               FEMTET, FEMTET_XPBD, STRAINTET, SHAPEMATCHING]
INIT_TYPES = [ISOBENDING, FEMTRIANGLE, STRAINTRIANGLE, FEMTET, STRAINTET]


def rest_tuple(rng, ctype):
    """A well-conditioned O(1)-sized rest stencil for the type (4 points; unused ones ignored)."""
    if ctype in (DISTANCE, DISTANCE_XPBD):
        x = rng.uniform(-1, 1, (4, 3))
    elif ctype in (DIHEDRAL, ISOBENDING, ISOBENDING_XPBD):
        # two triangles sharing the edge (x2,x3); x0, x1 are the opposite vertices
        x = np.array([[0.0, 1.0, 0.0], [0.0, -1.0, 0.3], [-1.0, 0.0, 0.0], [1.0, 0.0, 0.0]]) + rng.uniform(-0.15, 0.15, (4, 3))
    elif ctype in (FEMTRIANGLE, STRAINTRIANGLE):
        x = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.2], [0.2, 0.0, 1.0], [0, 0, 0]]) + rng.uniform(-0.1, 0.1, (4, 3))
        x[3] = 0
    else:
        x = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]]) + rng.uniform(-0.15, 0.15, (4, 3))
    return x + rng.uniform(-2, 2, (1, 3))  # translate away from the origin a little


def material(rng, ctype):
    if ctype in (DISTANCE, DIHEDRAL, ISOBENDING, VOLUME):
        return [rng.choice([1.0, 0.5, 0.01])]
    if ctype in (DISTANCE_XPBD, ISOBENDING_XPBD, VOLUME_XPBD):
        return [rng.choice([1.0e5, 100.0, 0.0])]
    if ctype == FEMTRIANGLE:
        return [rng.choice([1.0, 1000.0]), rng.choice([1.0, 800.0]), rng.choice([1.0, 400.0]), 0.3, 0.25]
    if ctype == STRAINTRIANGLE:
        return [1.0, 0.8, 0.6, float(rng.randint(2)), float(rng.randint(2))]
    if ctype in (FEMTET, FEMTET_XPBD):
        return [rng.choice([1.0e6, 1.0e3, 1.0]), rng.choice([0.3, 0.45, 0.0])]
    if ctype == STRAINTET:
        return [1.0, 0.7, float(rng.randint(2)), float(rng.randint(2))]
    if ctype == SHAPEMATCHING:
        return [rng.choice([1.0, 0.5]), 1.0, 1.0, 1.0, 1.0]  # stiffness, numClusters = 1 (raw solver answer == applied correction)
    raise ValueError(ctype)


def make_kat(ref, n_per_type=24, seed=1234):
    """Inputs + the reference's answers.  Each case is generated through a 1-constraint model so that the rest data is the
    reference's own initConstraint result."""
    rng = np.random.RandomState(seed)
    cases = []
    for ctype in SOLVE_TYPES:
        for c in range(n_per_type):
            x0 = rest_tuple(rng, ctype)
            mat = material(rng, ctype)
            ref.reset()
            ref.add_triangle_model(x0, np.zeros((0, 3), np.uint32))  # 4 free particles
            ok = ref.add_constraint(ctype, list(range(pyoracle.NBODIES[ctype])), mat)
            if not ok:
                continue
            _, _, params, _ = ref.constraints()
            p = params[0]
            w = rng.choice([1.0, 0.5, 2.0, 0.0], size=4, p=[0.5, 0.2, 0.15, 0.15])
            if ctype == SHAPEMATCHING:
                w[:] = 1.0  # the constraint uses the inverse masses frozen at creation (all 1 here), not these
            amp = rng.choice([0.02, 0.2, 0.0], p=[0.5, 0.4, 0.1])
            x = x0 + rng.uniform(-amp, amp, (4, 3))
            if ctype in (FEMTET, FEMTET_XPBD) and c % 4 == 3:   # inverted element -> SVD branch
                x[3] = x[0] + (x[0] - x[3]) * 0.7
            if ctype in (DISTANCE, DISTANCE_XPBD) and c % 8 == 7:  # coincident points (d <= 1e-6 early-out / zero normal)
                x[1] = x[0]
            if c % 12 == 11 and ctype != SHAPEMATCHING:
                w[:] = 0.0  # all static
            lam0 = rng.choice([0.0, rng.uniform(-1e-3, 1e-3)])
            hinv = (c % 4 == 3) or (c % 5 == 0)
            dt = rng.choice([0.005, 0.001])
            res, corr, lam1 = ref.kat_solve(ctype, x, w, p, dt=dt, handle_inversion=hinv, lam=lam0)
            cases.append((ctype, x0, x, w, p, dt, float(hinv), lam0, res, corr, lam1))
    inits = []
    for ctype in INIT_TYPES:
        for c in range(8):
            x0 = rest_tuple(rng, ctype)
            if c == 7:  # degenerate rest state -> init returns false
                x0[2] = x0[1]; x0[3] = x0[1]
            res, out = ref.kat_init(ctype, x0)
            inits.append((ctype, x0, res, out))
    svds = []
    for c in range(16):
        A = rng.uniform(-1, 1, (3, 3)) + np.eye(3) * rng.choice([1.0, -1.0, 0.0])
        if c % 5 == 4:
            A[:, 2] = A[:, 0] * 0.5  # rank deficient
        s, U, VT = ref.kat_svd(A)
        svds.append((A, s, U, VT))
    integ = []
    for c in range(8):
        h = rng.choice([0.005, 0.001]); mass = rng.choice([1.0, 0.0, 2.0])
        x = rng.uniform(-2, 2, 3); v = rng.uniform(-1, 1, 3); a = np.array([0, -9.81, 0.0]); old = x - rng.uniform(-0.01, 0.01, 3); last = old - rng.uniform(-0.01, 0.01, 3)
        x1, v1 = ref.kat_integrate(h, mass, x, v, a)
        vf = ref.kat_velocity_update(0, h, mass, x, old, last, v)
        vs = ref.kat_velocity_update(1, h, mass, x, old, last, v)
        integ.append((h, mass, x, v, a, old, last, x1, v1, vf, vs))
    return dict(
        solve_type=np.array([c[0] for c in cases], np.int32), solve_x0=np.array([c[1] for c in cases]), solve_x=np.array([c[2] for c in cases]),
        solve_w=np.array([c[3] for c in cases]), solve_p=np.array([c[4] for c in cases]), solve_dt=np.array([c[5] for c in cases]),
        solve_hinv=np.array([c[6] for c in cases]), solve_lam0=np.array([c[7] for c in cases]), solve_res=np.array([c[8] for c in cases], np.int32),
        solve_corr=np.array([c[9] for c in cases]), solve_lam1=np.array([c[10] for c in cases]),
        init_type=np.array([c[0] for c in inits], np.int32), init_x0=np.array([c[1] for c in inits]), init_res=np.array([c[2] for c in inits], np.int32),
        init_out=np.array([c[3] for c in inits]),
        svd_A=np.array([c[0] for c in svds]), svd_s=np.array([c[1] for c in svds]), svd_U=np.array([c[2] for c in svds]), svd_VT=np.array([c[3] for c in svds]),
        integ=np.array([np.concatenate([[c[0], c[1]], *c[2:]]) for c in integ]))


STRUCT_SCENES = {
    "cfg1_50x50": lambda m: scenes.cfg1(m, 50),
    "cloth_xpbd_33x17": lambda m: scenes.cloth(m, 33, 17, 4, 3, dist_k=1e5, bend_k=100.0),
    "cloth_fem_dihedral_12x9": lambda m: scenes.cloth(m, 12, 9, 2, 1),
    "cloth_strain_10x10": lambda m: scenes.cloth(m, 10, 10, 3, 0),
    "bar_dist_vol_8x4x3": lambda m: scenes.bar(m, 8, 4, 3, 1),
    "bar_fem_vol_9x4x4": lambda m: scenes.bar(m, 9, 4, 4, 2, extra_volume=True),
    "bar_strain_6x3x3": lambda m: scenes.bar(m, 6, 3, 3, 4),
    "bar_xpbd_6x3x3": lambda m: scenes.bar(m, 6, 3, 3, 6, k=1e5, vol_k=1e5),
    "bar_shapematching_6x3x3": lambda m: scenes.bar(m, 6, 3, 3, 5, k=0.5),
    "cfg4_small_coupling": lambda m: scenes.cfg4(m, 14, (5, 3, 3)),
    "bar_femx_6x3x3": lambda m: scenes.bar(m, 6, 3, 3, 3),
}

TRAJ_SCENES = {
    # name: (builder, perturbation, steps)
    "cloth_distance_16": (lambda m: scenes.cloth(m, 16, 16, 1, 0, max_iter=4), 0.02, 3),
    "cloth_xpbd_16": (lambda m: scenes.cloth(m, 16, 16, 4, 3, dist_k=1e5, bend_k=100.0, max_iter=4), 0.02, 3),
    "cloth_isobend_16": (lambda m: scenes.cloth(m, 16, 16, 1, 2, bend_k=0.5, max_iter=4), 0.02, 3),
    "cloth_dihedral_16": (lambda m: scenes.cloth(m, 16, 16, 1, 1, bend_k=0.5, max_iter=4), 0.02, 3),
    "cloth_fem_16": (lambda m: scenes.cloth(m, 16, 16, 2, 0, fem=(1000.0, 1000.0, 500.0, 0.3, 0.3), max_iter=4), 0.02, 3),
    "cloth_strain_16": (lambda m: scenes.cloth(m, 16, 16, 3, 0, max_iter=4), 0.02, 3),
    "bar_dist_vol": (lambda m: scenes.bar(m, 7, 4, 4, 1, k=1.0, sub_steps=2, max_iter=3), 0.01, 3),
    "bar_fem": (lambda m: scenes.bar(m, 7, 4, 4, 2, k=1e6, sub_steps=2, max_iter=3), 0.01, 3),
    "bar_femx": (lambda m: scenes.bar(m, 7, 4, 4, 3, k=1e4, sub_steps=2, max_iter=3), 0.01, 1),  # XPBD-FEM is chaotic beyond one step
    "bar_strain": (lambda m: scenes.bar(m, 7, 4, 4, 4, k=1.0, sub_steps=2, max_iter=3), 0.01, 3),
    "bar_shapematching": (lambda m: scenes.bar(m, 7, 4, 4, 5, k=0.5, sub_steps=2, max_iter=3), 0.02, 3),
    "bar_xpbd": (lambda m: scenes.bar(m, 7, 4, 4, 6, k=1e5, vol_k=1e5, sub_steps=2, max_iter=3), 0.01, 3),
    "bar_fem_vol": (lambda m: scenes.bar(m, 7, 4, 4, 2, k=1e6, extra_volume=True, sub_steps=3, max_iter=2), 0.01, 3),
    "cloth_second_order_12": (lambda m: scenes.cloth(m, 12, 12, 1, 0, max_iter=3, sub_steps=2, vel_method=1), 0.02, 4),
    # (FEMTet at E=1e6 amplifies rounding differences ~30x per step even in fp64: keep the horizon short)
    "cfg4_small_coupling": (lambda m: scenes.cfg4(m, 14, (5, 3, 3), cloth_method=1), 0.0, 3),
}


def main():
    if not (pyoracle.available("ref", "f32") and pyoracle.available("ref", "f64")):
        pyoracle.build(ref=True)
    for prec in ("f32", "f64"):
        ref = pyoracle.CpuPbd("ref", prec)
        np.savez_compressed(os.path.join(HERE, "kat_%s.npz" % prec), **make_kat(ref))
    ref = pyoracle.CpuPbd("ref", "f64")
    st = {}
    for name, build in STRUCT_SCENES.items():
        ref.reset(); build(ref)
        t, b, p, nb = ref.constraints()
        off, ids = ref.groups()
        st[name + "/types"] = t; st[name + "/bodies"] = b; st[name + "/params"] = p; st[name + "/group_off"] = off; st[name + "/group_ids"] = ids
        st[name + "/x0"] = ref.get("x0")
        if "coupling" in name:
            st[name + "/rigid_bodies"] = ref.rigid_bodies()
        if name.startswith("cloth") or name.startswith("cfg1"):
            st[name + "/tri_edges"] = ref.tri_edges(0); st[name + "/tri_faces"] = ref.tri_faces(0)
        if name.startswith("bar"):
            st[name + "/tet_edges"] = ref.tet_edges(0); st[name + "/tet_tets"] = ref.tet_tets(0)
    np.savez_compressed(os.path.join(HERE, "structure.npz"), **st)
    tr = {}
    for prec in ("f32", "f64"):
        ref = pyoracle.CpuPbd("ref", prec)
        for name, (build, amp, steps) in TRAJ_SCENES.items():
            ref.reset(); build(ref)
            xs = perturb([ref], amp)
            tr["%s/%s/start" % (name, prec)] = xs
            ref.step(steps)
            tr["%s/%s/x" % (name, prec)] = ref.get("x"); tr["%s/%s/v" % (name, prec)] = ref.get("v")
            if "coupling" in name:
                tr["%s/%s/rb" % (name, prec)] = ref.rigid_bodies()
            assert np.isfinite(ref.get("x")).all(), name
            moved = np.abs(ref.get("x") - xs).max()
            assert moved > 1e-5, (name, "nothing moved")
    np.savez_compressed(os.path.join(HERE, "trajectories.npz"), **tr)
    for f in ("kat_f32.npz", "kat_f64.npz", "structure.npz", "trajectories.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
