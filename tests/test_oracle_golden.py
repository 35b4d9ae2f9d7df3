"""The C restatement (oracle/) against the committed golden fixtures generated from the unmodified reference
(tests/golden/make_golden.py).  These pin the oracle on every box, including ones without /root/reference.
Tolerances: integers exact; fp64 build 1e-9 relative (same algorithm, different operation order); fp32 build 2e-5 of the
scale of the quantity (fp32 rounding + FMA contraction differences) -- stated per assertion.
"""
import os
import numpy as np
import pytest

import scenes
from golden.make_golden import STRUCT_SCENES, TRAJ_SCENES
from oracle.pyoracle import NPARAMS, ISOBENDING, ISOBENDING_XPBD, SHAPEMATCHING

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOLS = {"f32": 3e-5, "f64": 1e-9}


def _scale(a):
    return max(float(np.abs(a).max()), 1e-12)


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_known_answers_solvers(prec, cpu_libs):
    d = np.load(os.path.join(G, "kat_%s.npz" % prec))
    o = cpu_libs.CpuPbd("oracle", prec)
    tol = TOLS[prec]
    worst = {}
    for i in range(len(d["solve_type"])):
        t = int(d["solve_type"][i])
        res, corr, lam = o.kat_solve(t, d["solve_x"][i], d["solve_w"][i], d["solve_p"][i][:NPARAMS[t]], dt=float(d["solve_dt"][i]),
                                     handle_inversion=bool(d["solve_hinv"][i]), lam=float(d["solve_lam0"][i]))
        assert res == int(d["solve_res"][i]), (t, i)
        ref = d["solve_corr"][i]
        # scale: the correction itself, floored at 1e-3 (stencils are O(1)) so that rest-state cases, whose answer is pure rounding noise, compare cleanly
        sc = max(_scale(ref), 1e-3)
        # isometric bending evaluates a cancelling quadratic form on absolute positions; its fp32 noise is set by
        # |Q||x|^2 ulp, not by the size of the correction (DESIGN.md "Parity")
        loose = 50.0 if (prec == "f32" and t in (ISOBENDING, ISOBENDING_XPBD)) else 1.0
        err = float(np.abs(corr - ref).max() / sc)
        worst[t] = max(worst.get(t, 0.0), err)
        assert err <= tol * 40 * loose, (t, i, err)
        assert abs(lam - float(d["solve_lam1"][i])) <= tol * 40 * loose * max(abs(float(d["solve_lam1"][i])), 1e-3), (t, i)
    print(prec, {k: "%.1e" % v for k, v in sorted(worst.items())})


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_known_answers_init_svd_integration(prec, cpu_libs):
    d = np.load(os.path.join(G, "kat_%s.npz" % prec))
    o = cpu_libs.CpuPbd("oracle", prec)
    tol = TOLS[prec]
    for i in range(len(d["init_type"])):
        res, out = o.kat_init(int(d["init_type"][i]), d["init_x0"][i])
        assert res == int(d["init_res"][i])
        if res:
            ref = d["init_out"][i]
            # a degenerate bending stencil makes the reference return true with non-finite Q (cot of a zero angle):
            # the restatement must produce the same non-finite pattern
            assert (np.isfinite(out) == np.isfinite(ref)).all()
            fin = np.isfinite(ref)
            if fin.any():
                assert np.abs(out[fin] - ref[fin]).max() <= tol * 20 * _scale(ref[fin])
    for i in range(len(d["svd_A"])):
        s, U, VT = o.kat_svd(d["svd_A"][i])
        # singular values and the reconstruction are robust quantities; U/VT individually are not for (near-)repeated values
        assert np.abs(s - d["svd_s"][i]).max() <= (2e-3 if prec == "f32" else 1e-7)
        A1 = U @ np.diag(s) @ VT; A0 = d["svd_U"][i] @ np.diag(d["svd_s"][i]) @ d["svd_VT"][i]
        assert np.abs(A1 - A0).max() <= (2e-3 if prec == "f32" else 1e-7)
    for row in d["integ"]:
        h, mass = row[0], row[1]
        x, v, a, old, last, x1, v1, vf, vs = [row[2 + 3 * k: 5 + 3 * k] for k in range(9)]
        xo, vo = o.kat_integrate(h, mass, x, v, a)
        assert np.abs(xo - x1).max() <= tol * 10 and np.abs(vo - v1).max() <= tol * 10
        assert np.abs(o.kat_velocity_update(0, h, mass, x, old, last, v) - vf).max() <= tol * 50 * max(_scale(vf), 1)
        assert np.abs(o.kat_velocity_update(1, h, mass, x, old, last, v) - vs).max() <= tol * 50 * max(_scale(vs), 1)


@pytest.mark.parametrize("name", sorted(STRUCT_SCENES))
def test_structure_matches_reference(name, cpu_libs):
    d = np.load(os.path.join(G, "structure.npz"))
    o = cpu_libs.CpuPbd("oracle", "f64")
    STRUCT_SCENES[name](o)
    t, b, p, nb = o.constraints()
    off, ids = o.groups()
    assert (t == d[name + "/types"]).all()
    nbod = np.array([2 if tt >= 13 or tt < 2 else 4 for tt in t])
    assert all((b[i][:nbod[i]] == d[name + "/bodies"][i][:nbod[i]]).all() for i in range(len(t)))
    assert (off == d[name + "/group_off"]).all() and (ids == d[name + "/group_ids"]).all()
    ref_p = d[name + "/params"]
    assert np.abs(p - ref_p).max() <= 1e-9 * _scale(ref_p)
    if name + "/rigid_bodies" in d:
        assert np.abs(o.rigid_bodies() - d[name + "/rigid_bodies"]).max() <= 1e-12
    assert np.abs(o.get("x0") - d[name + "/x0"]).max() <= 1e-12
    if name + "/tri_edges" in d:
        assert (o.tri_edges(0) == d[name + "/tri_edges"]).all() and (o.tri_faces(0) == d[name + "/tri_faces"]).all()
    if name + "/tet_edges" in d:
        assert (o.tet_edges(0) == d[name + "/tet_edges"]).all() and (o.tet_tets(0) == d[name + "/tet_tets"]).all()


@pytest.mark.parametrize("name", sorted(TRAJ_SCENES))
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_trajectories_match_reference(name, prec, cpu_libs):
    if prec == "f32" and "coupling" in name:
        pytest.skip("cfg4 puts IsometricBending (PBD) on a cloth at |x|~5: the fp32 reference is cancellation noise there "
                    "(1e-2 relative between two fp32 evaluation orders); the fp64 fixture pins this scene")
    d = np.load(os.path.join(G, "trajectories.npz"))
    build, amp, steps = TRAJ_SCENES[name]
    o = cpu_libs.CpuPbd("oracle", prec)
    build(o)
    o.set("x", d["%s/%s/start" % (name, prec)])
    o.step(steps)
    x_ref = d["%s/%s/x" % (name, prec)]
    err = np.abs(o.get("x") - x_ref).max() / _scale(x_ref)
    if "%s/%s/rb" % (name, prec) in d:
        rb_ref = d["%s/%s/rb" % (name, prec)]
        err = max(err, np.abs(o.rigid_bodies() - rb_ref).max() / _scale(rb_ref))
    tol = 1e-9 if prec == "f64" else 2e-5
    if prec == "f64" and "coupling" in name:
        tol = 1e-8
    if prec == "f32" and ("isobend" in name or "coupling" in name):  # cfg4 uses IsometricBending (PBD) on the cloth
        tol = 2e-3 if "isobend" in name else 1e-2  # fp32 cancellation noise of the reference's own bending evaluation (see DESIGN.md "Parity")
    if prec == "f32" and ("dihedral" in name or "femx" in name):
        tol = 2e-4  # acos near a flat hinge / sqrt(2U') near the rest state: ill-conditioned in fp32 on both sides
    assert err <= tol, (name, prec, err)
