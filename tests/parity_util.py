"""Shared helpers of the parity tests: seeded perturbations and error metrics."""
import numpy as np


def perturb(models, amplitude, seed=1234, pinned_too=False):
    """Displace the current positions of every model by the same seeded noise (uniform in [-amplitude, amplitude])."""
    x = np.asarray(models[0].get("x"), dtype=np.float64)
    rng = np.random.RandomState(seed)
    noise = rng.uniform(-amplitude, amplitude, size=x.shape)
    if not pinned_too:
        m, w = models[0].masses()
        noise[np.asarray(w) == 0] = 0.0
    xp = (x + noise).astype(np.float32)  # identical fp32 start state for every implementation
    for mdl in models:
        mdl.set("x", xp)
    return xp


def rel_position_error(x, x_ref):
    """max |dx| / max |x_ref|  -- the north-star's "relative on particle positions"."""
    x = np.asarray(x, dtype=np.float64); x_ref = np.asarray(x_ref, dtype=np.float64)
    return float(np.abs(x - x_ref).max() / max(np.abs(x_ref).max(), 1e-30))


def rel_displacement_error(x, x_ref, x_start):
    """max |dx| / max |x_ref - x_start|: error relative to how far the particles actually moved (stricter)."""
    x = np.asarray(x, dtype=np.float64); x_ref = np.asarray(x_ref, dtype=np.float64); x_start = np.asarray(x_start, dtype=np.float64)
    return float(np.abs(x - x_ref).max() / max(np.abs(x_ref - x_start).max(), 1e-30))


TOL_POS = 1e-4    # north_star: relative on particle positions
TOL_DISP = 5e-3   # relative to the distance the particles moved in the test (what makes the gate discriminating)


def parity_errors(x, x_ref, x_start):
    return rel_position_error(x, x_ref), rel_displacement_error(x, x_ref, x_start)


def assert_parity(x, x_ref, x_start, tol_pos=TOL_POS, tol_disp=TOL_DISP, what=""):
    """The parity gate: positions within tol_pos of the checker relative to the scene size AND within tol_disp relative
    to the largest displacement of the test.  The second bound is what a step that did nothing (or skipped the projections)
    cannot meet; `assert_gate_rejects` is the negative control."""
    e_pos, e_disp = parity_errors(x, x_ref, x_start)
    assert np.isfinite(np.asarray(x)).all(), what
    assert e_pos <= tol_pos, "%s: relative position error %.3e > %.1e" % (what, e_pos, tol_pos)
    assert e_disp <= tol_disp, "%s: error relative to the displacement %.3e > %.1e" % (what, e_disp, tol_disp)
    return e_pos, e_disp


def assert_gate_rejects(x_bad, x_ref, x_start, what=""):
    """Negative control: the gate must FAIL for x_bad (e.g. the untouched start state, or a step without projections)."""
    try:
        assert_parity(x_bad, x_ref, x_start, what=what)
    except AssertionError:
        return
    raise AssertionError("%s: the parity gate accepted a state that must fail it (gate not discriminating)" % what)
