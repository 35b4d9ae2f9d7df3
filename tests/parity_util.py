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
