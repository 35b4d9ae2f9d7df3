"""ctypes binding of the engine-level C ABI (include/pbd_b200.h) exported by libpbd_b200.so.

This is plumbing only: every call goes straight into the CUDA library.  There is no Python or CPU fallback --
if the shared library is missing or no CUDA device is present the calls raise.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PBD_B200_LIB") or os.path.join(_HERE, "libpbd_b200.so")  # the override is a development aid (A/B builds)

(DISTANCE, DISTANCE_XPBD, DIHEDRAL, ISOBENDING, ISOBENDING_XPBD, FEMTRIANGLE, STRAINTRIANGLE, VOLUME, VOLUME_XPBD,
 FEMTET, FEMTET_XPBD, STRAINTET, SHAPEMATCHING, BALLJOINT, RB_PARTICLE_BALLJOINT) = range(15)
NUM_TYPES = 15
TYPE_NAMES = ["Distance", "Distance_XPBD", "Dihedral", "IsometricBending", "IsometricBending_XPBD", "FEMTriangle",
              "StrainTriangle", "Volume", "Volume_XPBD", "FEMTet", "FEMTet_XPBD", "StrainTet", "ShapeMatching", "BallJoint",
              "RigidBodyParticleBallJoint"]
ATTR_X, ATTR_V, ATTR_X0, ATTR_OLDX, ATTR_LASTX = range(5)
MODE_GRAPH, MODE_RESIDENT, MODE_LAUNCH, MODE_JACOBI, MODE_AUTO = 0, 1, 2, 3, 4

# every symbol include/pbd_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = ["pbd_last_error", "pbd_device_count", "pbd_create", "pbd_destroy", "pbd_set_particles", "pbd_set_attr",
           "pbd_get_attr", "pbd_set_masses", "pbd_set_rigid_bodies", "pbd_get_rigid_bodies", "pbd_clear_constraints", "pbd_add_constraints", "pbd_num_bodies",
           "pbd_num_params", "pbd_set_groups", "pbd_color_first_fit", "pbd_color_first_fit_device", "pbd_pin_host", "pbd_unpin_host", "pbd_get_num_groups", "pbd_get_groups",
           "pbd_set_params", "pbd_set_mode", "pbd_get_mode", "pbd_set_bucket_sort", "pbd_step", "pbd_sync", "pbd_step_host", "pbd_step_host_async", "pbd_step_host_wait",
           "pbd_get_lambdas", "pbd_get_stats", "pbd_profile_step",
           "pbd_set_colliders", "pbd_set_contact_params", "pbd_record_contacts", "pbd_get_contacts"]


SHAPE_BOX, SHAPE_SPHERE, SHAPE_TORUS, SHAPE_CYLINDER, SHAPE_HOLLOW_SPHERE, SHAPE_HOLLOW_BOX = range(6)


class ParticleCollider(C.Structure):
    _fields_ = [("offset", C.c_uint), ("count", C.c_uint), ("restitution", C.c_float), ("friction", C.c_float)]


class RigidCollider(C.Structure):
    _fields_ = [("shape", C.c_int), ("body", C.c_uint), ("dim", C.c_float * 3), ("thickness", C.c_float), ("invert_sdf", C.c_int),
                ("restitution", C.c_float), ("friction", C.c_float), ("R", C.c_float * 9), ("v1", C.c_float * 3), ("v2", C.c_float * 3),
                ("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3)]


class Contact(C.Structure):
    _fields_ = [("particle", C.c_uint), ("body", C.c_uint), ("cp0", C.c_float * 3), ("cp1", C.c_float * 3), ("normal", C.c_float * 3), ("dist", C.c_float)]


class Stats(C.Structure):
    _fields_ = [("projections", C.c_ulonglong), ("kernel_launches", C.c_ulonglong), ("steps", C.c_ulonglong),
                ("num_particles", C.c_uint), ("num_constraints", C.c_uint), ("num_groups", C.c_uint),
                ("num_buckets", C.c_uint), ("constraints_per_type", C.c_uint * NUM_TYPES),
                ("bytes_per_step", C.c_double), ("last_step_ms", C.c_float)]


class PbdError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libpbd_b200.so (built in-tree by __graft_entry__.build()).  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PbdError("native library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no Python/CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.pbd_last_error.restype = C.c_char_p
        _lib.pbd_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        for name in SYMBOLS:
            fn = getattr(_lib, name)
            if name not in ("pbd_last_error",):
                fn.restype = C.c_int
        for name in ("pbd_destroy", "pbd_clear_constraints", "pbd_color_first_fit", "pbd_sync"):
            getattr(_lib, name).argtypes = [C.c_void_p]
        _lib.pbd_color_first_fit_device.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_uint)]
        _lib.pbd_set_particles.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.pbd_set_attr.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.pbd_get_attr.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.pbd_set_masses.argtypes = [C.c_void_p, C.c_void_p]
        _lib.pbd_set_rigid_bodies.argtypes = [C.c_void_p, C.c_uint] + [C.c_void_p] * 6
        _lib.pbd_get_rigid_bodies.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        _lib.pbd_add_constraints.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.pbd_set_groups.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
        _lib.pbd_get_num_groups.argtypes = [C.c_void_p, C.POINTER(C.c_uint)]
        _lib.pbd_get_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.pbd_set_params.argtypes = [C.c_void_p, C.c_float, C.c_uint, C.c_uint, C.c_int, C.c_void_p]
        _lib.pbd_set_mode.argtypes = [C.c_void_p, C.c_int]
        _lib.pbd_get_mode.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib.pbd_set_bucket_sort.argtypes = [C.c_void_p, C.c_int]
        _lib.pbd_step.argtypes = [C.c_void_p, C.c_uint]
        _lib.pbd_step_host.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.pbd_step_host_async.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.pbd_step_host_wait.argtypes = [C.c_void_p, C.c_uint]
        _lib.pbd_set_colliders.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint, C.c_void_p]
        _lib.pbd_set_contact_params.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_uint]
        _lib.pbd_record_contacts.argtypes = [C.c_void_p, C.c_uint]
        _lib.pbd_get_contacts.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.POINTER(C.c_uint)]
        _lib.pbd_get_lambdas.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.pbd_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        _lib.pbd_profile_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _ck(rc):
    if rc != 0:
        raise PbdError(lib().pbd_last_error().decode())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count():
    n = C.c_int(0)
    rc = lib().pbd_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def num_bodies(t):
    return lib().pbd_num_bodies(t)


def num_params(t):
    return lib().pbd_num_params(t)


class Engine:
    """One engine <-> one CUDA device + one stream (pbd_create / pbd_destroy)."""

    def __init__(self, device=0, stream=None):
        self._h = C.c_void_p()
        _ck(lib().pbd_create(int(device), C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.n = 0
        self._nc = 0  # constraints added so far (sizes the buffer of groups())

    def close(self):
        if self._h:
            lib().pbd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # particles -------------------------------------------------------------------------------------
    def set_particles(self, x, mass, x0=None, v=None):
        x = _f32(x).reshape(-1, 3); mass = _f32(mass).reshape(-1)
        x0 = _f32(x0).reshape(-1, 3) if x0 is not None else None
        v = _f32(v).reshape(-1, 3) if v is not None else None
        self.n = len(x)
        _ck(lib().pbd_set_particles(self._h, self.n, _ptr(x), _ptr(x0), _ptr(v), _ptr(mass)))

    def set_attr(self, attr, a):
        a = _f32(a).reshape(-1, 3); assert len(a) == self.n
        _ck(lib().pbd_set_attr(self._h, attr, _ptr(a)))

    def get_attr(self, attr=ATTR_X, out=None):
        if out is None:
            out = np.empty((self.n, 3), dtype=np.float32)
        _ck(lib().pbd_get_attr(self._h, attr, _ptr(out)))
        return out

    def set_masses(self, mass):
        mass = _f32(mass).reshape(-1); assert len(mass) == self.n
        _ck(lib().pbd_set_masses(self._h, _ptr(mass)))

    # rigid bodies ----------------------------------------------------------------------------------
    def set_rigid_bodies(self, mass, x, q, inertia, v=None, omega=None):
        """q rows = (w, x, y, z); inertia rows = principal moments."""
        mass = _f32(mass).reshape(-1); self.n_rb = len(mass)
        x = _f32(x).reshape(-1, 3); q = _f32(q).reshape(-1, 4); inertia = _f32(inertia).reshape(-1, 3)
        v = _f32(v).reshape(-1, 3) if v is not None else None
        omega = _f32(omega).reshape(-1, 3) if omega is not None else None
        _ck(lib().pbd_set_rigid_bodies(self._h, self.n_rb, _ptr(mass), _ptr(x), _ptr(q), _ptr(inertia), _ptr(v), _ptr(omega)))

    def get_rigid_bodies(self):
        """[n, 13]: x(3) q(w,x,y,z) v(3) omega(3)."""
        n = getattr(self, "n_rb", 0)
        x = np.zeros((max(n, 1), 3), np.float32); q = np.zeros((max(n, 1), 4), np.float32)
        v = np.zeros((max(n, 1), 3), np.float32); w = np.zeros((max(n, 1), 3), np.float32)
        _ck(lib().pbd_get_rigid_bodies(self._h, _ptr(x), _ptr(q), _ptr(v), _ptr(w)))
        return np.concatenate([x, q, v, w], axis=1)[:n]

    # constraints -----------------------------------------------------------------------------------
    def clear_constraints(self):
        _ck(lib().pbd_clear_constraints(self._h))
        self._nc = 0

    def add_constraints(self, ctype, bodies, params, ids=None):
        nb, npar = num_bodies(ctype), num_params(ctype)
        bodies = np.ascontiguousarray(bodies, dtype=np.uint32).reshape(-1, nb)
        params = _f32(params).reshape(-1, npar)
        assert len(bodies) == len(params)
        ids = np.ascontiguousarray(ids, dtype=np.uint32) if ids is not None else None
        _ck(lib().pbd_add_constraints(self._h, ctype, len(bodies), _ptr(bodies), _ptr(params), _ptr(ids)))
        self._nc += len(bodies)

    def add_flat(self, types, bodies, params):
        """Insert a whole flat constraint list (types[n], bodies[n,4], params[n,24]) keeping insertion ids."""
        types = np.asarray(types)
        for t in range(NUM_TYPES):
            sel = np.nonzero(types == t)[0]
            if len(sel):
                self.add_constraints(t, np.asarray(bodies)[sel][:, :num_bodies(t)], np.asarray(params)[sel][:, :num_params(t)],
                                     ids=sel.astype(np.uint32))

    def set_groups(self, offsets, ids):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32); ids = np.ascontiguousarray(ids, dtype=np.uint32)
        _ck(lib().pbd_set_groups(self._h, len(offsets) - 1, _ptr(offsets), _ptr(ids)))

    def color_first_fit(self):
        _ck(lib().pbd_color_first_fit(self._h))

    def color_first_fit_device(self):
        """Exact first-fit colouring on the GPU; returns (device ms, wavefronts)."""
        ms = C.c_float(0.0); wf = C.c_uint(0)
        _ck(lib().pbd_color_first_fit_device(self._h, C.byref(ms), C.byref(wf)))
        return ms.value, wf.value

    def groups(self):
        ng = C.c_uint(0)
        _ck(lib().pbd_get_num_groups(self._h, C.byref(ng)))
        off = np.zeros(ng.value + 1, dtype=np.uint32)
        ids = np.zeros(max(self._nc, 1), dtype=np.uint32)  # the groups partition all constraints added so far
        _ck(lib().pbd_get_groups(self._h, _ptr(off), _ptr(ids)))
        return off, ids[:off[-1]]

    # parameters / stepping -------------------------------------------------------------------------
    def set_params(self, dt=0.005, sub_steps=5, max_iter=1, vel_method=0, gravity=(0.0, -9.81, 0.0)):
        g = _f32(gravity)
        _ck(lib().pbd_set_params(self._h, float(dt), int(sub_steps), int(max_iter), int(vel_method), _ptr(g)))

    def set_mode(self, mode):
        _ck(lib().pbd_set_mode(self._h, int(mode)))

    def get_mode(self):
        """(requested, active): active is what the current image runs in (the resolution of MODE_AUTO after the first step)."""
        r, a = C.c_int(0), C.c_int(0)
        _ck(lib().pbd_get_mode(self._h, C.byref(r), C.byref(a)))
        return r.value, a.value

    def set_bucket_sort(self, enable):
        _ck(lib().pbd_set_bucket_sort(self._h, int(bool(enable))))

    def step(self, n=1):
        _ck(lib().pbd_step(self._h, int(n)))

    def sync(self):
        _ck(lib().pbd_sync(self._h))

    def step_host(self, n, x_in, v_in, x_out, v_out=None):
        _ck(lib().pbd_step_host(self._h, int(n), _ptr(x_in), _ptr(v_in), _ptr(x_out), _ptr(v_out)))

    def step_host_async(self, n, x_in, v_in, x_out, v_out=None):
        """pbd_step_host_async: enqueue only; the arrays belong to the engine until step_host_wait covers the call."""
        _ck(lib().pbd_step_host_async(self._h, int(n), _ptr(x_in), _ptr(v_in), _ptr(x_out), _ptr(v_out)))

    def step_host_wait(self, lag=0):
        _ck(lib().pbd_step_host_wait(self._h, int(lag)))

    def set_colliders(self, particle_colliders, rigid_colliders):
        """pbd_set_colliders: lists of ParticleCollider / RigidCollider (rigid ones in the order of the reference's collision object list)."""
        pa = (ParticleCollider * max(len(particle_colliders), 1))(*particle_colliders)
        ra = (RigidCollider * max(len(rigid_colliders), 1))(*rigid_colliders)
        _ck(lib().pbd_set_colliders(self._h, len(particle_colliders), C.cast(pa, C.c_void_p), len(rigid_colliders), C.cast(ra, C.c_void_p)))

    def set_contact_params(self, tolerance=0.01, stiffness=100.0, max_iter_v=5):
        _ck(lib().pbd_set_contact_params(self._h, float(tolerance), float(stiffness), int(max_iter_v)))

    def record_contacts(self, capacity):
        _ck(lib().pbd_record_contacts(self._h, int(capacity)))

    def contacts(self, capacity=1 << 16):
        buf = (Contact * capacity)(); cnt = C.c_uint(0)
        _ck(lib().pbd_get_contacts(self._h, C.cast(buf, C.c_void_p), capacity, C.byref(cnt)))
        return [buf[i] for i in range(min(cnt.value, capacity))], cnt.value

    def lambdas(self, ctype):
        cnt = self.stats().constraints_per_type[ctype]
        lam = np.zeros(max(cnt, 1), dtype=np.float32); ids = np.zeros(max(cnt, 1), dtype=np.uint32)
        _ck(lib().pbd_get_lambdas(self._h, ctype, _ptr(lam), _ptr(ids)))
        return lam[:cnt], ids[:cnt]

    def stats(self, flatten=True):
        s = Stats()
        _ck(lib().pbd_get_stats(self._h, C.byref(s)))
        return s

    def profile_step(self):
        ms = np.zeros(NUM_TYPES, dtype=np.float32); launches = np.zeros(NUM_TYPES, dtype=np.uint32)
        mi = C.c_float(0); mv = C.c_float(0)
        _ck(lib().pbd_profile_step(self._h, _ptr(ms), C.byref(mi), C.byref(mv), _ptr(launches)))
        return ms, mi.value, mv.value, launches
