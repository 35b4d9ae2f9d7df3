"""Host model mirror over the C ABI of include/pbd_b200_model.h (ctypes; no compute happens in Python).

`HostModel` carries the snake_case builder surface that tests/scenes.py drives (the same surface the CPU checkers of the test suite expose), `TimeStep` wraps the engine-backed TimeStepController.  The camelCase API that mirrors
pyPBD one-to-one lives in positionbaseddynamics_b200/pypbd.py on top of these.
"""
import ctypes as C
import numpy as np
from . import _capi
from ._capi import PbdError, lib, Engine, Stats

MODEL_SYMBOLS = [
    "pbdm_model_create", "pbdm_model_destroy", "pbdm_model_reset", "pbdm_model_cleanup", "pbdm_add_regular_triangle_model",
    "pbdm_add_regular_tet_model", "pbdm_add_triangle_model", "pbdm_add_tet_model", "pbdm_num_particles", "pbdm_set_mass",
    "pbdm_get_mass", "pbdm_get_inv_mass", "pbdm_get_masses", "pbdm_get_particle", "pbdm_set_particle", "pbdm_get_particles", "pbdm_set_particles",
    "pbdm_vertices", "pbdm_add_rigid_body", "pbdm_num_rigid_bodies", "pbdm_set_rigid_body_mass", "pbdm_get_rigid_body_mass", "pbdm_get_rigid_bodies", "pbdm_add_constraint", "pbdm_add_cloth_constraints", "pbdm_add_bending_constraints",
    "pbdm_add_solid_constraints", "pbdm_num_constraints", "pbdm_get_constraint", "pbdm_get_constraints",
    "pbdm_init_constraint_groups", "pbdm_num_groups", "pbdm_get_groups", "pbdm_set_model_param", "pbdm_num_triangle_models",
    "pbdm_tri_num_edges", "pbdm_tri_num_faces", "pbdm_tri_index_offset", "pbdm_tri_get_edges", "pbdm_tri_get_faces",
    "pbdm_num_tet_models", "pbdm_tet_num_edges", "pbdm_tet_num_tets", "pbdm_tet_index_offset", "pbdm_tet_get_edges",
    "pbdm_tet_get_tets", "pbdm_first_fit_colouring", "pbdm_timestep_create", "pbdm_timestep_destroy", "pbdm_timestep_set_uint",
    "pbdm_timestep_get_uint", "pbdm_timestep_set_int", "pbdm_timestep_get_int", "pbdm_timestep_set_time_step_size",
    "pbdm_timestep_get_time_step_size", "pbdm_timestep_get_time", "pbdm_timestep_set_time", "pbdm_timestep_set_gravitation",
    "pbdm_timestep_set_mode", "pbdm_timestep_step", "pbdm_timestep_error", "pbdm_timestep_engine",
    "pbdm_cd_create", "pbdm_cd_destroy", "pbdm_cd_set_tolerance", "pbdm_cd_get_tolerance", "pbdm_cd_add_collision_shape",
    "pbdm_cd_add_collision_object_without_geometry", "pbdm_cd_num_collision_objects", "pbdm_set_contact_coefficients",
    "pbdm_set_contact_stiffness_particle_rigid_body", "pbdm_timestep_set_collision_detection", "pbdm_set_rigid_body_geometry_frame"]

_F = C.c_float
_vp = C.c_void_p
_configured = False


def _l():
    global _configured
    L = lib()
    if not _configured:
        L.pbdm_model_create.restype = _vp
        L.pbdm_timestep_create.restype = _vp
        L.pbdm_timestep_create.argtypes = [C.c_int, _vp]
        L.pbdm_timestep_engine.restype = _vp
        L.pbdm_timestep_error.restype = C.c_char_p
        L.pbdm_vertices.restype = C.POINTER(C.c_float)
        for n in ("pbdm_get_mass", "pbdm_get_inv_mass", "pbdm_timestep_get_time_step_size", "pbdm_timestep_get_time"):
            getattr(L, n).restype = C.c_float
        for n in ("pbdm_num_particles", "pbdm_num_constraints", "pbdm_num_groups", "pbdm_num_triangle_models", "pbdm_num_tet_models",
                  "pbdm_tri_num_edges", "pbdm_tri_num_faces", "pbdm_tri_index_offset", "pbdm_tet_num_edges", "pbdm_tet_num_tets",
                  "pbdm_tet_index_offset", "pbdm_timestep_get_uint", "pbdm_first_fit_colouring"):
            getattr(L, n).restype = C.c_uint
        # first argument of every model / timestep call is an opaque pointer
        one = [_vp]
        L.pbdm_model_destroy.argtypes = one; L.pbdm_model_reset.argtypes = one; L.pbdm_model_cleanup.argtypes = one
        L.pbdm_add_regular_triangle_model.argtypes = [_vp, C.c_int, C.c_int, _vp, _vp, _vp]
        L.pbdm_add_regular_tet_model.argtypes = [_vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp]
        L.pbdm_add_triangle_model.argtypes = [_vp, C.c_uint, C.c_uint, _vp, _vp]
        L.pbdm_add_tet_model.argtypes = [_vp, C.c_uint, C.c_uint, _vp, _vp]
        L.pbdm_num_particles.argtypes = one
        L.pbdm_set_mass.argtypes = [_vp, C.c_uint, _F]
        L.pbdm_get_mass.argtypes = [_vp, C.c_uint]; L.pbdm_get_inv_mass.argtypes = [_vp, C.c_uint]
        L.pbdm_get_masses.argtypes = [_vp, _vp, _vp]
        L.pbdm_get_particle.argtypes = [_vp, C.c_int, C.c_uint, _vp]; L.pbdm_set_particle.argtypes = [_vp, C.c_int, C.c_uint, _vp]
        L.pbdm_get_particles.argtypes = [_vp, C.c_int, _vp]; L.pbdm_set_particles.argtypes = [_vp, C.c_int, _vp]
        L.pbdm_vertices.argtypes = one
        L.pbdm_add_constraint.argtypes = [_vp, C.c_int, _vp, _vp]
        L.pbdm_add_rigid_body.argtypes = [_vp, _F, _vp, _vp, _vp]; L.pbdm_add_rigid_body.restype = C.c_uint
        L.pbdm_num_rigid_bodies.argtypes = one; L.pbdm_num_rigid_bodies.restype = C.c_uint
        L.pbdm_get_rigid_bodies.argtypes = [_vp, _vp]
        L.pbdm_set_rigid_body_mass.argtypes = [_vp, C.c_uint, _F]
        L.pbdm_get_rigid_body_mass.argtypes = [_vp, C.c_uint]; L.pbdm_get_rigid_body_mass.restype = C.c_float
        L.pbdm_add_cloth_constraints.argtypes = [_vp, C.c_uint, C.c_uint, _F, _F, _F, _F, _F, _F, C.c_int, C.c_int]
        L.pbdm_add_bending_constraints.argtypes = [_vp, C.c_uint, C.c_uint, _F]
        L.pbdm_add_solid_constraints.argtypes = [_vp, C.c_uint, C.c_uint, _F, _F, _F, C.c_int, C.c_int]
        L.pbdm_num_constraints.argtypes = one
        L.pbdm_get_constraint.argtypes = [_vp, C.c_uint, _vp, _vp, _vp]
        L.pbdm_get_constraints.argtypes = [_vp, _vp, _vp, _vp]
        L.pbdm_init_constraint_groups.argtypes = one; L.pbdm_num_groups.argtypes = one
        L.pbdm_get_groups.argtypes = [_vp, _vp, _vp]
        L.pbdm_set_model_param.argtypes = [_vp, C.c_int, _F]
        L.pbdm_num_triangle_models.argtypes = one; L.pbdm_num_tet_models.argtypes = one
        for n in ("pbdm_tri_num_edges", "pbdm_tri_num_faces", "pbdm_tri_index_offset", "pbdm_tet_num_edges", "pbdm_tet_num_tets", "pbdm_tet_index_offset"):
            getattr(L, n).argtypes = [_vp, C.c_uint]
        for n in ("pbdm_tri_get_edges", "pbdm_tri_get_faces", "pbdm_tet_get_edges", "pbdm_tet_get_tets"):
            getattr(L, n).argtypes = [_vp, C.c_uint, _vp]
        L.pbdm_first_fit_colouring.argtypes = [C.c_uint, C.c_uint, _vp, _vp, _vp]
        L.pbdm_timestep_destroy.argtypes = one
        L.pbdm_timestep_set_uint.argtypes = [_vp, C.c_int, C.c_uint]; L.pbdm_timestep_get_uint.argtypes = [_vp, C.c_int]
        L.pbdm_timestep_set_int.argtypes = [_vp, C.c_int, C.c_int]; L.pbdm_timestep_get_int.argtypes = [_vp, C.c_int]
        L.pbdm_timestep_set_time_step_size.argtypes = [_vp, _F]; L.pbdm_timestep_get_time_step_size.argtypes = one
        L.pbdm_timestep_get_time.argtypes = one; L.pbdm_timestep_set_time.argtypes = [_vp, _F]
        L.pbdm_timestep_set_gravitation.argtypes = [_vp, _vp]; L.pbdm_timestep_set_mode.argtypes = [_vp, C.c_int]
        L.pbdm_timestep_step.argtypes = [_vp, _vp]; L.pbdm_timestep_error.argtypes = one; L.pbdm_timestep_engine.argtypes = one
        L.pbdm_cd_create.restype = _vp; L.pbdm_cd_destroy.argtypes = one
        L.pbdm_cd_set_tolerance.argtypes = [_vp, _F]; L.pbdm_cd_get_tolerance.argtypes = one; L.pbdm_cd_get_tolerance.restype = C.c_float
        L.pbdm_cd_add_collision_shape.argtypes = [_vp, C.c_uint, C.c_uint, C.c_int, _vp, _F, _vp, C.c_uint, C.c_int, C.c_int]
        L.pbdm_cd_add_collision_object_without_geometry.argtypes = [_vp, C.c_uint, C.c_uint, C.c_int]
        L.pbdm_cd_num_collision_objects.argtypes = one; L.pbdm_cd_num_collision_objects.restype = C.c_uint
        L.pbdm_set_contact_coefficients.argtypes = [_vp, C.c_int, C.c_uint, _F, _F]
        L.pbdm_set_contact_stiffness_particle_rigid_body.argtypes = [_vp, _F]
        L.pbdm_timestep_set_collision_detection.argtypes = [_vp, _vp, _vp]
        L.pbdm_set_rigid_body_geometry_frame.argtypes = [_vp, C.c_uint, _vp, _vp]
        _configured = True
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None else None


NUM_SUB_STEPS, MAX_ITERATIONS, MAX_ITERATIONS_V, VELOCITY_UPDATE_METHOD = 0, 1, 2, 3
RIGID_BODY_COLLISION_OBJECT, TRIANGLE_MODEL_COLLISION_OBJECT, TET_MODEL_COLLISION_OBJECT = 0, 1, 2  # CollisionObject::*CollisionObjectType


class CollisionDetection:
    """Host mirror of PBD::DistanceFieldCollisionDetection: the registry of collision objects (the tests run on the GPU)."""

    def __init__(self):
        self._h = _vp(_l().pbdm_cd_create())

    def close(self):
        if self._h:
            _l().pbdm_cd_destroy(self._h); self._h = _vp()

    def set_tolerance(self, t): _l().pbdm_cd_set_tolerance(self._h, float(t))
    def get_tolerance(self): return _l().pbdm_cd_get_tolerance(self._h)
    def num_collision_objects(self): return _l().pbdm_cd_num_collision_objects(self._h)

    def add_shape(self, body_index, body_type, shape, dims, thickness=0.0, vertices=None, test_mesh=True, invert_sdf=False):
        d = np.zeros(3, np.float32); dd = np.atleast_1d(np.asarray(dims, dtype=np.float32)); d[:len(dd)] = dd
        v = _f32(vertices).reshape(-1, 3) if vertices is not None else None
        if _l().pbdm_cd_add_collision_shape(self._h, int(body_index), int(body_type), int(shape), _p(d), float(thickness), _p(v), 0 if v is None else len(v),
                                            int(bool(test_mesh)), int(bool(invert_sdf))):
            raise PbdError("unknown collision shape %r" % (shape,))

    def add_object_without_geometry(self, body_index, body_type, test_mesh=True):
        _l().pbdm_cd_add_collision_object_without_geometry(self._h, int(body_index), int(body_type), int(bool(test_mesh)))



def first_fit_colouring(num_bodies, body_off, bodies):
    body_off = np.ascontiguousarray(body_off, dtype=np.uint32); bodies = np.ascontiguousarray(bodies, dtype=np.uint32)
    out = np.zeros(max(len(body_off) - 1, 1), dtype=np.uint32)
    n = _l().pbdm_first_fit_colouring(int(num_bodies), len(body_off) - 1, _p(body_off), _p(bodies), _p(out))
    return n, out[:len(body_off) - 1]


class HostModel:
    """SimulationModel mirror (host C++); scene construction needs no GPU."""
    ATTR = {"x": 0, "v": 1, "x0": 2, "oldX": 3, "lastX": 4}

    def __init__(self):
        self._h = _vp(_l().pbdm_model_create())
        self._ts = None
        self._params = dict(dt=0.005, sub_steps=5, max_iter=1, vel_method=0, gravity=(0.0, -9.81, 0.0))

    def close(self):
        if self._ts is not None:
            self._ts.close(); self._ts = None
        if self._h:
            _l().pbdm_model_destroy(self._h); self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- builder surface (snake_case; tests/scenes.py drives it) ---------------------------------------
    def add_regular_triangle_model(self, w, h, t=(0, 0, 0), R=np.eye(3), scale=(1, 1)):
        _l().pbdm_add_regular_triangle_model(self._h, w, h, _p(_f32(t)), _p(_f32(R)), _p(_f32(scale)))

    def add_regular_tet_model(self, w, h, d, t=(0, 0, 0), R=np.eye(3), scale=(1, 1, 1)):
        _l().pbdm_add_regular_tet_model(self._h, w, h, d, _p(_f32(t)), _p(_f32(R)), _p(_f32(scale)))

    def add_triangle_model(self, pts, faces):
        pts = _f32(pts); faces = np.ascontiguousarray(faces, dtype=np.uint32)
        _l().pbdm_add_triangle_model(self._h, len(pts), len(faces), _p(pts), _p(faces))

    def add_tet_model(self, pts, tets):
        pts = _f32(pts); tets = np.ascontiguousarray(tets, dtype=np.uint32)
        _l().pbdm_add_tet_model(self._h, len(pts), len(tets), _p(pts), _p(tets))

    def set_mass(self, i, m):
        _l().pbdm_set_mass(self._h, int(i), float(m))

    def add_cloth_constraints(self, tm, method, dist_k=1.0, xx=1.0, yy=1.0, xy=1.0, pxy=0.3, pyx=0.3, norm_stretch=False, norm_shear=False):
        _l().pbdm_add_cloth_constraints(self._h, tm, method, dist_k, xx, yy, xy, pxy, pyx, int(norm_stretch), int(norm_shear))

    def add_bending_constraints(self, tm, method, k):
        _l().pbdm_add_bending_constraints(self._h, tm, method, k)

    def add_solid_constraints(self, tm, method, k=1.0, nu=0.3, vol_k=1.0, norm_stretch=False, norm_shear=False):
        _l().pbdm_add_solid_constraints(self._h, tm, method, k, nu, vol_k, int(norm_stretch), int(norm_shear))

    def add_constraint(self, ctype, bodies, args):
        b = np.zeros(4, dtype=np.uint32); b[:len(bodies)] = bodies
        a = np.zeros(8, dtype=np.float32); a[:len(args)] = args
        return _l().pbdm_add_constraint(self._h, ctype, _p(b), _p(a))

    def set_contact_coefficients(self, kind, index, restitution, friction):
        """kind 0 rigid body, 1 triangle model, 2 tet model (setRestitutionCoeff / setFrictionCoeff of the three classes)."""
        if _l().pbdm_set_contact_coefficients(self._h, int(kind), int(index), float(restitution), float(friction)):
            raise PbdError("no such object: kind %d index %d" % (kind, index))

    def set_rigid_body_geometry_frame(self, i, R, t):
        """Frame of the body's geometry (principal-axes matrix R, centre of mass t in geometry coordinates): x_local = R R(q)^T (x_w - x) + t."""
        if _l().pbdm_set_rigid_body_geometry_frame(self._h, int(i), _p(_f32(R).reshape(3, 3)), _p(_f32(t))):
            raise PbdError("rigid body index %d out of range" % i)

    def set_contact_stiffness_particle_rigid_body(self, k):
        _l().pbdm_set_contact_stiffness_particle_rigid_body(self._h, float(k))

    def add_rigid_body(self, mass, x, inertia, q=(1, 0, 0, 0)):
        return _l().pbdm_add_rigid_body(self._h, float(mass), _p(_f32(x)), _p(_f32(inertia)), _p(_f32(q)))

    def set_rigid_body_mass(self, i, mass):
        if _l().pbdm_set_rigid_body_mass(self._h, int(i), float(mass)):
            raise PbdError("rigid body index %d out of range" % i)

    def rigid_body_mass(self, i):
        return float(_l().pbdm_get_rigid_body_mass(self._h, int(i)))

    def add_ball_joint(self, rb0, rb1, pos):
        return self.add_constraint(_capi.BALLJOINT, [rb0, rb1], list(pos))

    def add_rb_particle_ball_joint(self, rb, particle):
        return self.add_constraint(_capi.RB_PARTICLE_BALLJOINT, [rb, particle], [])

    def rigid_bodies(self):
        n = _l().pbdm_num_rigid_bodies(self._h)
        out = np.zeros((max(n, 1), 13), dtype=np.float32); _l().pbdm_get_rigid_bodies(self._h, _p(out)); return out[:n]

    def set_params(self, dt=0.005, sub_steps=5, max_iter=1, vel_method=0, gravity=(0, -9.81, 0)):
        self._params = dict(dt=dt, sub_steps=sub_steps, max_iter=max_iter, vel_method=vel_method, gravity=tuple(gravity))
        if self._ts is not None:
            self._ts.apply(self._params)

    # -- structure -----------------------------------------------------------------------------------
    def num_particles(self):
        return _l().pbdm_num_particles(self._h)

    def num_constraints(self):
        return _l().pbdm_num_constraints(self._h)

    def init_groups(self):
        _l().pbdm_init_constraint_groups(self._h)

    def groups(self):
        ng, nc = _l().pbdm_num_groups(self._h), self.num_constraints()
        off = np.zeros(ng + 1, dtype=np.uint32); ids = np.zeros(max(nc, 1), dtype=np.uint32)
        _l().pbdm_get_groups(self._h, _p(off), _p(ids))
        return off, ids[:nc]

    def constraints(self):
        nc = self.num_constraints()
        types = np.zeros(max(nc, 1), dtype=np.int32); bodies = np.zeros((max(nc, 1), 4), dtype=np.uint32)
        params = np.zeros((max(nc, 1), 24), dtype=np.float32)
        _l().pbdm_get_constraints(self._h, _p(types), _p(bodies), _p(params))
        nb = np.array([_capi.num_bodies(int(t)) for t in range(_capi.NUM_TYPES)], dtype=np.int32)[types[:nc]] if nc else np.zeros(0, np.int32)
        return types[:nc], bodies[:nc], params[:nc], nb

    def tri_edges(self, tm=0):
        out = np.zeros((_l().pbdm_tri_num_edges(self._h, tm), 4), dtype=np.uint32); _l().pbdm_tri_get_edges(self._h, tm, _p(out)); return out

    def tri_faces(self, tm=0):
        out = np.zeros((_l().pbdm_tri_num_faces(self._h, tm), 3), dtype=np.uint32); _l().pbdm_tri_get_faces(self._h, tm, _p(out)); return out

    def tet_edges(self, tm=0):
        out = np.zeros((_l().pbdm_tet_num_edges(self._h, tm), 2), dtype=np.uint32); _l().pbdm_tet_get_edges(self._h, tm, _p(out)); return out

    def tet_tets(self, tm=0):
        out = np.zeros((_l().pbdm_tet_num_tets(self._h, tm), 4), dtype=np.uint32); _l().pbdm_tet_get_tets(self._h, tm, _p(out)); return out

    def tri_index_offset(self, tm=0):
        return _l().pbdm_tri_index_offset(self._h, tm)

    def tet_index_offset(self, tm=0):
        return _l().pbdm_tet_index_offset(self._h, tm)

    def set_model_param(self, which, value):
        if _l().pbdm_set_model_param(self._h, int(which), float(value)):
            raise PbdError("unknown model parameter %r" % which)

    # -- state ---------------------------------------------------------------------------------------
    def get(self, name="x"):
        out = np.zeros((self.num_particles(), 3), dtype=np.float32)
        if _l().pbdm_get_particles(self._h, self.ATTR[name], _p(out)):
            raise PbdError("bad attribute")
        return out

    def set(self, name, arr):
        arr = _f32(arr); assert arr.shape == (self.num_particles(), 3)
        _l().pbdm_set_particles(self._h, self.ATTR[name], _p(arr))

    def masses(self):
        n = self.num_particles()
        m = np.zeros(max(n, 1), dtype=np.float32); w = np.zeros(max(n, 1), dtype=np.float32)
        _l().pbdm_get_masses(self._h, _p(m), _p(w))
        return m[:n], w[:n]

    def vertices_view(self):
        """Zero-copy numpy view of the host positions (pyPBD getVertices); pulls from the device first."""
        n = self.num_particles()
        ptr = _l().pbdm_vertices(self._h)
        return np.ctypeslib.as_array(ptr, shape=(n, 3)) if n else np.zeros((0, 3), np.float32)

    # -- stepping (needs a CUDA device) --------------------------------------------------------------
    def time_step(self, device=0, stream=None):
        if self._ts is None:
            self._ts = TimeStep(device, stream)
            self._ts.apply(self._params)
        return self._ts

    def step(self, n=1, device=0):
        ts = self.time_step(device)
        for _ in range(int(n)):
            ts.step(self)


class TimeStep:
    """Engine-backed TimeStepController (+ TimeManager, gravity)."""

    def __init__(self, device=0, stream=None):
        h = _l().pbdm_timestep_create(int(device), _vp(stream) if stream else None)
        if not h:
            raise PbdError(lib().pbd_last_error().decode())
        self._h = _vp(h)

    def close(self):
        if self._h:
            _l().pbdm_timestep_destroy(self._h); self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def apply(self, p):
        self.set_uint(NUM_SUB_STEPS, p["sub_steps"]); self.set_uint(MAX_ITERATIONS, p["max_iter"])
        self.set_int(VELOCITY_UPDATE_METHOD, p["vel_method"]); self.set_time_step_size(p["dt"]); self.set_gravitation(p["gravity"])

    def set_uint(self, pid, v):
        if _l().pbdm_timestep_set_uint(self._h, pid, int(v)):
            raise PbdError("parameter %d rejects value %r" % (pid, v))

    def get_uint(self, pid):
        return _l().pbdm_timestep_get_uint(self._h, pid)

    def set_int(self, pid, v):
        if _l().pbdm_timestep_set_int(self._h, pid, int(v)):
            raise PbdError("parameter %d rejects value %r" % (pid, v))

    def get_int(self, pid):
        return _l().pbdm_timestep_get_int(self._h, pid)

    def set_time_step_size(self, h):
        _l().pbdm_timestep_set_time_step_size(self._h, float(h))

    def get_time_step_size(self):
        return _l().pbdm_timestep_get_time_step_size(self._h)

    def get_time(self):
        return _l().pbdm_timestep_get_time(self._h)

    def set_time(self, t):
        _l().pbdm_timestep_set_time(self._h, float(t))

    def set_gravitation(self, g):
        _l().pbdm_timestep_set_gravitation(self._h, _p(_f32(g)))

    def set_mode(self, mode):
        _l().pbdm_timestep_set_mode(self._h, int(mode))

    def set_collision_detection(self, model, cd):
        """TimeStep::setCollisionDetection; cd = None detaches."""
        self._cd = cd  # keep it alive
        _l().pbdm_timestep_set_collision_detection(self._h, model._h, cd._h if cd is not None else None)

    def step(self, model):
        if _l().pbdm_timestep_step(self._h, model._h):
            raise PbdError(_l().pbdm_timestep_error(self._h).decode())

    def engine_handle(self):
        return _vp(_l().pbdm_timestep_engine(self._h))

    def stats(self):
        s = Stats()
        if lib().pbd_get_stats(self.engine_handle(), C.byref(s)):
            raise PbdError(lib().pbd_last_error().decode())
        return s

    def sync(self):
        if lib().pbd_sync(self.engine_handle()):
            raise PbdError(lib().pbd_last_error().decode())
