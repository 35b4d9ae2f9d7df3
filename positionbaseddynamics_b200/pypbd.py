"""pyPBD-compatible facade (camelCase names of the reference's Python module `pypbd`) over the host model mirror.

Covers the calls the reference's own example scripts make for the accelerated path
(pyPBD/examples/cloth_model.py:18-124, beam_model.py:15-87):
  Simulation.getCurrent/initDefault/getModel/getTimeStep/reset, TimeManager.getCurrent().getTime/setTimeStepSize,
  SimulationModel.addRegularTriangleModel/addRegularTetModel/addTriangleModel/addTetModel/addClothConstraints/
  addBendingConstraints/addSolidConstraints/add<X>Constraint/getParticles/getTriangleModels/getTetModels/getConstraints/
  getConstraintGroups/cleanup, ParticleData.setMass/getMass/getPosition/setPosition/getVertices/size,
  TimeStepController.NUM_SUB_STEPS/MAX_ITERATIONS/..., ts.setValueUInt/getValueUInt/setValueInt, ts.step(model).
GUI, scene files, rigid bodies and collision detection are not part of this path.
"""
import math
import numpy as np
from . import _capi, model as _m
from ._capi import PbdError  # noqa: F401


class ParticleData:
    def __init__(self, host):
        self._h = host

    def size(self):
        return self._h.num_particles()

    getNumberOfParticles = size

    def setMass(self, i, mass):
        self._h.set_mass(i, mass)

    def getMass(self, i):
        return float(_m._l().pbdm_get_mass(self._h._h, int(i)))

    def getInvMass(self, i):
        return float(_m._l().pbdm_get_inv_mass(self._h._h, int(i)))

    def _get(self, attr, i):
        out = np.zeros(3, dtype=np.float32)
        if _m._l().pbdm_get_particle(self._h._h, attr, int(i), _m._p(out)):
            raise IndexError(i)
        return out

    def _set(self, attr, i, v):
        v = _m._f32(v)
        if _m._l().pbdm_set_particle(self._h._h, attr, int(i), _m._p(v)):
            raise IndexError(i)

    def getPosition(self, i): return self._get(_capi.ATTR_X, i)
    def getPosition0(self, i): return self._get(_capi.ATTR_X0, i)
    def getVelocity(self, i): return self._get(_capi.ATTR_V, i)
    def getOldPosition(self, i): return self._get(_capi.ATTR_OLDX, i)
    def getLastPosition(self, i): return self._get(_capi.ATTR_LASTX, i)
    def setPosition(self, i, v): self._set(_capi.ATTR_X, i, v)
    def setPosition0(self, i, v): self._set(_capi.ATTR_X0, i, v)
    def setVelocity(self, i, v): self._set(_capi.ATTR_V, i, v)

    def getVertices(self):
        """Zero-copy view of the positions (pyPBD/ParticleDataModule.cpp:54-58); pulled from the device lazily."""
        return self._h.vertices_view()


class _Mesh:
    def __init__(self, host, idx, tri):
        self._h, self._i, self._tri = host, idx, tri

    def numFaces(self): return _m._l().pbdm_tri_num_faces(self._h._h, self._i)
    def numEdges(self): return (_m._l().pbdm_tri_num_edges if self._tri else _m._l().pbdm_tet_num_edges)(self._h._h, self._i)
    def numTets(self): return _m._l().pbdm_tet_num_tets(self._h._h, self._i)
    def getFaces(self): return self._h.tri_faces(self._i).reshape(-1)
    def getTets(self): return self._h.tet_tets(self._i).reshape(-1)
    def getEdges(self): return self._h.tri_edges(self._i) if self._tri else self._h.tet_edges(self._i)


class _Coefficients:
    """setRestitutionCoeff / setFrictionCoeff of RigidBody, TriangleModel and TetModel (defaults 0.6 / 0.2 as in the reference)."""
    _kind = 0

    def _coeffs(self):
        return self._h.__dict__.setdefault("_contact_coeffs", {}).setdefault((self._kind, self._i), [0.6, 0.2])

    def getRestitutionCoeff(self): return self._coeffs()[0]
    def getFrictionCoeff(self): return self._coeffs()[1]

    def setRestitutionCoeff(self, v):
        c = self._coeffs(); c[0] = float(v); self._h.set_contact_coefficients(self._kind, self._i, c[0], c[1])

    def setFrictionCoeff(self, v):
        c = self._coeffs(); c[1] = float(v); self._h.set_contact_coefficients(self._kind, self._i, c[0], c[1])


class TriangleModel(_Coefficients):
    _kind = 1
    def __init__(self, host, idx): self._h, self._i = host, idx
    def getIndexOffset(self): return self._h.tri_index_offset(self._i)
    def getParticleMesh(self): return _Mesh(self._h, self._i, True)
    def updateMeshNormals(self, pd): pass  # rendering helper of the reference; no normals are kept on this path


class TetModel(_Coefficients):
    _kind = 2
    def __init__(self, host, idx): self._h, self._i = host, idx
    def getIndexOffset(self): return self._h.tet_index_offset(self._i)
    def getParticleMesh(self): return _Mesh(self._h, self._i, False)
    def updateMeshNormals(self, pd): pass  # rendering helper of the reference


def mass_properties(vertices, faces, density):
    """Mass, centre of mass and inertia tensor (about the centre of mass) of the closed triangle mesh, as
    Utilities::VolumeIntegration (Utils/VolumeIntegration.cpp) provides them to RigidBody::determineMassProperties
    (Simulation/RigidBody.h:209-262): exact polyhedral integrals, here summed over signed tetrahedra from the origin."""
    v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3); f = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    vol6 = np.einsum("ij,ij->i", a, np.cross(b, c))              # 6 x signed tetra volume
    vol = vol6.sum() / 6.0
    com = (vol6[:, None] * (a + b + c)).sum(axis=0) / (24.0 * vol)
    s = a + b + c
    cov = (vol6[:, None, None] * (s[:, :, None] * s[:, None, :] + a[:, :, None] * a[:, None, :] + b[:, :, None] * b[:, None, :]
                                  + c[:, :, None] * c[:, None, :])).sum(axis=0) / 120.0   # integral of x x^T
    cov -= vol * np.outer(com, com)
    inertia = density * (np.trace(cov) * np.eye(3) - cov)
    return density * vol, com, inertia


class RigidBody(_Coefficients):
    """View of one rigid body of the model (pyPBD RigidBodyModule.cpp subset: what the coupling example touches)."""
    def __init__(self, host, idx): self._h, self._i = host, idx
    def _row(self): return self._h.rigid_bodies()[self._i]
    def setMass(self, mass): self._h.set_rigid_body_mass(self._i, mass)
    def getMass(self): return self._h.rigid_body_mass(self._i)
    def getPosition(self): return self._row()[0:3].copy()
    def getRotation(self): return self._row()[3:7].copy()        # (w, x, y, z)
    def getVelocity(self): return self._row()[7:10].copy()
    def getAngularVelocity(self): return self._row()[10:13].copy()


class SimulationModel:
    def __init__(self):
        self._host = _m.HostModel()

    def init(self): pass
    def reset(self): _m._l().pbdm_model_reset(self._host._h)
    def cleanup(self): _m._l().pbdm_model_cleanup(self._host._h)
    def getParticles(self): return ParticleData(self._host)
    def getTriangleModels(self): return [TriangleModel(self._host, i) for i in range(_m._l().pbdm_num_triangle_models(self._host._h))]
    def getTetModels(self): return [TetModel(self._host, i) for i in range(_m._l().pbdm_num_tet_models(self._host._h))]

    # The builders return the new model like pyPBD does (SimulationModelModule.cpp:98-229).  uvIndices / uvs (rendering) and
    # testMesh (collision detection against signed distance fields) belong to parts of the reference outside this path: accepted
    # for source compatibility, testMesh=True is refused because no collision detection runs here.
    @staticmethod
    def _no_collision(testMesh):
        if testMesh:
            raise _capi.PbdError("testMesh=True needs the reference's collision detection, which is outside this engine's path")

    def addRegularTriangleModel(self, width, height, translation=(0, 0, 0), rotation=np.eye(3), scale=(1, 1), testMesh=False):
        self._no_collision(testMesh)
        self._host.add_regular_triangle_model(width, height, translation, rotation, scale)
        return self.getTriangleModels()[-1]

    def addRegularTetModel(self, width, height, depth, translation=(0, 0, 0), rotation=np.eye(3), scale=(1, 1, 1), testMesh=False):
        self._no_collision(testMesh)
        self._host.add_regular_tet_model(width, height, depth, translation, rotation, scale)
        return self.getTetModels()[-1]

    def addTriangleModel(self, points, indices, uvIndices=None, uvs=None, testMesh=False):
        self._no_collision(testMesh)
        self._host.add_triangle_model(np.asarray(points).reshape(-1, 3), np.asarray(indices).reshape(-1, 3))
        return self.getTriangleModels()[-1]

    def addTetModel(self, points, indices, testMesh=False, **unused):
        self._no_collision(testMesh)
        self._host.add_tet_model(np.asarray(points).reshape(-1, 3), np.asarray(indices).reshape(-1, 4))
        return self.getTetModels()[-1]

    def addClothConstraints(self, tm, clothMethod, distanceStiffness, xxStiffness, yyStiffness, xyStiffness, xyPoissonRatio, yxPoissonRatio,
                            normalizeStretch, normalizeShear):
        self._host.add_cloth_constraints(tm._i, clothMethod, distanceStiffness, xxStiffness, yyStiffness, xyStiffness, xyPoissonRatio,
                                         yxPoissonRatio, normalizeStretch, normalizeShear)

    def addBendingConstraints(self, tm, bendingMethod, stiffness):
        self._host.add_bending_constraints(tm._i, bendingMethod, stiffness)

    def addSolidConstraints(self, tm, solidMethod, stiffness, poissonRatio, volumeStiffness, normalizeStretch, normalizeShear):
        self._host.add_solid_constraints(tm._i, solidMethod, stiffness, poissonRatio, volumeStiffness, normalizeStretch, normalizeShear)

    # ---- rigid bodies coupled through ball joints (pyPBD SimulationModelModule.cpp:306-376, 396-420) -------------------
    def addRigidBody(self, density, vertices, mesh, translation=(0, 0, 0), rotation=np.eye(3), scale=(1, 1, 1), testMesh=False,
                     generateCollisionObject=False, resolution=None, sdf=None):
        """RigidBody::initBody(density, x, rotation, vertices, mesh, scale): mass and principal inertia from the scaled mesh, the body
        frame moved to the centre of mass and rotated into the principal axes (RigidBody.h:122-262).  `vertices` is an (n, 3) array,
        `mesh` an (m, 3) face array (or an object with getFaces()).  Collision objects / signed distance fields are outside this
        engine's path: requesting them is refused.  For repeated eigenvalues (cube, sphere) the principal frame is not unique; it may
        differ from Eigen's choice without changing the dynamics."""
        self._no_collision(testMesh)
        if generateCollisionObject or sdf is not None:
            raise _capi.PbdError("collision objects need the reference's collision detection, which is outside this engine's path")
        faces = np.asarray(mesh.getFaces() if hasattr(mesh, "getFaces") else mesh).reshape(-1, 3)
        v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3) * np.asarray(scale, dtype=np.float64)
        mass, com, J = mass_properties(v, faces, float(density))
        w, R = np.linalg.eigh(J)
        if np.linalg.det(R) < 0.0:
            R = -R
        R0 = np.asarray(rotation, dtype=np.float64).reshape(3, 3)
        x = R0 @ com + np.asarray(translation, dtype=np.float64)
        Rw = R0 @ R
        # rotation matrix -> unit quaternion (w, x, y, z)
        q = np.empty(4); t = np.trace(Rw)
        if t > 0.0:
            r = math.sqrt(1.0 + t); q[0] = 0.5 * r; r = 0.5 / r
            q[1] = (Rw[2, 1] - Rw[1, 2]) * r; q[2] = (Rw[0, 2] - Rw[2, 0]) * r; q[3] = (Rw[1, 0] - Rw[0, 1]) * r
        else:
            i = int(np.argmax(np.diag(Rw))); j = (i + 1) % 3; k = (i + 2) % 3
            r = math.sqrt(1.0 + Rw[i, i] - Rw[j, j] - Rw[k, k]); q[1 + i] = 0.5 * r; r = 0.5 / r
            q[0] = (Rw[k, j] - Rw[j, k]) * r; q[1 + j] = (Rw[j, i] + Rw[i, j]) * r; q[1 + k] = (Rw[k, i] + Rw[i, k]) * r
        idx = self._host.add_rigid_body(mass, x, w, q / np.linalg.norm(q))
        self._host.set_rigid_body_geometry_frame(idx, R, com)  # distance fields on the body live in the mesh's coordinates
        return RigidBody(self._host, idx)

    def getRigidBodies(self): return [RigidBody(self._host, i) for i in range(len(self._host.rigid_bodies()))]
    def addBallJoint(self, rbIndex1, rbIndex2, pos): return bool(self._host.add_ball_joint(rbIndex1, rbIndex2, pos))
    def addRigidBodyParticleBallJoint(self, rbIndex, particleIndex): return bool(self._host.add_rb_particle_ball_joint(rbIndex, particleIndex))

    def addDistanceConstraint(self, p1, p2, k): return bool(self._host.add_constraint(_capi.DISTANCE, [p1, p2], [k]))
    def addDistanceConstraint_XPBD(self, p1, p2, k): return bool(self._host.add_constraint(_capi.DISTANCE_XPBD, [p1, p2], [k]))
    def addDihedralConstraint(self, p1, p2, p3, p4, k): return bool(self._host.add_constraint(_capi.DIHEDRAL, [p1, p2, p3, p4], [k]))
    def addIsometricBendingConstraint(self, p1, p2, p3, p4, k): return bool(self._host.add_constraint(_capi.ISOBENDING, [p1, p2, p3, p4], [k]))
    def addIsometricBendingConstraint_XPBD(self, p1, p2, p3, p4, k): return bool(self._host.add_constraint(_capi.ISOBENDING_XPBD, [p1, p2, p3, p4], [k]))
    def addFEMTriangleConstraint(self, p1, p2, p3, xx, yy, xy, nuxy, nuyx): return bool(self._host.add_constraint(_capi.FEMTRIANGLE, [p1, p2, p3], [xx, yy, xy, nuxy, nuyx]))
    def addStrainTriangleConstraint(self, p1, p2, p3, xx, yy, xy, ns, nsh): return bool(self._host.add_constraint(_capi.STRAINTRIANGLE, [p1, p2, p3], [xx, yy, xy, float(ns), float(nsh)]))
    def addVolumeConstraint(self, p1, p2, p3, p4, k): return bool(self._host.add_constraint(_capi.VOLUME, [p1, p2, p3, p4], [k]))
    def addVolumeConstraint_XPBD(self, p1, p2, p3, p4, k): return bool(self._host.add_constraint(_capi.VOLUME_XPBD, [p1, p2, p3, p4], [k]))
    def addFEMTetConstraint(self, p1, p2, p3, p4, k, nu): return bool(self._host.add_constraint(_capi.FEMTET, [p1, p2, p3, p4], [k, nu]))
    def addFEMTetConstraint_XPBD(self, p1, p2, p3, p4, k, nu): return bool(self._host.add_constraint(_capi.FEMTET_XPBD, [p1, p2, p3, p4], [k, nu]))
    def addStrainTetConstraint(self, p1, p2, p3, p4, ks, kh, ns, nsh): return bool(self._host.add_constraint(_capi.STRAINTET, [p1, p2, p3, p4], [ks, kh, float(ns), float(nsh)]))

    def addShapeMatchingConstraint(self, n, indices, numClusters, k):
        return bool(self._host.add_constraint(_capi.SHAPEMATCHING, list(indices)[:4], [k] + [float(c) for c in list(numClusters)[:4]])) if n == 4 else False

    def initConstraintGroups(self): self._host.init_groups()

    def getConstraintGroups(self):
        off, ids = self._host.groups()
        return [ids[off[g]:off[g + 1]] for g in range(len(off) - 1)]

    def getConstraints(self):
        t, b, p, nb = self._host.constraints()
        return [dict(type=_capi.TYPE_NAMES[int(t[i])], bodies=b[i][:nb[i]], params=p[i][:_capi.num_params(int(t[i]))]) for i in range(len(t))]

    def numConstraints(self): return self._host.num_constraints()

    # global setters (Simulation/SimulationModel.cpp:1351-1485)
    def setClothStiffness(self, v): self._host.set_model_param(0, v)
    def setClothStiffnessXX(self, v): self._host.set_model_param(1, v)
    def setClothStiffnessYY(self, v): self._host.set_model_param(2, v)
    def setClothStiffnessXY(self, v): self._host.set_model_param(3, v)
    def setClothPoissonRatioXY(self, v): self._host.set_model_param(4, v)
    def setClothPoissonRatioYX(self, v): self._host.set_model_param(5, v)
    def setClothBendingStiffness(self, v): self._host.set_model_param(6, v)
    def setSolidStiffness(self, v): self._host.set_model_param(9, v)
    def setSolidPoissonRatio(self, v): self._host.set_model_param(10, v)
    def setSolidVolumeStiffness(self, v): self._host.set_model_param(11, v)


class LogLevel:
    """Utils/Logger.h levels (pyPBD exposes them as an enum)."""
    DEBUG, INFO, WARN, ERR = 0, 1, 2, 3


class Logger:
    """pyPBD's Logger.addConsoleSink / addFileSink (UtilitiesModule.cpp): this path has nothing of its own to log, the calls are
    kept so that the reference's example scripts run unchanged; errors surface as PbdError."""
    level = LogLevel.INFO

    @staticmethod
    def addConsoleSink(level=LogLevel.INFO): Logger.level = level

    @staticmethod
    def addFileSink(level=LogLevel.INFO, path=None): Logger.level = level


class Timing:
    """pyPBD's Timing.printAverageTimes / reset: the reference prints its START_TIMING sections (Utils/Timing.h); here the one
    section that exists is the step itself, timed on the device with CUDA events."""
    _steps = 0; _ms = 0.0
    enabled = False  # recording synchronises after every step; switch on with Timing.enabled = True

    @staticmethod
    def _record(ms): Timing._steps += 1; Timing._ms += ms

    @staticmethod
    def reset(): Timing._steps = 0; Timing._ms = 0.0

    @staticmethod
    def averageStepMs(): return Timing._ms / Timing._steps if Timing._steps else 0.0

    @staticmethod
    def printAverageTimes():
        print("---------------------------------------------------------------------------")
        print("Average times:")
        print("SimStep (device): %.4f ms over %d steps" % (Timing.averageStepMs(), Timing._steps))
        print("---------------------------------------------------------------------------")

    @staticmethod
    def printTimeSums(): print("SimStep (device): %.4f ms in %d steps" % (Timing._ms, Timing._steps))


class TimeManager:
    _current = None

    def __init__(self, ts): self._ts = ts

    @staticmethod
    def getCurrent(): return TimeManager(Simulation.getCurrent().getTimeStep())
    def getTime(self): return self._ts._ts.get_time()
    def setTime(self, t): self._ts._ts.set_time(t)
    def getTimeStepSize(self): return self._ts._ts.get_time_step_size()
    def setTimeStepSize(self, h): self._ts._ts.set_time_step_size(h)


class TimeStepController:
    NUM_SUB_STEPS, MAX_ITERATIONS, MAX_ITERATIONS_V, VELOCITY_UPDATE_METHOD = 0, 1, 2, 3
    ENUM_VUPDATE_FIRST_ORDER, ENUM_VUPDATE_SECOND_ORDER = 0, 1

    def __init__(self, device=0, stream=None):
        self._ts = _m.TimeStep(device, stream)   # raises PbdError without a CUDA device: no CPU fallback

    def init(self): pass
    def reset(self): pass
    def setValueUInt(self, pid, v): self._ts.set_uint(pid, v)
    def getValueUInt(self, pid): return self._ts.get_uint(pid)
    def setValueInt(self, pid, v): self._ts.set_int(pid, v)
    def getValueInt(self, pid): return self._ts.get_int(pid)
    def setCollisionDetection(self, model, cd):
        self._ts.set_collision_detection(model._host, cd._cd if cd is not None else None)

    def step(self, model):
        self._ts.step(model._host)
        if Timing.enabled:
            self._ts.sync(); Timing._record(self._ts.stats().last_step_ms)


from .loaders import TetGenLoader, OBJLoader, MeshFaceIndices, VertexData  # Utilities::TetGenLoader / OBJLoader under their pyPBD names


class CollisionObject:
    RigidBodyCollisionObjectType, TriangleModelCollisionObjectType, TetModelCollisionObjectType = 0, 1, 2


class DistanceFieldCollisionDetection:
    """pyPBD's DistanceFieldCollisionDetection (pyPBD/CollisionDetectionModule.cpp): the add* calls of the reference; the tests and the
    velocity-level contact solve run on the GPU for particles against static rigid bodies (include/pbd_b200.h "Contact path")."""

    def __init__(self):
        self._cd = _m.CollisionDetection()

    def init(self): pass
    def cleanup(self): pass
    def getTolerance(self): return self._cd.get_tolerance()
    def setTolerance(self, t): self._cd.set_tolerance(t)
    def numCollisionObjects(self): return self._cd.num_collision_objects()
    def addCollisionBox(self, bodyIndex, bodyType, vertices, numVertices, box, testMesh=True, invertSDF=False):
        self._cd.add_shape(bodyIndex, bodyType, _capi.SHAPE_BOX, box, 0.0, vertices, testMesh, invertSDF)
    def addCollisionSphere(self, bodyIndex, bodyType, vertices, numVertices, radius, testMesh=True, invertSDF=False):
        self._cd.add_shape(bodyIndex, bodyType, _capi.SHAPE_SPHERE, [radius], 0.0, vertices, testMesh, invertSDF)
    def addCollisionTorus(self, bodyIndex, bodyType, vertices, numVertices, radii, testMesh=True, invertSDF=False):
        self._cd.add_shape(bodyIndex, bodyType, _capi.SHAPE_TORUS, radii, 0.0, vertices, testMesh, invertSDF)
    def addCollisionCylinder(self, bodyIndex, bodyType, vertices, numVertices, dim, testMesh=True, invertSDF=False):
        self._cd.add_shape(bodyIndex, bodyType, _capi.SHAPE_CYLINDER, dim, 0.0, vertices, testMesh, invertSDF)
    def addCollisionHollowSphere(self, bodyIndex, bodyType, vertices, numVertices, radius, thickness, testMesh=True, invertSDF=False):
        self._cd.add_shape(bodyIndex, bodyType, _capi.SHAPE_HOLLOW_SPHERE, [radius], thickness, vertices, testMesh, invertSDF)
    def addCollisionHollowBox(self, bodyIndex, bodyType, vertices, numVertices, box, thickness, testMesh=True, invertSDF=False):
        self._cd.add_shape(bodyIndex, bodyType, _capi.SHAPE_HOLLOW_BOX, box, thickness, vertices, testMesh, invertSDF)
    def addCollisionObjectWithoutGeometry(self, bodyIndex, bodyType, vertices, numVertices, testMesh):
        self._cd.add_object_without_geometry(bodyIndex, bodyType, testMesh)


class Simulation:
    GRAVITATION = 0
    _current = None

    def __init__(self):
        self._model = None; self._ts = None; self._gravity = (0.0, -9.81, 0.0); self._device = 0

    @staticmethod
    def getCurrent():
        if Simulation._current is None:
            Simulation._current = Simulation()
        return Simulation._current

    @staticmethod
    def hasCurrent(): return Simulation._current is not None

    def initDefault(self, device=0):
        self._model = SimulationModel(); self._device = device

    def getModel(self): return self._model
    def setModel(self, m): self._model = m

    def getTimeStep(self):
        if self._ts is None:
            self._ts = TimeStepController(self._device)
            self._ts._ts.set_gravitation(self._gravity)
        return self._ts

    def setTimeStep(self, ts): self._ts = ts

    def setVecValueReal(self, pid, v):
        if pid == Simulation.GRAVITATION:
            self._gravity = tuple(float(c) for c in v)
            if self._ts is not None:
                self._ts._ts.set_gravitation(self._gravity)

    def getVecValueReal(self, pid): return list(self._gravity)

    def reset(self):
        if self._model is not None:
            self._model.reset()
        if self._ts is not None:
            self._ts._ts.set_time(0.0)
