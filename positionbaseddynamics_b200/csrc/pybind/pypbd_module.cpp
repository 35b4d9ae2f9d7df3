// positionbaseddynamics_b200/csrc/pybind/pypbd_module.cpp
//
// Compiled pybind11 module `pypbd_b200` (north_star: "pyPBD via pybind"): the names and call shapes of the reference's Python
// module for the accelerated path -- pyPBD/SimulationModule.cpp:14-45, SimulationModelModule.cpp:98-294, ParticleDataModule.cpp,
// TimeStepModule.cpp:15-31, TimeModule.cpp, UtilitiesModule.cpp -- bound directly onto the C++ host mirror (csrc/host/pbd_model.h)
// that drives libpbd_b200.so.  What the reference's example scripts call (pyPBD/examples/cloth_model.py:18-124,
// beam_model.py:15-87) works unchanged:
//     import pypbd_b200 as pbd
//     sim = pbd.Simulation.getCurrent(); sim.initDefault(); model = sim.getModel()
//     model.addRegularTriangleModel(...); model.addClothConstraints(tm, ...); ts = sim.getTimeStep()
//     ts.setValueUInt(pbd.TimeStepController.NUM_SUB_STEPS, 1); ts.step(model); model.getParticles().getVertices()
// GUI, scene files and collision detection are outside this path; the Python-level facade
// positionbaseddynamics_b200/pypbd.py offers the same surface over ctypes for environments without a compiler.
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>
#include <memory>
#include <stdexcept>
#include "../host/pbd_model.h"

namespace py = pybind11;
using namespace pbd_b200;

namespace {

Vector3r vec3(const py::object &o) {
    auto a = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(o);
    if (!a || a.size() != 3) throw std::invalid_argument("expected 3 numbers");
    return Vector3r((Real)a.data()[0], (Real)a.data()[1], (Real)a.data()[2]);
}
Matrix3r mat3(const py::object &o) {
    auto a = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(o);
    if (!a || a.size() != 9) throw std::invalid_argument("expected a 3x3 matrix");
    Matrix3r m;
    for (int i = 0; i < 9; i++) m.m[i] = (Real)a.data()[i];
    return m;
}
py::array_t<float> to_np(const Vector3r &v) { py::array_t<float> r(3); for (int k = 0; k < 3; k++) r.mutable_data()[k] = v[k]; return r; }

// the process-wide objects the reference keeps as singletons (Simulation.cpp:11,30-38, TimeManager.cpp:5,18-25)
struct Sim {
    std::unique_ptr<SimulationModel> model;
    std::unique_ptr<TimeStepController> ts;
    int device = 0;
    Vector3r gravity = Vector3r(0, (Real)-9.81, 0);
    static Sim &current() { static Sim s; return s; }
    TimeStepController *timeStep() {
        if (!ts) {
            ts.reset(new TimeStepController(device));
            if (!ts->valid()) { const std::string e = ts->error(); ts.reset(); throw std::runtime_error("pypbd_b200: " + e + " (no CPU fallback)"); }
            ts->setGravitation(gravity);
        }
        return ts.get();
    }
};
struct TimeManagerRef {};  // TimeManager.getCurrent(): forwards to the current time step's manager

}  // namespace

PYBIND11_MODULE(pypbd_b200, m) {
    m.doc() = "pyPBD-compatible compiled module over the B200 PBD/XPBD engine (libpbd_b200.so)";

    py::class_<ParticleData>(m, "ParticleData")
        .def("size", &ParticleData::size)
        .def("getNumberOfParticles", &ParticleData::getNumberOfParticles)
        .def("addVertex", [](ParticleData &pd, const py::object &v) { pd.addVertex(vec3(v)); })
        .def("getMass", &ParticleData::getMass)
        .def("getInvMass", &ParticleData::getInvMass)
        .def("setMass", &ParticleData::setMass)
        .def("getPosition", [](const ParticleData &pd, unsigned i) { if (i >= pd.size()) throw py::index_error(); return to_np(pd.getPosition(i)); })
        .def("getPosition0", [](const ParticleData &pd, unsigned i) { if (i >= pd.size()) throw py::index_error(); return to_np(pd.getPosition0(i)); })
        .def("getVelocity", [](const ParticleData &pd, unsigned i) { if (i >= pd.size()) throw py::index_error(); return to_np(pd.getVelocity(i)); })
        .def("getOldPosition", [](const ParticleData &pd, unsigned i) { if (i >= pd.size()) throw py::index_error(); return to_np(pd.getOldPosition(i)); })
        .def("getLastPosition", [](const ParticleData &pd, unsigned i) { if (i >= pd.size()) throw py::index_error(); return to_np(pd.getLastPosition(i)); })
        .def("setPosition", [](ParticleData &pd, unsigned i, const py::object &v) { if (i >= pd.size()) throw py::index_error(); pd.setPosition(i, vec3(v)); })
        .def("setPosition0", [](ParticleData &pd, unsigned i, const py::object &v) { if (i >= pd.size()) throw py::index_error(); pd.setPosition0(i, vec3(v)); })
        .def("setVelocity", [](ParticleData &pd, unsigned i, const py::object &v) { if (i >= pd.size()) throw py::index_error(); pd.setVelocity(i, vec3(v)); })
        // ParticleDataModule.cpp:54-58: a zero-copy (n, 3) float view of the positions; the host copy is refreshed from the device first
        .def("getVertices", [](py::object self) {
            const ParticleData &pd = self.cast<const ParticleData &>();
            const std::vector<Vector3r> &x = pd.getVertices();
            return py::array_t<float>({(py::ssize_t)x.size(), (py::ssize_t)3}, {(py::ssize_t)sizeof(Vector3r), (py::ssize_t)sizeof(float)},
                                      x.empty() ? nullptr : &x[0].v[0], self);
        });

    py::class_<IndexedFaceMesh>(m, "IndexedFaceMesh")
        .def("numVertices", &IndexedFaceMesh::numVertices).def("numFaces", &IndexedFaceMesh::numFaces).def("numEdges", &IndexedFaceMesh::numEdges)
        .def("getFaces", [](const IndexedFaceMesh &mm) { return py::array_t<unsigned>((py::ssize_t)mm.getFaces().size(), mm.getFaces().data()); })
        .def("getEdges", [](const IndexedFaceMesh &mm) {
            py::array_t<unsigned> r({(py::ssize_t)mm.getEdges().size(), (py::ssize_t)2});
            for (size_t i = 0; i < mm.getEdges().size(); i++) { r.mutable_at(i, 0) = mm.getEdges()[i].m_vert[0]; r.mutable_at(i, 1) = mm.getEdges()[i].m_vert[1]; }
            return r; });
    py::class_<IndexedTetMesh>(m, "IndexedTetMesh")
        .def("numVertices", &IndexedTetMesh::numVertices).def("numTets", &IndexedTetMesh::numTets).def("numEdges", &IndexedTetMesh::numEdges)
        .def("getTets", [](const IndexedTetMesh &mm) { return py::array_t<unsigned>((py::ssize_t)mm.getTets().size(), mm.getTets().data()); });
    py::class_<TriangleModel>(m, "TriangleModel")
        .def("getIndexOffset", &TriangleModel::getIndexOffset)
        .def("getParticleMesh", [](TriangleModel &t) -> IndexedFaceMesh & { return t.getParticleMesh(); }, py::return_value_policy::reference_internal)
        .def("getRestitutionCoeff", &TriangleModel::getRestitutionCoeff).def("setRestitutionCoeff", &TriangleModel::setRestitutionCoeff)
        .def("getFrictionCoeff", &TriangleModel::getFrictionCoeff).def("setFrictionCoeff", &TriangleModel::setFrictionCoeff)
        .def("updateMeshNormals", [](TriangleModel &, const ParticleData &) {});  // rendering helper of the reference: no normals on this path
    py::class_<TetModel>(m, "TetModel")
        .def("getIndexOffset", &TetModel::getIndexOffset)
        .def("getParticleMesh", [](TetModel &t) -> IndexedTetMesh & { return t.getParticleMesh(); }, py::return_value_policy::reference_internal)
        .def("getRestitutionCoeff", &TetModel::getRestitutionCoeff).def("setRestitutionCoeff", &TetModel::setRestitutionCoeff)
        .def("getFrictionCoeff", &TetModel::getFrictionCoeff).def("setFrictionCoeff", &TetModel::setFrictionCoeff)
        .def("updateMeshNormals", [](TetModel &, const ParticleData &) {});
    // RigidBody (Simulation/RigidBody.h): the state the path uses; bodies are created with SimulationModel.addRigidBody below
    py::class_<RigidBody>(m, "RigidBody")
        .def("getMass", &RigidBody::getMass)
        .def("getPosition", [](RigidBody &b) { return to_np(b.getPosition()); })
        .def("getVelocity", [](RigidBody &b) { return to_np(b.getVelocity()); })
        .def("getAngularVelocity", [](RigidBody &b) { return to_np(b.getAngularVelocity()); })
        .def("getRotation", [](RigidBody &b) { const Quaternionr &q = b.getRotation(); return py::make_tuple(q.w, q.x, q.y, q.z); })
        .def("getRestitutionCoeff", &RigidBody::getRestitutionCoeff).def("setRestitutionCoeff", &RigidBody::setRestitutionCoeff)
        .def("getFrictionCoeff", &RigidBody::getFrictionCoeff).def("setFrictionCoeff", &RigidBody::setFrictionCoeff);

    py::class_<SimulationModel>(m, "SimulationModel")
        .def(py::init<>())
        .def("init", &SimulationModel::init).def("reset", &SimulationModel::reset).def("cleanup", &SimulationModel::cleanup)
        .def("getParticles", &SimulationModel::getParticles, py::return_value_policy::reference_internal)
        .def("getTriangleModels", [](SimulationModel &s) { return s.getTriangleModels(); }, py::return_value_policy::reference_internal)
        .def("getTetModels", [](SimulationModel &s) { return s.getTetModels(); }, py::return_value_policy::reference_internal)
        .def("addRegularTriangleModel", [](SimulationModel &s, int w, int h, const py::object &t, const py::object &r, const py::object &sc, bool testMesh) {
                if (testMesh) throw std::runtime_error("testMesh=True needs the reference's collision detection, which is outside this engine's path");
                auto sa = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(sc);
                if (!sa || sa.size() != 2) throw std::invalid_argument("scale: expected 2 numbers");
                s.addRegularTriangleModel(w, h, vec3(t), mat3(r), Vector2r{{(Real)sa.data()[0], (Real)sa.data()[1]}});
                return s.getTriangleModels().back(); },
             py::arg("width"), py::arg("height"), py::arg("translation") = py::make_tuple(0, 0, 0),
             py::arg("rotation") = py::make_tuple(py::make_tuple(1, 0, 0), py::make_tuple(0, 1, 0), py::make_tuple(0, 0, 1)),
             py::arg("scale") = py::make_tuple(1, 1), py::arg("testMesh") = false, py::return_value_policy::reference_internal)
        .def("addRegularTetModel", [](SimulationModel &s, int w, int h, int d, const py::object &t, const py::object &r, const py::object &sc, bool testMesh) {
                if (testMesh) throw std::runtime_error("testMesh=True needs the reference's collision detection, which is outside this engine's path");
                s.addRegularTetModel(w, h, d, vec3(t), mat3(r), vec3(sc));
                return s.getTetModels().back(); },
             py::arg("width"), py::arg("height"), py::arg("depth"), py::arg("translation") = py::make_tuple(0, 0, 0),
             py::arg("rotation") = py::make_tuple(py::make_tuple(1, 0, 0), py::make_tuple(0, 1, 0), py::make_tuple(0, 0, 1)),
             py::arg("scale") = py::make_tuple(1, 1, 1), py::arg("testMesh") = false, py::return_value_policy::reference_internal)
        .def("addTriangleModel", [](SimulationModel &s, py::array_t<float, py::array::c_style | py::array::forcecast> pts,
                                    py::array_t<unsigned, py::array::c_style | py::array::forcecast> idx) {
                s.addTriangleModel((unsigned)(pts.size() / 3), (unsigned)(idx.size() / 3), reinterpret_cast<const Vector3r *>(pts.data()), idx.data());
                return s.getTriangleModels().back(); }, py::return_value_policy::reference_internal)
        .def("addTetModel", [](SimulationModel &s, py::array_t<float, py::array::c_style | py::array::forcecast> pts,
                               py::array_t<unsigned, py::array::c_style | py::array::forcecast> idx) {
                s.addTetModel((unsigned)(pts.size() / 3), (unsigned)(idx.size() / 4), reinterpret_cast<const Vector3r *>(pts.data()), idx.data());
                return s.getTetModels().back(); }, py::return_value_policy::reference_internal)
        .def("addClothConstraints", &SimulationModel::addClothConstraints)
        .def("addBendingConstraints", &SimulationModel::addBendingConstraints)
        .def("addSolidConstraints", &SimulationModel::addSolidConstraints)
        .def("addDistanceConstraint", &SimulationModel::addDistanceConstraint)
        .def("addDistanceConstraint_XPBD", &SimulationModel::addDistanceConstraint_XPBD)
        .def("addDihedralConstraint", &SimulationModel::addDihedralConstraint)
        .def("addIsometricBendingConstraint", &SimulationModel::addIsometricBendingConstraint)
        .def("addIsometricBendingConstraint_XPBD", &SimulationModel::addIsometricBendingConstraint_XPBD)
        .def("addFEMTriangleConstraint", &SimulationModel::addFEMTriangleConstraint)
        .def("addStrainTriangleConstraint", &SimulationModel::addStrainTriangleConstraint)
        .def("addVolumeConstraint", &SimulationModel::addVolumeConstraint)
        .def("addVolumeConstraint_XPBD", &SimulationModel::addVolumeConstraint_XPBD)
        .def("addFEMTetConstraint", &SimulationModel::addFEMTetConstraint)
        .def("addFEMTetConstraint_XPBD", &SimulationModel::addFEMTetConstraint_XPBD)
        .def("addStrainTetConstraint", &SimulationModel::addStrainTetConstraint)
        // RigidBody::initBody(mass, x, inertiaTensor, rotation, ...) without the mesh arguments (the ctypes facade pypbd.py carries the
        // density / mesh form with its volume integration); returns the body index
        .def("addRigidBody", [](SimulationModel &s, Real mass, const py::object &x, const py::object &inertia, const py::object &q) {
                RigidBody *rb = new RigidBody();
                auto qa = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(q);
                if (!qa || qa.size() != 4) throw std::invalid_argument("rotation: expected (w, x, y, z)");
                rb->initBody(mass, vec3(x), vec3(inertia), Quaternionr((Real)qa.data()[0], (Real)qa.data()[1], (Real)qa.data()[2], (Real)qa.data()[3]));
                s.getRigidBodies().push_back(rb);
                s.m_groupsInitialized = false; s.rigidBodiesDirty = true;
                return (unsigned)s.getRigidBodies().size() - 1; },
             py::arg("mass"), py::arg("x"), py::arg("inertiaTensor"), py::arg("rotation") = py::make_tuple(1.0, 0.0, 0.0, 0.0))
        .def("getRigidBodies", [](SimulationModel &s) { py::list l; for (RigidBody *b : s.getRigidBodies()) l.append(py::cast(b, py::return_value_policy::reference)); return l; })
        .def("getContactStiffnessParticleRigidBody", &SimulationModel::getContactStiffnessParticleRigidBody)
        .def("setContactStiffnessParticleRigidBody", &SimulationModel::setContactStiffnessParticleRigidBody)
        .def("addBallJoint", [](SimulationModel &s, unsigned a, unsigned b, const py::object &p) { return s.addBallJoint(a, b, vec3(p)); })
        .def("addRigidBodyParticleBallJoint", &SimulationModel::addRigidBodyParticleBallJoint)
        .def("initConstraintGroups", &SimulationModel::initConstraintGroups)
        .def("getConstraintGroups", [](SimulationModel &s) { if (!s.m_groupsInitialized) s.initConstraintGroups(); return s.getConstraintGroups(); })
        .def("numConstraints", &SimulationModel::numConstraints)
        .def("getConstraints", [](SimulationModel &s) {  // type name, bodies, parameters (flat layout of include/pbd_b200.h)
                py::list out;
                for (unsigned i = 0; i < s.numConstraints(); i++) {
                    const ConstraintView c = s.getConstraint(i);
                    py::dict d;
                    d["type"] = c.type;
                    d["bodies"] = py::array_t<unsigned>((py::ssize_t)c.numberOfBodies, c.m_bodies);
                    d["params"] = py::array_t<float>((py::ssize_t)c.numParams, c.params);
                    out.append(d);
                }
                return out; })
        .def("setClothStiffness", &SimulationModel::setClothStiffness).def("setClothStiffnessXX", &SimulationModel::setClothStiffnessXX)
        .def("setClothStiffnessYY", &SimulationModel::setClothStiffnessYY).def("setClothStiffnessXY", &SimulationModel::setClothStiffnessXY)
        .def("setClothPoissonRatioXY", &SimulationModel::setClothPoissonRatioXY).def("setClothPoissonRatioYX", &SimulationModel::setClothPoissonRatioYX)
        .def("setClothBendingStiffness", &SimulationModel::setClothBendingStiffness)
        .def("setSolidStiffness", &SimulationModel::setSolidStiffness).def("setSolidPoissonRatio", &SimulationModel::setSolidPoissonRatio)
        .def("setSolidVolumeStiffness", &SimulationModel::setSolidVolumeStiffness);

    // CollisionDetection / DistanceFieldCollisionDetection (pyPBD/CollisionDetectionModule.cpp): the add* calls of the reference; `vertices` is an
    // (n, 3) array or None
    auto verts = [](const py::object &o, std::vector<Vector3r> &v) {
        v.clear();
        if (o.is_none()) return;
        auto a = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(o);
        if (!a || a.size() % 3) throw std::invalid_argument("vertices: expected an (n, 3) array");
        for (ssize_t i = 0; i < a.size() / 3; i++) v.emplace_back((Real)a.data()[3 * i], (Real)a.data()[3 * i + 1], (Real)a.data()[3 * i + 2]);
    };
    auto vec2 = [](const py::object &o) {
        auto a = py::array_t<double, py::array::c_style | py::array::forcecast>::ensure(o);
        if (!a || a.size() != 2) throw std::invalid_argument("expected 2 numbers");
        return Vector2r{{(Real)a.data()[0], (Real)a.data()[1]}};
    };
    py::class_<CollisionDetection> cdb(m, "CollisionDetection");
    cdb.def("getTolerance", &CollisionDetection::getTolerance).def("setTolerance", &CollisionDetection::setTolerance)
       .def("numCollisionObjects", [](CollisionDetection &c) { return c.getCollisionObjects().size(); })
       .def("cleanup", &CollisionDetection::cleanup);
    py::class_<CollisionDetection::CollisionObject> co(cdb, "CollisionObject");
    co.attr("RigidBodyCollisionObjectType") = (unsigned)CollisionDetection::CollisionObject::RigidBodyCollisionObjectType;
    co.attr("TriangleModelCollisionObjectType") = (unsigned)CollisionDetection::CollisionObject::TriangleModelCollisionObjectType;
    co.attr("TetModelCollisionObjectType") = (unsigned)CollisionDetection::CollisionObject::TetModelCollisionObjectType;
    py::class_<DistanceFieldCollisionDetection, CollisionDetection>(m, "DistanceFieldCollisionDetection")
        .def(py::init<>())
        .def("init", [](DistanceFieldCollisionDetection &) {})
        .def("addCollisionBox", [verts](DistanceFieldCollisionDetection &c, unsigned bi, unsigned bt, const py::object &v, unsigned, const py::object &box, bool tm, bool inv) {
                std::vector<Vector3r> vv; verts(v, vv); c.addCollisionBox(bi, bt, vv.data(), (unsigned)vv.size(), vec3(box), tm, inv); },
             py::arg("bodyIndex"), py::arg("bodyType"), py::arg("vertices"), py::arg("numVertices"), py::arg("box"), py::arg("testMesh") = true, py::arg("invertSDF") = false)
        .def("addCollisionSphere", [verts](DistanceFieldCollisionDetection &c, unsigned bi, unsigned bt, const py::object &v, unsigned, Real r, bool tm, bool inv) {
                std::vector<Vector3r> vv; verts(v, vv); c.addCollisionSphere(bi, bt, vv.data(), (unsigned)vv.size(), r, tm, inv); },
             py::arg("bodyIndex"), py::arg("bodyType"), py::arg("vertices"), py::arg("numVertices"), py::arg("radius"), py::arg("testMesh") = true, py::arg("invertSDF") = false)
        .def("addCollisionTorus", [verts, vec2](DistanceFieldCollisionDetection &c, unsigned bi, unsigned bt, const py::object &v, unsigned, const py::object &radii, bool tm, bool inv) {
                std::vector<Vector3r> vv; verts(v, vv); c.addCollisionTorus(bi, bt, vv.data(), (unsigned)vv.size(), vec2(radii), tm, inv); },
             py::arg("bodyIndex"), py::arg("bodyType"), py::arg("vertices"), py::arg("numVertices"), py::arg("radii"), py::arg("testMesh") = true, py::arg("invertSDF") = false)
        .def("addCollisionCylinder", [verts, vec2](DistanceFieldCollisionDetection &c, unsigned bi, unsigned bt, const py::object &v, unsigned, const py::object &dim, bool tm, bool inv) {
                std::vector<Vector3r> vv; verts(v, vv); c.addCollisionCylinder(bi, bt, vv.data(), (unsigned)vv.size(), vec2(dim), tm, inv); },
             py::arg("bodyIndex"), py::arg("bodyType"), py::arg("vertices"), py::arg("numVertices"), py::arg("dim"), py::arg("testMesh") = true, py::arg("invertSDF") = false)
        .def("addCollisionHollowSphere", [verts](DistanceFieldCollisionDetection &c, unsigned bi, unsigned bt, const py::object &v, unsigned, Real r, Real th, bool tm, bool inv) {
                std::vector<Vector3r> vv; verts(v, vv); c.addCollisionHollowSphere(bi, bt, vv.data(), (unsigned)vv.size(), r, th, tm, inv); },
             py::arg("bodyIndex"), py::arg("bodyType"), py::arg("vertices"), py::arg("numVertices"), py::arg("radius"), py::arg("thickness"), py::arg("testMesh") = true, py::arg("invertSDF") = false)
        .def("addCollisionHollowBox", [verts](DistanceFieldCollisionDetection &c, unsigned bi, unsigned bt, const py::object &v, unsigned, const py::object &box, Real th, bool tm, bool inv) {
                std::vector<Vector3r> vv; verts(v, vv); c.addCollisionHollowBox(bi, bt, vv.data(), (unsigned)vv.size(), vec3(box), th, tm, inv); },
             py::arg("bodyIndex"), py::arg("bodyType"), py::arg("vertices"), py::arg("numVertices"), py::arg("box"), py::arg("thickness"), py::arg("testMesh") = true, py::arg("invertSDF") = false)
        .def("addCollisionObjectWithoutGeometry", [](DistanceFieldCollisionDetection &c, unsigned bi, unsigned bt, const py::object &, unsigned, bool tm) {
                c.addCollisionObjectWithoutGeometry(bi, bt, nullptr, 0, tm); },
             py::arg("bodyIndex"), py::arg("bodyType"), py::arg("vertices"), py::arg("numVertices"), py::arg("testMesh"));

    py::class_<TimeStepController> ts(m, "TimeStepController");
    ts.def(py::init([](int device) {
                auto t = std::unique_ptr<TimeStepController>(new TimeStepController(device));
                if (!t->valid()) throw std::runtime_error("pypbd_b200: " + t->error() + " (no CPU fallback)");
                return t; }), py::arg("device") = 0)
      .def("init", &TimeStepController::init).def("reset", &TimeStepController::reset)
        .def("setValueUInt", [](TimeStepController &t, int id, unsigned v) { if (!t.setValueUInt(id, v)) throw std::invalid_argument("TimeStepController.setValueUInt: bad parameter id or value"); })
        .def("getValueUInt", &TimeStepController::getValueUInt)
        .def("setValueInt", [](TimeStepController &t, int id, int v) { if (!t.setValueInt(id, v)) throw std::invalid_argument("TimeStepController.setValueInt: bad parameter id or value"); })
        .def("getValueInt", &TimeStepController::getValueInt)
        .def("setSolverMode", &TimeStepController::setSolverMode)
        .def("setCollisionDetection", [](TimeStepController &t, SimulationModel &model, CollisionDetection *cd) { t.setCollisionDetection(model, cd); },
             py::keep_alive<1, 3>())  // TimeStep::setCollisionDetection: the time step refers to the collision detection
        .def("getCollisionDetection", &TimeStepController::getCollisionDetection, py::return_value_policy::reference)
        .def("step", [](TimeStepController &t, SimulationModel &model) {  // TimeStepModule.cpp:15-31: the GIL is held for the whole step there too
                if (!t.step(model)) throw std::runtime_error("pypbd_b200: " + t.error()); });
    ts.attr("NUM_SUB_STEPS") = (int)TimeStepController::NUM_SUB_STEPS;
    ts.attr("MAX_ITERATIONS") = (int)TimeStepController::MAX_ITERATIONS;
    ts.attr("MAX_ITERATIONS_V") = (int)TimeStepController::MAX_ITERATIONS_V;
    ts.attr("VELOCITY_UPDATE_METHOD") = (int)TimeStepController::VELOCITY_UPDATE_METHOD;
    ts.attr("ENUM_VUPDATE_FIRST_ORDER") = (int)TimeStepController::ENUM_VUPDATE_FIRST_ORDER;
    ts.attr("ENUM_VUPDATE_SECOND_ORDER") = (int)TimeStepController::ENUM_VUPDATE_SECOND_ORDER;

    py::class_<TimeManagerRef>(m, "TimeManager")
        .def_static("getCurrent", []() { return TimeManagerRef(); })
        .def("getTime", [](TimeManagerRef &) { return Sim::current().timeStep()->timeManager().getTime(); })
        .def("setTime", [](TimeManagerRef &, Real t) { Sim::current().timeStep()->timeManager().setTime(t); })
        .def("getTimeStepSize", [](TimeManagerRef &) { return Sim::current().timeStep()->timeManager().getTimeStepSize(); })
        .def("setTimeStepSize", [](TimeManagerRef &, Real h) { Sim::current().timeStep()->timeManager().setTimeStepSize(h); });

    py::class_<Sim> sim(m, "Simulation");
    sim.def_static("getCurrent", []() -> Sim & { return Sim::current(); }, py::return_value_policy::reference)
        .def_static("hasCurrent", []() { return true; })
        .def("initDefault", [](Sim &s, int device) { s.model.reset(new SimulationModel()); s.model->init(); s.ts.reset(); s.device = device; }, py::arg("device") = 0)
        .def("getModel", [](Sim &s) -> SimulationModel * { return s.model.get(); }, py::return_value_policy::reference_internal)
        .def("getTimeStep", [](Sim &s) -> TimeStepController * { return s.timeStep(); }, py::return_value_policy::reference_internal)
        .def("setVecValueReal", [](Sim &s, int, const py::object &v) { s.gravity = vec3(v); if (s.ts) s.ts->setGravitation(s.gravity); })
        .def("getVecValueReal", [](Sim &s, int) { return to_np(s.gravity); })
        .def("reset", [](Sim &s) { if (s.model) s.model->reset(); if (s.ts) { s.ts->reset(); s.ts->timeManager().setTime(0); } });
    sim.attr("GRAVITATION") = 0;

    // pyPBD's Logger / Timing entry points the example scripts call (UtilitiesModule.cpp); nothing of its own to log on this path
    struct LoggerT {}; struct TimingT {};
    py::class_<LoggerT>(m, "Logger")
        .def_static("addConsoleSink", [](int) {}, py::arg("level") = 1)
        .def_static("addFileSink", [](int, const std::string &) {}, py::arg("level") = 1, py::arg("path") = std::string());
    py::class_<TimingT>(m, "Timing")
        .def_static("printAverageTimes", []() {}).def_static("printTimeSums", []() {}).def_static("reset", []() {});
    py::module_ ll = m.def_submodule("LogLevel");
    ll.attr("DEBUG") = 0; ll.attr("INFO") = 1; ll.attr("WARN") = 2; ll.attr("ERR") = 3;
}
