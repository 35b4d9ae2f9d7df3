// integration/GpuTimeStepController.h
//
// The reference-side binding of libpbd_b200.so: a PBD::TimeStep subclass (Simulation/TimeStep.h:13-48) that a maintainer adds to
// the reference tree and installs with
//     auto *ts = new PBD::GpuTimeStepController(); ts->init(); delete sim->getTimeStep(); sim->setTimeStep(ts);
// exactly as Demos/PositionBasedElasticRodsDemo/PositionBasedElasticRodsDemo.cpp:51-54 installs its own time step.  It replaces
// the body of TimeStepController::step (Simulation/TimeStepController.cpp:75-241) for the particle constraints and the two
// coupling joints; everything else of the reference (scene construction, colouring, GUI, IO) stays as it is.
//
// This header includes the reference's own headers and is compiled INSIDE the reference (it is part of neither
// libpbd_b200.so nor the Python package).  In this repository it is compiled by oracle/Makefile (target `refgpu`) against the
// reference sources under /root/reference and exercised by tests/test_gpu_parity.py::test_reference_side_adapter_*: the
// reference builds the scene and colours it, this class steps it on the GPU, the reference's own TimeStepController steps a
// twin on the CPU, and the two trajectories are compared.
//
// Exactly one translation unit defines PBD_GPU_TIMESTEP_IMPLEMENTATION before including it (static parameter ids).
#pragma once
#include "Simulation/TimeStep.h"
#include "Simulation/SimulationModel.h"
#include "Simulation/Constraints.h"
#include "Simulation/RigidBody.h"
#include "Simulation/TimeManager.h"
#include "Simulation/Simulation.h"
#include "Simulation/DistanceFieldCollisionDetection.h"
#include "pbd_b200.h"
#include <cstring>
#include <string>
#include <vector>

namespace PBD {

class GpuTimeStepController : public TimeStep {
public:
    // same parameter names and meaning as TimeStepController.h:16-22 / TimeStepController.cpp:38-73
    static int NUM_SUB_STEPS, MAX_ITERATIONS, MAX_ITERATIONS_V, VELOCITY_UPDATE_METHOD;
    static int ENUM_VUPDATE_FIRST_ORDER, ENUM_VUPDATE_SECOND_ORDER;

    explicit GpuTimeStepController(int device = 0) {
        m_collisionDetection = NULL;
        if (pbd_create(device, nullptr, &m_engine)) { m_error = pbd_last_error(); m_engine = nullptr; }
    }
    ~GpuTimeStepController() override { unpin(); pbd_destroy(m_engine); }

    const std::string &lastError() const { return m_error; }
    bool ok() const { return m_engine != nullptr && m_error.empty(); }
    void setMode(int mode) { if (m_engine) pbd_set_mode(m_engine, mode); }

    void reset() override { TimeStep::reset(); m_boundConstraints = ~size_t(0); }

    // Force a re-flatten at the next step.  The adapter notices on its own: added / removed constraints, particles and bodies
    // (m_groupsInitialized, sizes), changed masses (compared every step) and the model-wide parameter setters (setClothStiffness,
    // setSolidStiffness, ... rewrite every constraint of a type: sentinel constraints are compared every step).  Call this after
    // editing members of individual constraints or rigid bodies by hand.
    void invalidate() { m_boundConstraints = ~size_t(0); }
    // Host-state policy.  Default (true): the model's positions and velocities are uploaded before every step, so edits made by user
    // code between steps (dragging particles, resetting velocities) are honoured exactly like the reference does -- at the price of
    // 24 bytes per particle of PCIe traffic per step.  false: the device state is authoritative between steps; positions and
    // velocities are still downloaded into the model after every step (rendering keeps working), uploads happen only at bind time and
    // after invalidateState().  For pure simulation loops this removes two thirds of the per-step copies.
    void setHostStateAuthoritative(bool on) { m_hostAuthoritative = on; m_stateInvalid = true; }
    void invalidateState() { m_stateInvalid = true; }
    // The device owns oldX / lastX while this time step is installed (they feed the second-order velocity update); copy them back
    // into the model before handing it to another TimeStep (TimeStepController.cpp:112-118 reads them).
    bool downloadHistory(SimulationModel &model) {
        ParticleData &pd = model.getParticles();
        const unsigned n = pd.size();
        if (!m_engine || n != m_boundParticles || n == 0) return false;
        std::vector<float> t(3 * (size_t)n);
        if (pbd_get_attr(m_engine, PBD_ATTR_OLDX, t.data())) { fail(); return false; }
        for (unsigned i = 0; i < n; i++) pd.setOldPosition(i, Vector3r((Real)t[3 * i], (Real)t[3 * i + 1], (Real)t[3 * i + 2]));
        if (pbd_get_attr(m_engine, PBD_ATTR_LASTX, t.data())) { fail(); return false; }
        for (unsigned i = 0; i < n; i++) pd.setLastPosition(i, Vector3r((Real)t[3 * i], (Real)t[3 * i + 1], (Real)t[3 * i + 2]));
        return true;
    }

    // TimeStepController::step for the engine's path.  Host state in, host state out: the model's ParticleData / RigidBody
    // objects hold the result when this returns, so rendering and user code keep working unchanged.
    void step(SimulationModel &model) override {
        if (!m_engine) return;
        // The contact path (TimeStepController.cpp:189-196: collision detection, then velocityConstraintProjection over the contact
        // constraints, :298-357).  On the GPU: particles of triangle / tet models against analytic distance fields on static rigid
        // bodies (collectColliders below); anything else -- dynamic bodies in contact, rigid-rigid or particle-tet contacts, another
        // CollisionDetection class, hand-made contact constraints -- is refused instead of being simulated without its contacts.
        const bool haveObjects = m_collisionDetection && !m_collisionDetection->getCollisionObjects().empty();
        if (!haveObjects && (!model.getRigidBodyContactConstraints().empty() || !model.getParticleRigidBodyContactConstraints().empty() ||
                             !model.getParticleSolidContactConstraints().empty())) {
            m_error = "GpuTimeStepController: the model holds contact constraints but no collision detection that produces them; the "
                      "contact path of such a model (TimeStepController.cpp:298-357) runs on the CPU TimeStepController only";
            return;
        }
        std::vector<pbd_particle_collider> pcs;
        std::vector<pbd_rigid_collider> rcs;
        if (haveObjects && !collectColliders(model, pcs, rcs)) return;  // m_error says why
        m_error.clear();  // lastError() describes the current step
        ParticleData &pd = model.getParticles();
        SimulationModel::RigidBodyVector &rbs = model.getRigidBodies();
        const unsigned n = pd.size();
        // (re)bind when the model changed: any add*Constraint clears m_groupsInitialized (SimulationModel.cpp); parameter setters
        // and hand edits do not (SimulationModel.h:266), hence the sentinel comparison
        if (!model.m_groupsInitialized || model.getConstraints().size() != m_boundConstraints || n != m_boundParticles ||
            rbs.size() != m_boundBodies || sentinelSignature(model) != m_signature) {
            if (!bind(model)) return;
        } else if (m_checkMasses && massesChanged(pd, n)) {
            if (pbd_set_masses(m_engine, m_mass.data())) return fail();  // ParticleData::setMass after the first step
        }
        TimeManager *tm = TimeManager::getCurrent();
        const Vector3r g(Simulation::getCurrent()->getVecValue<Real>(Simulation::GRAVITATION));
        const float grav[3] = {(float)g[0], (float)g[1], (float)g[2]};
        if (pbd_set_params(m_engine, (float)tm->getTimeStepSize(), m_subSteps, m_maxIterations, m_velocityUpdateMethod, grav)) return fail();
        if (haveObjects || m_hadColliders) {
            if (!sameColliders(pcs, rcs)) {
                if (pbd_set_colliders(m_engine, (unsigned)pcs.size(), pcs.data(), (unsigned)rcs.size(), rcs.data())) return fail();
                m_pcs = pcs; m_rcs = rcs;
            }
            m_hadColliders = haveObjects;
            if (haveObjects) {
                if (pbd_set_contact_params(m_engine, (float)m_collisionDetection->getTolerance(), (float)model.getContactStiffnessParticleRigidBody(), m_maxIterationsV)) return fail();
                model.resetContacts();  // DistanceFieldCollisionDetection::collisionDetection starts with this; the contacts of the step live on the device
            }
        }
        // rigid bodies: uploaded at bind time, device state is authoritative afterwards (their history feeds the second-order
        // velocity update); particles: x and v come from the host every step, so user edits between steps are honoured
        if (sizeof(Real) == sizeof(float) && n) {
            // Real == float: std::vector<Vector3r> is already n x 3 packed floats (Common/Common.h:31); the engine copies straight
            // from and into the model's own arrays (page-locked at bind time), no host-side conversion at all
            float *x = reinterpret_cast<float *>(&pd.getPosition(0)[0]), *v = reinterpret_cast<float *>(&pd.getVelocity(0)[0]);
            if (!m_pinFailed && (x != m_pinnedX || v != m_pinnedV)) pin(x, v, n);  // the vectors were (re)allocated
            const bool up = m_hostAuthoritative || m_stateInvalid;
            if (pbd_step_host(m_engine, 1, up ? x : nullptr, up ? v : nullptr, x, v)) return fail();
            m_stateInvalid = false;
        } else {
            const bool up = m_hostAuthoritative || m_stateInvalid;
            if (up) packParticles(pd, n); else { m_x.resize(3 * (size_t)n); m_v.resize(3 * (size_t)n); }
            if (pbd_step_host(m_engine, 1, (n && up) ? m_x.data() : nullptr, (n && up) ? m_v.data() : nullptr, n ? m_x.data() : nullptr, n ? m_v.data() : nullptr)) return fail();
            m_stateInvalid = false;
#pragma omp parallel for schedule(static)
            for (int i = 0; i < (int)n; i++) {
                pd.getPosition(i) = Vector3r((Real)m_x[3 * i], (Real)m_x[3 * i + 1], (Real)m_x[3 * i + 2]);
                pd.getVelocity(i) = Vector3r((Real)m_v[3 * i], (Real)m_v[3 * i + 1], (Real)m_v[3 * i + 2]);
            }
        }
        if (!rbs.empty() && !downloadBodies(rbs)) return;
        tm->setTime(tm->getTime() + tm->getTimeStepSize());  // TimeStepController.cpp:239
    }

protected:
    unsigned int m_subSteps = 5, m_maxIterations = 1, m_maxIterationsV = 5;
    int m_velocityUpdateMethod = 0;
    pbd_engine *m_engine = nullptr;
    size_t m_boundConstraints = ~size_t(0);
    unsigned m_boundParticles = ~0u;
    size_t m_boundBodies = ~size_t(0);
    std::string m_error;
    std::vector<float> m_x, m_v, m_mass;
    float *m_pinnedX = nullptr, *m_pinnedV = nullptr;
    bool m_pinFailed = false, m_checkMasses = true, m_hostAuthoritative = true, m_stateInvalid = true;
    std::vector<float> m_signature;
    std::vector<pbd_particle_collider> m_pcs;
    std::vector<pbd_rigid_collider> m_rcs;
    bool m_hadColliders = false;

    bool sameColliders(const std::vector<pbd_particle_collider> &p, const std::vector<pbd_rigid_collider> &r) const {
        return p.size() == m_pcs.size() && r.size() == m_rcs.size() &&
               (p.empty() || std::memcmp(p.data(), m_pcs.data(), p.size() * sizeof(p[0])) == 0) &&
               (r.empty() || std::memcmp(r.data(), m_rcs.data(), r.size() * sizeof(r[0])) == 0);
    }
    // The collision objects of the attached DistanceFieldCollisionDetection in the layout of pbd_set_colliders, following the pair
    // dispatch of DistanceFieldCollisionDetection::collisionDetection (DistanceFieldCollisionDetection.cpp:96-165): a pair produces
    // contacts when co1 has m_testMesh and co2 is a rigid body or a tet model, both being distance-field objects.
    bool collectColliders(SimulationModel &model, std::vector<pbd_particle_collider> &pcs, std::vector<pbd_rigid_collider> &rcs) {
        typedef DistanceFieldCollisionDetection D;
        D *cd = dynamic_cast<D *>(m_collisionDetection);
        if (!cd) { m_error = "GpuTimeStepController: the attached collision detection is not a DistanceFieldCollisionDetection; its contact path runs on the CPU TimeStepController only"; return false; }
        SimulationModel::RigidBodyVector &rbs = model.getRigidBodies();
        unsigned tetObjects = 0;
        for (CollisionDetection::CollisionObject *co : cd->getCollisionObjects()) {
            if (!cd->isDistanceFieldCollisionObject(co)) continue;  // the reference skips every pair with such an object
            D::DistanceFieldCollisionObject *dco = static_cast<D::DistanceFieldCollisionObject *>(co);
            const int id = co->getTypeId();
            if (co->m_bodyType == CollisionDetection::CollisionObject::TriangleModelCollisionObjectType ||
                co->m_bodyType == CollisionDetection::CollisionObject::TetModelCollisionObjectType) {
                const bool tet = (co->m_bodyType == CollisionDetection::CollisionObject::TetModelCollisionObjectType);
                if (tet && ++tetObjects > 1) { m_error = "GpuTimeStepController: two tet models as collision objects produce particle-tet contacts (collisionDetectionSolidSolid), which run on the CPU TimeStepController only"; return false; }
                if (!dco->m_testMesh) continue;  // never a co1, and as co2 only a tet model matters (counted above)
                pbd_particle_collider pc;
                if (tet) { TetModel *tm = model.getTetModels()[co->m_bodyIndex]; pc.offset = tm->getIndexOffset(); pc.count = tm->getParticleMesh().numVertices();
                           pc.restitution = (float)tm->getRestitutionCoeff(); pc.friction = (float)tm->getFrictionCoeff(); }
                else { TriangleModel *tm = model.getTriangleModels()[co->m_bodyIndex]; pc.offset = tm->getIndexOffset(); pc.count = tm->getParticleMesh().numVertices();
                       pc.restitution = (float)tm->getRestitutionCoeff(); pc.friction = (float)tm->getFrictionCoeff(); }
                pcs.push_back(pc);
                continue;
            }
            if (co->m_bodyType != CollisionDetection::CollisionObject::RigidBodyCollisionObjectType) continue;
            pbd_rigid_collider rc;
            std::memset(&rc, 0, sizeof(rc));
            if (id == D::DistanceFieldCollisionBox::TYPE_ID) { rc.shape = PBD_SHAPE_BOX; auto *o = static_cast<D::DistanceFieldCollisionBox *>(co); for (int k = 0; k < 3; k++) rc.dim[k] = (float)o->m_box[k]; }
            else if (id == D::DistanceFieldCollisionSphere::TYPE_ID) { rc.shape = PBD_SHAPE_SPHERE; rc.dim[0] = (float)static_cast<D::DistanceFieldCollisionSphere *>(co)->m_radius; }
            else if (id == D::DistanceFieldCollisionTorus::TYPE_ID) { rc.shape = PBD_SHAPE_TORUS; auto *o = static_cast<D::DistanceFieldCollisionTorus *>(co); rc.dim[0] = (float)o->m_radii[0]; rc.dim[1] = (float)o->m_radii[1]; }
            else if (id == D::DistanceFieldCollisionCylinder::TYPE_ID) { rc.shape = PBD_SHAPE_CYLINDER; auto *o = static_cast<D::DistanceFieldCollisionCylinder *>(co); rc.dim[0] = (float)o->m_dim[0]; rc.dim[1] = (float)o->m_dim[1]; }
            else if (id == D::DistanceFieldCollisionHollowSphere::TYPE_ID) { rc.shape = PBD_SHAPE_HOLLOW_SPHERE; auto *o = static_cast<D::DistanceFieldCollisionHollowSphere *>(co); rc.dim[0] = (float)o->m_radius; rc.thickness = (float)o->m_thickness; }
            else if (id == D::DistanceFieldCollisionHollowBox::TYPE_ID) { rc.shape = PBD_SHAPE_HOLLOW_BOX; auto *o = static_cast<D::DistanceFieldCollisionHollowBox *>(co); for (int k = 0; k < 3; k++) rc.dim[k] = (float)o->m_box[k]; rc.thickness = (float)o->m_thickness; }
            else continue;  // a rigid body without geometry: never a co2 that produces contacts
            RigidBody *rb = rbs[co->m_bodyIndex];
            if (rb->getMass() != 0.0) { m_error = "GpuTimeStepController: rigid body " + std::to_string(co->m_bodyIndex) + " is a dynamic collision object; contacts with dynamic bodies (and rigid-rigid contacts) run on the CPU TimeStepController only"; return false; }
            if (dco->m_invertSDF < 0) { m_error = "GpuTimeStepController: collision object with an inverted distance field; its candidate pruning is specific to the reference's bounding-sphere hierarchy (CPU TimeStepController only)"; return false; }
            rc.body = co->m_bodyIndex; rc.invert_sdf = 0;
            rc.restitution = (float)rb->getRestitutionCoeff(); rc.friction = (float)rb->getFrictionCoeff();
            const Matrix3r &R = rb->getTransformationR();
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) rc.R[3 * r + c] = (float)R(r, c);
            for (int k = 0; k < 3; k++) { rc.v1[k] = (float)rb->getTransformationV1()[k]; rc.v2[k] = (float)rb->getTransformationV2()[k]; }
            cd->updateAABB(model, co);  // what collisionDetection does first (DistanceFieldCollisionDetection.cpp:63)
            for (int k = 0; k < 3; k++) { rc.aabb_min[k] = (float)co->m_aabb.m_p[0][k]; rc.aabb_max[k] = (float)co->m_aabb.m_p[1][k]; }
            rcs.push_back(rc);
        }
        return true;
    }

    // values of the first and the last constraint of the model plus one in the middle, per call: O(1)
    static void appendConstraintValues(Constraint *c, std::vector<float> &sig) {
        const int tid = c->getTypeId();
        sig.push_back((float)tid);
        if (tid == DistanceConstraint::TYPE_ID) sig.push_back((float)static_cast<DistanceConstraint *>(c)->m_stiffness);
        else if (tid == DistanceConstraint_XPBD::TYPE_ID) sig.push_back((float)static_cast<DistanceConstraint_XPBD *>(c)->m_stiffness);
        else if (tid == DihedralConstraint::TYPE_ID) sig.push_back((float)static_cast<DihedralConstraint *>(c)->m_stiffness);
        else if (tid == IsometricBendingConstraint::TYPE_ID) sig.push_back((float)static_cast<IsometricBendingConstraint *>(c)->m_stiffness);
        else if (tid == IsometricBendingConstraint_XPBD::TYPE_ID) sig.push_back((float)static_cast<IsometricBendingConstraint_XPBD *>(c)->m_stiffness);
        else if (tid == FEMTriangleConstraint::TYPE_ID) { auto *d = static_cast<FEMTriangleConstraint *>(c);
            for (Real v : {d->m_xxStiffness, d->m_yyStiffness, d->m_xyStiffness, d->m_xyPoissonRatio, d->m_yxPoissonRatio}) sig.push_back((float)v); }
        else if (tid == StrainTriangleConstraint::TYPE_ID) { auto *d = static_cast<StrainTriangleConstraint *>(c);
            for (Real v : {d->m_xxStiffness, d->m_yyStiffness, d->m_xyStiffness}) sig.push_back((float)v);
            sig.push_back(d->m_normalizeStretch ? 1.f : 0.f); sig.push_back(d->m_normalizeShear ? 1.f : 0.f); }
        else if (tid == VolumeConstraint::TYPE_ID) sig.push_back((float)static_cast<VolumeConstraint *>(c)->m_stiffness);
        else if (tid == VolumeConstraint_XPBD::TYPE_ID) sig.push_back((float)static_cast<VolumeConstraint_XPBD *>(c)->m_stiffness);
        else if (tid == FEMTetConstraint::TYPE_ID) { auto *d = static_cast<FEMTetConstraint *>(c); sig.push_back((float)d->m_stiffness); sig.push_back((float)d->m_poissonRatio); }
        else if (tid == XPBD_FEMTetConstraint::TYPE_ID) { auto *d = static_cast<XPBD_FEMTetConstraint *>(c); sig.push_back((float)d->m_stiffness); sig.push_back((float)d->m_poissonRatio); }
        else if (tid == StrainTetConstraint::TYPE_ID) { auto *d = static_cast<StrainTetConstraint *>(c); sig.push_back((float)d->m_stretchStiffness); sig.push_back((float)d->m_shearStiffness);
            sig.push_back(d->m_normalizeStretch ? 1.f : 0.f); sig.push_back(d->m_normalizeShear ? 1.f : 0.f); }
        else if (tid == ShapeMatchingConstraint::TYPE_ID) sig.push_back((float)static_cast<ShapeMatchingConstraint *>(c)->m_stiffness);
    }
    // The model-wide setters rewrite every constraint of a type, so the first and the last constraint of each run of equal type
    // ids (the add*Constraints calls append type by type) see them; O(#runs) per step.
    static std::vector<float> sentinelSignature(SimulationModel &model) {
        std::vector<float> sig;
        SimulationModel::ConstraintVector &cs = model.getConstraints();
        const size_t N = cs.size();
        if (N == 0) return sig;
        // sample positions: both ends, and a binary subdivision down to 64 samples: finds every run boundary cheaply enough and
        // does not depend on how the constraints were appended
        const size_t samples = N < 64 ? N : 64;
        for (size_t k = 0; k < samples; k++) appendConstraintValues(cs[(size_t)((double)k * (double)(N - 1) / (double)(samples > 1 ? samples - 1 : 1))], sig);
        return sig;
    }
    bool massesChanged(ParticleData &pd, unsigned n) {
        if (m_mass.size() != n) { m_mass.resize(n); for (unsigned i = 0; i < n; i++) m_mass[i] = (float)pd.getMass(i); return true; }
        int changed = 0;
#pragma omp parallel for schedule(static) reduction(| : changed)
        for (int i = 0; i < (int)n; i++) {
            const float m = (float)pd.getMass(i);
            if (m != m_mass[i]) { m_mass[i] = m; changed |= 1; }
        }
        return changed != 0;
    }

    void fail() { m_error = pbd_last_error(); }
    void unpin() {
        if (m_pinnedX) pbd_unpin_host(m_pinnedX);
        if (m_pinnedV) pbd_unpin_host(m_pinnedV);
        m_pinnedX = m_pinnedV = nullptr;
    }
    void pin(float *x, float *v, unsigned n) {  // best effort: an unpinned array still works, only slower
        unpin();
        if (m_pinFailed) return;  // do not retry every step
        if (pbd_pin_host(x, (size_t)n * 3 * sizeof(float)) == 0) m_pinnedX = x; else m_pinFailed = true;
        if (pbd_pin_host(v, (size_t)n * 3 * sizeof(float)) == 0) m_pinnedV = v; else m_pinFailed = true;
    }

    void initParameters() override {
        TimeStep::initParameters();
        NUM_SUB_STEPS = createNumericParameter("subSteps", "# sub steps", &m_subSteps);
        setGroup(NUM_SUB_STEPS, "Simulation|PBD");
        setDescription(NUM_SUB_STEPS, "Number of sub steps of the solver.");
        static_cast<GenParam::NumericParameter<unsigned int> *>(getParameter(NUM_SUB_STEPS))->setMinValue(1);
        MAX_ITERATIONS = createNumericParameter("maxIterations", "Max. iterations", &m_maxIterations);
        setGroup(MAX_ITERATIONS, "Simulation|PBD");
        setDescription(MAX_ITERATIONS, "Maximal number of iterations of the solver.");
        static_cast<GenParam::NumericParameter<unsigned int> *>(getParameter(MAX_ITERATIONS))->setMinValue(1);
        MAX_ITERATIONS_V = createNumericParameter("maxIterationsV", "Max. velocity iterations", &m_maxIterationsV);
        setGroup(MAX_ITERATIONS_V, "Simulation|PBD");
        setDescription(MAX_ITERATIONS_V, "Maximal number of iterations of the velocity solver (particle / static-body contacts).");
        VELOCITY_UPDATE_METHOD = createEnumParameter("velocityUpdateMethod", "Velocity update method", &m_velocityUpdateMethod);
        setGroup(VELOCITY_UPDATE_METHOD, "Simulation|PBD");
        setDescription(VELOCITY_UPDATE_METHOD, "Velocity method.");
        GenParam::EnumParameter *ep = static_cast<GenParam::EnumParameter *>(getParameter(VELOCITY_UPDATE_METHOD));
        ep->addEnumValue("First Order Update", ENUM_VUPDATE_FIRST_ORDER);
        ep->addEnumValue("Second Order Update", ENUM_VUPDATE_SECOND_ORDER);
    }

    void packParticles(ParticleData &pd, unsigned n) {
        m_x.resize(3 * (size_t)n); m_v.resize(3 * (size_t)n);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < (int)n; i++)
            for (int k = 0; k < 3; k++) { m_x[3 * i + k] = (float)pd.getPosition(i)[k]; m_v[3 * i + k] = (float)pd.getVelocity(i)[k]; }
    }

    bool uploadBodies(SimulationModel::RigidBodyVector &rbs) {
        const size_t m = rbs.size();
        std::vector<float> mass(m), x(3 * m), q(4 * m), I(3 * m), v(3 * m), w(3 * m);
        for (size_t i = 0; i < m; i++) {
            RigidBody *rb = rbs[i];
            mass[i] = (float)rb->getMass();
            const Quaternionr &r = rb->getRotation();
            q[4 * i] = (float)r.w(); q[4 * i + 1] = (float)r.x(); q[4 * i + 2] = (float)r.y(); q[4 * i + 3] = (float)r.z();
            for (int k = 0; k < 3; k++) {
                x[3 * i + k] = (float)rb->getPosition()[k]; I[3 * i + k] = (float)rb->getInertiaTensor()[k];
                v[3 * i + k] = (float)rb->getVelocity()[k]; w[3 * i + k] = (float)rb->getAngularVelocity()[k];
            }
        }
        if (pbd_set_rigid_bodies(m_engine, (unsigned)m, mass.data(), x.data(), q.data(), I.data(), v.data(), w.data())) { fail(); return false; }
        return true;
    }
    bool downloadBodies(SimulationModel::RigidBodyVector &rbs) {
        const size_t m = rbs.size();
        std::vector<float> x(3 * m), q(4 * m), v(3 * m), w(3 * m);
        if (pbd_get_rigid_bodies(m_engine, x.data(), q.data(), v.data(), w.data())) { fail(); return false; }
        for (size_t i = 0; i < m; i++) {
            RigidBody *rb = rbs[i];
            if (rb->getMass() == 0.0) continue;
            rb->setPosition(Vector3r((Real)x[3 * i], (Real)x[3 * i + 1], (Real)x[3 * i + 2]));
            rb->setRotation(Quaternionr((Real)q[4 * i], (Real)q[4 * i + 1], (Real)q[4 * i + 2], (Real)q[4 * i + 3]));
            rb->rotationUpdated();  // RigidBody.h:190-207
            rb->setVelocity(Vector3r((Real)v[3 * i], (Real)v[3 * i + 1], (Real)v[3 * i + 2]));
            rb->setAngularVelocity(Vector3r((Real)w[3 * i], (Real)w[3 * i + 1], (Real)w[3 * i + 2]));
        }
        return true;
    }

    // flatten the reference model through its public members (the data members of Constraints.h:255-457 are all public);
    // insertion index = position in getConstraints() = the id the reference's colouring refers to
    bool bind(SimulationModel &model) {
        ParticleData &pd = model.getParticles();
        SimulationModel::RigidBodyVector &rbs = model.getRigidBodies();
        const unsigned n = pd.size();
        std::vector<float> mass(n), x0(3 * (size_t)n);
        packParticles(pd, n);
        for (unsigned i = 0; i < n; i++) { mass[i] = (float)pd.getMass(i); for (int k = 0; k < 3; k++) x0[3 * i + k] = (float)pd.getPosition0(i)[k]; }
        if (pbd_set_particles(m_engine, n, m_x.data(), x0.data(), m_v.data(), mass.data())) { fail(); return false; }
        if (n) {
            for (unsigned i = 0; i < n; i++) for (int k = 0; k < 3; k++) x0[3 * i + k] = (float)pd.getOldPosition(i)[k];
            if (pbd_set_attr(m_engine, PBD_ATTR_OLDX, x0.data())) { fail(); return false; }
            for (unsigned i = 0; i < n; i++) for (int k = 0; k < 3; k++) x0[3 * i + k] = (float)pd.getLastPosition(i)[k];
            if (pbd_set_attr(m_engine, PBD_ATTR_LASTX, x0.data())) { fail(); return false; }
        }
        if (!uploadBodies(rbs)) return false;
        if (pbd_clear_constraints(m_engine)) { fail(); return false; }

        SimulationModel::ConstraintVector &cs = model.getConstraints();
        std::vector<unsigned> bodies[PBD_NUM_TYPES], ids[PBD_NUM_TYPES];
        std::vector<float> params[PBD_NUM_TYPES];
        for (unsigned id = 0; id < cs.size(); id++) {
            Constraint *c = cs[id];
            float p[24]; int np = 0, type = -1;
            const int tid = c->getTypeId();  // runtime ids from IDFactory (Constraints.cpp:17-49): compare, never hard-code
            auto mat = [&](const auto &M, int rows, int cols) { for (int r = 0; r < rows; r++) for (int cc = 0; cc < cols; cc++) p[np++] = (float)M(r, cc); };
            if (tid == DistanceConstraint::TYPE_ID) { auto *d = static_cast<DistanceConstraint *>(c); type = PBD_DISTANCE; p[np++] = (float)d->m_restLength; p[np++] = (float)d->m_stiffness; }
            else if (tid == DistanceConstraint_XPBD::TYPE_ID) { auto *d = static_cast<DistanceConstraint_XPBD *>(c); type = PBD_DISTANCE_XPBD; p[np++] = (float)d->m_restLength; p[np++] = (float)d->m_stiffness; }
            else if (tid == DihedralConstraint::TYPE_ID) { auto *d = static_cast<DihedralConstraint *>(c); type = PBD_DIHEDRAL; p[np++] = (float)d->m_restAngle; p[np++] = (float)d->m_stiffness; }
            else if (tid == IsometricBendingConstraint::TYPE_ID) { auto *d = static_cast<IsometricBendingConstraint *>(c); type = PBD_ISOBENDING; p[np++] = (float)d->m_stiffness; mat(d->m_Q, 4, 4); }
            else if (tid == IsometricBendingConstraint_XPBD::TYPE_ID) { auto *d = static_cast<IsometricBendingConstraint_XPBD *>(c); type = PBD_ISOBENDING_XPBD; p[np++] = (float)d->m_stiffness; mat(d->m_Q, 4, 4); }
            else if (tid == FEMTriangleConstraint::TYPE_ID) { auto *d = static_cast<FEMTriangleConstraint *>(c); type = PBD_FEMTRIANGLE; p[np++] = (float)d->m_area; mat(d->m_invRestMat, 2, 2);
                p[np++] = (float)d->m_xxStiffness; p[np++] = (float)d->m_yyStiffness; p[np++] = (float)d->m_xyStiffness; p[np++] = (float)d->m_xyPoissonRatio; p[np++] = (float)d->m_yxPoissonRatio; }
            else if (tid == StrainTriangleConstraint::TYPE_ID) { auto *d = static_cast<StrainTriangleConstraint *>(c); type = PBD_STRAINTRIANGLE; mat(d->m_invRestMat, 2, 2);
                p[np++] = (float)d->m_xxStiffness; p[np++] = (float)d->m_yyStiffness; p[np++] = (float)d->m_xyStiffness; p[np++] = d->m_normalizeStretch ? 1.f : 0.f; p[np++] = d->m_normalizeShear ? 1.f : 0.f; }
            else if (tid == VolumeConstraint::TYPE_ID) { auto *d = static_cast<VolumeConstraint *>(c); type = PBD_VOLUME; p[np++] = (float)d->m_restVolume; p[np++] = (float)d->m_stiffness; }
            else if (tid == VolumeConstraint_XPBD::TYPE_ID) { auto *d = static_cast<VolumeConstraint_XPBD *>(c); type = PBD_VOLUME_XPBD; p[np++] = (float)d->m_restVolume; p[np++] = (float)d->m_stiffness; }
            else if (tid == FEMTetConstraint::TYPE_ID) { auto *d = static_cast<FEMTetConstraint *>(c); type = PBD_FEMTET; p[np++] = (float)d->m_volume; mat(d->m_invRestMat, 3, 3); p[np++] = (float)d->m_stiffness; p[np++] = (float)d->m_poissonRatio; }
            else if (tid == XPBD_FEMTetConstraint::TYPE_ID) { auto *d = static_cast<XPBD_FEMTetConstraint *>(c); type = PBD_FEMTET_XPBD; p[np++] = (float)d->m_volume; mat(d->m_invRestMat, 3, 3); p[np++] = (float)d->m_stiffness; p[np++] = (float)d->m_poissonRatio; }
            else if (tid == StrainTetConstraint::TYPE_ID) { auto *d = static_cast<StrainTetConstraint *>(c); type = PBD_STRAINTET; mat(d->m_invRestMat, 3, 3);
                p[np++] = (float)d->m_stretchStiffness; p[np++] = (float)d->m_shearStiffness; p[np++] = d->m_normalizeStretch ? 1.f : 0.f; p[np++] = d->m_normalizeShear ? 1.f : 0.f; }
            else if (tid == ShapeMatchingConstraint::TYPE_ID) { auto *d = static_cast<ShapeMatchingConstraint *>(c);
                if (d->numberOfBodies() != 4) { m_error = "ShapeMatchingConstraint: only clusters of 4 particles run on the GPU path"; return false; }
                type = PBD_SHAPEMATCHING; p[np++] = (float)d->m_stiffness;
                for (int k = 0; k < 3; k++) p[np++] = (float)d->m_restCm[k];
                for (int qq = 0; qq < 4; qq++) for (int k = 0; k < 3; k++) p[np++] = (float)d->m_x0[qq][k];
                for (int qq = 0; qq < 4; qq++) p[np++] = (float)d->m_w[qq];
                for (int qq = 0; qq < 4; qq++) p[np++] = (float)d->m_numClusters[qq]; }
            else if (tid == BallJoint::TYPE_ID) { auto *d = static_cast<BallJoint *>(c); type = PBD_BALLJOINT; for (int col = 0; col < 4; col++) for (int k = 0; k < 3; k++) p[np++] = (float)d->m_jointInfo(k, col); }
            else if (tid == RigidBodyParticleBallJoint::TYPE_ID) { auto *d = static_cast<RigidBodyParticleBallJoint *>(c); type = PBD_RB_PARTICLE_BALLJOINT; for (int col = 0; col < 2; col++) for (int k = 0; k < 3; k++) p[np++] = (float)d->m_jointInfo(k, col); }
            else { m_error = "constraint type id " + std::to_string(tid) + " is not on the GPU path (SURVEY section 8: out of scope)"; return false; }
            const int nb = pbd_num_bodies(type), npar = pbd_num_params(type);
            for (int k = 0; k < nb; k++) bodies[type].push_back(c->m_bodies[k]);
            for (int k = 0; k < npar; k++) params[type].push_back(k < np ? p[k] : 0.0f);
            ids[type].push_back(id);
        }
        for (int t = 0; t < PBD_NUM_TYPES; t++)
            if (!ids[t].empty() && pbd_add_constraints(m_engine, t, (unsigned)ids[t].size(), bodies[t].data(), params[t].data(), ids[t].data())) { fail(); return false; }

        if (!model.m_groupsInitialized) model.initConstraintGroups();  // the reference's own colouring (SimulationModel.cpp:1033-1094)
        SimulationModel::ConstraintGroupVector &groups = model.getConstraintGroups();
        std::vector<unsigned> off(groups.size() + 1, 0), gids;
        for (size_t gi = 0; gi < groups.size(); gi++) { gids.insert(gids.end(), groups[gi].begin(), groups[gi].end()); off[gi + 1] = (unsigned)gids.size(); }
        if (pbd_set_groups(m_engine, (unsigned)groups.size(), off.data(), gids.data())) { fail(); return false; }
        m_boundConstraints = cs.size(); m_boundParticles = n; m_boundBodies = rbs.size();
        m_signature = sentinelSignature(model);
        m_mass = mass;
        m_error.clear();
        return true;
    }
};

#ifdef PBD_GPU_TIMESTEP_IMPLEMENTATION
int GpuTimeStepController::NUM_SUB_STEPS = -1;
int GpuTimeStepController::MAX_ITERATIONS = -1;
int GpuTimeStepController::MAX_ITERATIONS_V = -1;
int GpuTimeStepController::VELOCITY_UPDATE_METHOD = -1;
int GpuTimeStepController::ENUM_VUPDATE_FIRST_ORDER = -1;
int GpuTimeStepController::ENUM_VUPDATE_SECOND_ORDER = -1;
#endif

}  // namespace PBD
