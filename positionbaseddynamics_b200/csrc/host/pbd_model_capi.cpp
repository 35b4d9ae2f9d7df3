// positionbaseddynamics_b200/csrc/host/pbd_model_capi.cpp -- extern "C" surface of include/pbd_b200_model.h over the
// host model mirror (pbd_model.h).  Plain pointers and sizes only.
#include "../../../include/pbd_b200_model.h"
#include "pbd_model.h"
#include <cstring>

using namespace pbd_b200;

struct pbdm_model { SimulationModel model; };
struct pbdm_timestep { TimeStepController ts; pbdm_timestep(int d, void *s) : ts(d, s) {} };
struct pbdm_collision_detection { DistanceFieldCollisionDetection cd; };

static inline Vector3r v3(const float *p) { return p ? Vector3r(p[0], p[1], p[2]) : Vector3r(); }
static inline Matrix3r m3(const float *p) { Matrix3r m = Matrix3r::Identity(); if (p) std::memcpy(m.m, p, sizeof(m.m)); return m; }

static std::vector<Vector3r> *attrVec(ParticleData &pd, int attr, bool forWrite) {
    if (forWrite) pd.touch(attr); else pd.pull(attr);
    switch (attr) {
    case PBD_ATTR_X: return &pd.m_x;
    case PBD_ATTR_V: return &pd.m_v;
    case PBD_ATTR_X0: return &pd.m_x0;
    case PBD_ATTR_OLDX: return &pd.m_oldX;
    case PBD_ATTR_LASTX: return &pd.m_lastX;
    default: return nullptr;
    }
}

extern "C" {

pbdm_model *pbdm_model_create(void) { return new pbdm_model(); }
void pbdm_model_destroy(pbdm_model *m) { delete m; }
void pbdm_model_reset(pbdm_model *m) { m->model.reset(); }
void pbdm_model_cleanup(pbdm_model *m) { m->model.cleanup(); }

void pbdm_add_regular_triangle_model(pbdm_model *m, int w, int h, const float *t, const float *R, const float *scale) {
    Vector2r s{{scale ? scale[0] : 1.0f, scale ? scale[1] : 1.0f}};
    m->model.addRegularTriangleModel(w, h, v3(t), m3(R), s);
}
void pbdm_add_regular_tet_model(pbdm_model *m, int w, int h, int d, const float *t, const float *R, const float *scale) {
    m->model.addRegularTetModel(w, h, d, v3(t), m3(R), scale ? v3(scale) : Vector3r(1, 1, 1));
}
void pbdm_add_triangle_model(pbdm_model *m, unsigned nPoints, unsigned nFaces, const float *points, const unsigned *indices) {
    m->model.addTriangleModel(nPoints, nFaces, reinterpret_cast<const Vector3r *>(points), indices);
}
void pbdm_add_tet_model(pbdm_model *m, unsigned nPoints, unsigned nTets, const float *points, const unsigned *indices) {
    m->model.addTetModel(nPoints, nTets, reinterpret_cast<const Vector3r *>(points), indices);
}

unsigned pbdm_num_particles(pbdm_model *m) { return m->model.getParticles().size(); }
void pbdm_set_mass(pbdm_model *m, unsigned i, float mass) { m->model.getParticles().setMass(i, mass); }
float pbdm_get_mass(pbdm_model *m, unsigned i) { return m->model.getParticles().getMass(i); }
float pbdm_get_inv_mass(pbdm_model *m, unsigned i) { return m->model.getParticles().getInvMass(i); }
void pbdm_get_masses(pbdm_model *m, float *mass, float *invMass) {
    const ParticleData &pd = m->model.getParticles();
    if (pd.size() == 0) return;
    if (mass) std::memcpy(mass, pd.m_masses.data(), pd.size() * sizeof(float));
    if (invMass) std::memcpy(invMass, pd.m_invMasses.data(), pd.size() * sizeof(float));
}
int pbdm_get_particle(pbdm_model *m, int attr, unsigned i, float *out3) {
    std::vector<Vector3r> *v = attrVec(m->model.getParticles(), attr, false);
    if (!v || i >= v->size()) return 1;
    std::memcpy(out3, (*v)[i].v, 3 * sizeof(float)); return 0;
}
int pbdm_set_particle(pbdm_model *m, int attr, unsigned i, const float *in3) {
    std::vector<Vector3r> *v = attrVec(m->model.getParticles(), attr, true);
    if (!v || i >= v->size()) return 1;
    std::memcpy((*v)[i].v, in3, 3 * sizeof(float)); return 0;
}
int pbdm_get_particles(pbdm_model *m, int attr, float *out) {
    std::vector<Vector3r> *v = attrVec(m->model.getParticles(), attr, false);
    if (!v) return 1;
    if (!v->empty()) std::memcpy(out, v->data(), v->size() * sizeof(Vector3r));
    return 0;
}
int pbdm_set_particles(pbdm_model *m, int attr, const float *in) {
    ParticleData &pd = m->model.getParticles();
    if (attr < 0 || attr > PBD_ATTR_LASTX) return 1;
    pd.aheadMask &= ~(1u << attr);  // a full overwrite needs no download first
    std::vector<Vector3r> *v = attrVec(pd, attr, true);
    if (!v->empty()) std::memcpy(v->data(), in, v->size() * sizeof(Vector3r));
    return 0;
}
const float *pbdm_vertices(pbdm_model *m) {
    const std::vector<Vector3r> &x = static_cast<const ParticleData &>(m->model.getParticles()).getVertices();
    return x.empty() ? nullptr : &x[0][0];
}

unsigned pbdm_add_rigid_body(pbdm_model *m, float mass, const float *x3, const float *inertia3, const float *q4) {
    RigidBody *rb = new RigidBody();
    rb->initBody(mass, v3(x3), v3(inertia3), q4 ? Quaternionr(q4[0], q4[1], q4[2], q4[3]) : Quaternionr());
    m->model.getRigidBodies().push_back(rb);
    m->model.m_groupsInitialized = false; m->model.rigidBodiesDirty = true;
    return (unsigned)m->model.getRigidBodies().size() - 1;
}
unsigned pbdm_num_rigid_bodies(pbdm_model *m) { return (unsigned)m->model.getRigidBodies().size(); }
int pbdm_set_rigid_body_mass(pbdm_model *m, unsigned i, float mass) {
    auto &rbs = m->model.getRigidBodies();
    if (i >= rbs.size()) return 1;
    rbs[i]->m_mass = mass; rbs[i]->m_invMass = (mass != 0.0f) ? 1.0f / mass : 0.0f;
    m->model.rigidBodiesDirty = true;
    return 0;
}
float pbdm_get_rigid_body_mass(pbdm_model *m, unsigned i) {
    auto &rbs = m->model.getRigidBodies();
    return i < rbs.size() ? rbs[i]->m_mass : 0.0f;
}
void pbdm_get_rigid_bodies(pbdm_model *m, float *out) {
    const auto &rbs = m->model.getRigidBodies();
    for (size_t i = 0; i < rbs.size(); i++) {
        float *o = out + 13 * i; const RigidBody &b = *rbs[i];
        for (int k = 0; k < 3; k++) { o[k] = b.m_x[k]; o[7 + k] = b.m_v[k]; o[10 + k] = b.m_omega[k]; }
        o[3] = b.m_q.w; o[4] = b.m_q.x; o[5] = b.m_q.y; o[6] = b.m_q.z;
    }
}

int pbdm_add_constraint(pbdm_model *m, int type, const unsigned *b, const float *a) {
    SimulationModel &M = m->model;
    switch (type) {
    case PBD_BALLJOINT: return M.addBallJoint(b[0], b[1], v3(a));
    case PBD_RB_PARTICLE_BALLJOINT: return M.addRigidBodyParticleBallJoint(b[0], b[1]);
    case PBD_DISTANCE: return M.addDistanceConstraint(b[0], b[1], a[0]);
    case PBD_DISTANCE_XPBD: return M.addDistanceConstraint_XPBD(b[0], b[1], a[0]);
    case PBD_DIHEDRAL: return M.addDihedralConstraint(b[0], b[1], b[2], b[3], a[0]);
    case PBD_ISOBENDING: return M.addIsometricBendingConstraint(b[0], b[1], b[2], b[3], a[0]);
    case PBD_ISOBENDING_XPBD: return M.addIsometricBendingConstraint_XPBD(b[0], b[1], b[2], b[3], a[0]);
    case PBD_FEMTRIANGLE: return M.addFEMTriangleConstraint(b[0], b[1], b[2], a[0], a[1], a[2], a[3], a[4]);
    case PBD_STRAINTRIANGLE: return M.addStrainTriangleConstraint(b[0], b[1], b[2], a[0], a[1], a[2], a[3] != 0, a[4] != 0);
    case PBD_VOLUME: return M.addVolumeConstraint(b[0], b[1], b[2], b[3], a[0]);
    case PBD_VOLUME_XPBD: return M.addVolumeConstraint_XPBD(b[0], b[1], b[2], b[3], a[0]);
    case PBD_FEMTET: return M.addFEMTetConstraint(b[0], b[1], b[2], b[3], a[0], a[1]);
    case PBD_FEMTET_XPBD: return M.addFEMTetConstraint_XPBD(b[0], b[1], b[2], b[3], a[0], a[1]);
    case PBD_STRAINTET: return M.addStrainTetConstraint(b[0], b[1], b[2], b[3], a[0], a[1], a[2] != 0, a[3] != 0);
    case PBD_SHAPEMATCHING: { const unsigned nc[4] = {(unsigned)a[1], (unsigned)a[2], (unsigned)a[3], (unsigned)a[4]}; return M.addShapeMatchingConstraint(4, b, nc, a[0]); }
    default: return 0;
    }
}
void pbdm_add_cloth_constraints(pbdm_model *m, unsigned tm, unsigned method, float dk, float xx, float yy, float xy, float pxy, float pyx, int ns, int nsh) {
    if (tm < m->model.getTriangleModels().size()) m->model.addClothConstraints(m->model.getTriangleModels()[tm], method, dk, xx, yy, xy, pxy, pyx, ns != 0, nsh != 0);
}
void pbdm_add_bending_constraints(pbdm_model *m, unsigned tm, unsigned method, float k) {
    if (tm < m->model.getTriangleModels().size()) m->model.addBendingConstraints(m->model.getTriangleModels()[tm], method, k);
}
void pbdm_add_solid_constraints(pbdm_model *m, unsigned tm, unsigned method, float k, float nu, float volK, int ns, int nsh) {
    if (tm < m->model.getTetModels().size()) m->model.addSolidConstraints(m->model.getTetModels()[tm], method, k, nu, volK, ns != 0, nsh != 0);
}
unsigned pbdm_num_constraints(pbdm_model *m) { return m->model.numConstraints(); }
int pbdm_get_constraint(pbdm_model *m, unsigned i, int *type, unsigned *bodies, float *params) {
    if (i >= m->model.numConstraints()) return 1;
    const ConstraintView v = m->model.getConstraint(i);
    *type = v.type;
    std::memcpy(bodies, v.m_bodies, v.numberOfBodies * sizeof(unsigned));
    std::memcpy(params, v.params, v.numParams * sizeof(float));
    return 0;
}
void pbdm_get_constraints(pbdm_model *m, int *types, unsigned *bodies, float *params) {
    const unsigned N = m->model.numConstraints();
    for (unsigned i = 0; i < N; i++) {
        const ConstraintView v = m->model.getConstraint(i);
        types[i] = v.type;
        for (unsigned k = 0; k < 4; k++) bodies[4 * (size_t)i + k] = k < v.numberOfBodies ? v.m_bodies[k] : 0xffffffffu;
        for (unsigned k = 0; k < 24; k++) params[24 * (size_t)i + k] = k < v.numParams ? v.params[k] : 0.0f;
    }
}
void pbdm_init_constraint_groups(pbdm_model *m) { m->model.initConstraintGroups(); }
unsigned pbdm_num_groups(pbdm_model *m) { m->model.initConstraintGroups(); return (unsigned)m->model.getConstraintGroups().size(); }
void pbdm_get_groups(pbdm_model *m, unsigned *offsets, unsigned *ids) {
    m->model.initConstraintGroups();
    const auto &g = m->model.getConstraintGroups();
    unsigned o = 0;
    for (size_t i = 0; i < g.size(); i++) { offsets[i] = o; for (unsigned id : g[i]) ids[o++] = id; }
    offsets[g.size()] = o;
}
int pbdm_set_model_param(pbdm_model *m, int which, float v) {
    SimulationModel &M = m->model;
    switch (which) {
    case PBDM_CLOTH_STIFFNESS: M.setClothStiffness(v); break;
    case PBDM_CLOTH_STIFFNESS_XX: M.setClothStiffnessXX(v); break;
    case PBDM_CLOTH_STIFFNESS_YY: M.setClothStiffnessYY(v); break;
    case PBDM_CLOTH_STIFFNESS_XY: M.setClothStiffnessXY(v); break;
    case PBDM_CLOTH_POISSON_XY: M.setClothPoissonRatioXY(v); break;
    case PBDM_CLOTH_POISSON_YX: M.setClothPoissonRatioYX(v); break;
    case PBDM_CLOTH_BENDING_STIFFNESS: M.setClothBendingStiffness(v); break;
    case PBDM_CLOTH_NORMALIZE_STRETCH: M.setClothNormalizeStretch(v != 0); break;
    case PBDM_CLOTH_NORMALIZE_SHEAR: M.setClothNormalizeShear(v != 0); break;
    case PBDM_SOLID_STIFFNESS: M.setSolidStiffness(v); break;
    case PBDM_SOLID_POISSON: M.setSolidPoissonRatio(v); break;
    case PBDM_SOLID_VOLUME_STIFFNESS: M.setSolidVolumeStiffness(v); break;
    case PBDM_SOLID_NORMALIZE_STRETCH: M.setSolidNormalizeStretch(v != 0); break;
    case PBDM_SOLID_NORMALIZE_SHEAR: M.setSolidNormalizeShear(v != 0); break;
    default: return 1;
    }
    return 0;
}

unsigned pbdm_num_triangle_models(pbdm_model *m) { return (unsigned)m->model.getTriangleModels().size(); }
unsigned pbdm_tri_num_edges(pbdm_model *m, unsigned tm) { return m->model.getTriangleModels()[tm]->getParticleMesh().numEdges(); }
unsigned pbdm_tri_num_faces(pbdm_model *m, unsigned tm) { return m->model.getTriangleModels()[tm]->getParticleMesh().numFaces(); }
unsigned pbdm_tri_index_offset(pbdm_model *m, unsigned tm) { return m->model.getTriangleModels()[tm]->getIndexOffset(); }
void pbdm_tri_get_edges(pbdm_model *m, unsigned tm, unsigned *out) {
    const auto &e = m->model.getTriangleModels()[tm]->getParticleMesh().getEdges();
    for (size_t i = 0; i < e.size(); i++) { out[4 * i] = e[i].m_vert[0]; out[4 * i + 1] = e[i].m_vert[1]; out[4 * i + 2] = e[i].m_face[0]; out[4 * i + 3] = e[i].m_face[1]; }
}
void pbdm_tri_get_faces(pbdm_model *m, unsigned tm, unsigned *out) {
    const auto &f = m->model.getTriangleModels()[tm]->getParticleMesh().getFaces();
    if (!f.empty()) std::memcpy(out, f.data(), f.size() * sizeof(unsigned));
}
unsigned pbdm_num_tet_models(pbdm_model *m) { return (unsigned)m->model.getTetModels().size(); }
unsigned pbdm_tet_num_edges(pbdm_model *m, unsigned tm) { return m->model.getTetModels()[tm]->getParticleMesh().numEdges(); }
unsigned pbdm_tet_num_tets(pbdm_model *m, unsigned tm) { return m->model.getTetModels()[tm]->getParticleMesh().numTets(); }
unsigned pbdm_tet_index_offset(pbdm_model *m, unsigned tm) { return m->model.getTetModels()[tm]->getIndexOffset(); }
void pbdm_tet_get_edges(pbdm_model *m, unsigned tm, unsigned *out) {
    const auto &e = m->model.getTetModels()[tm]->getParticleMesh().getEdges();
    for (size_t i = 0; i < e.size(); i++) { out[2 * i] = e[i].m_vert[0]; out[2 * i + 1] = e[i].m_vert[1]; }
}
void pbdm_tet_get_tets(pbdm_model *m, unsigned tm, unsigned *out) {
    const auto &t = m->model.getTetModels()[tm]->getParticleMesh().getTets();
    if (!t.empty()) std::memcpy(out, t.data(), t.size() * sizeof(unsigned));
}

unsigned pbdm_first_fit_colouring(unsigned numBodies, unsigned numConstraints, const unsigned *bodyOff, const unsigned *bodies, unsigned *colourOut) {
    std::vector<unsigned> colour;
    const unsigned n = firstFitColouring(numBodies, numConstraints, bodyOff, bodies, colour);
    if (numConstraints) std::memcpy(colourOut, colour.data(), numConstraints * sizeof(unsigned));
    return n;
}

pbdm_timestep *pbdm_timestep_create(int device, void *stream) {
    pbdm_timestep *t = new pbdm_timestep(device, stream);
    if (!t->ts.valid()) { delete t; return nullptr; }  // pbd_last_error() holds the reason (no CUDA device: no CPU fallback)
    return t;
}
void pbdm_timestep_destroy(pbdm_timestep *ts) { delete ts; }
int pbdm_timestep_set_uint(pbdm_timestep *ts, int id, unsigned v) { return ts->ts.setValueUInt(id, v) ? 0 : 1; }
unsigned pbdm_timestep_get_uint(pbdm_timestep *ts, int id) { return ts->ts.getValueUInt(id); }
int pbdm_timestep_set_int(pbdm_timestep *ts, int id, int v) { return ts->ts.setValueInt(id, v) ? 0 : 1; }
int pbdm_timestep_get_int(pbdm_timestep *ts, int id) { return ts->ts.getValueInt(id); }
void pbdm_timestep_set_time_step_size(pbdm_timestep *ts, float h) { ts->ts.timeManager().setTimeStepSize(h); }
float pbdm_timestep_get_time_step_size(pbdm_timestep *ts) { return ts->ts.timeManager().getTimeStepSize(); }
float pbdm_timestep_get_time(pbdm_timestep *ts) { return ts->ts.timeManager().getTime(); }
void pbdm_timestep_set_time(pbdm_timestep *ts, float t) { ts->ts.timeManager().setTime(t); }
void pbdm_timestep_set_gravitation(pbdm_timestep *ts, const float *g) { ts->ts.setGravitation(v3(g)); }
void pbdm_timestep_set_mode(pbdm_timestep *ts, int mode) { ts->ts.setSolverMode(mode); }
pbdm_collision_detection *pbdm_cd_create(void) { return new pbdm_collision_detection(); }
void pbdm_cd_destroy(pbdm_collision_detection *cd) { delete cd; }
void pbdm_cd_set_tolerance(pbdm_collision_detection *cd, float t) { cd->cd.setTolerance(t); }
float pbdm_cd_get_tolerance(pbdm_collision_detection *cd) { return cd->cd.getTolerance(); }
int pbdm_cd_add_collision_shape(pbdm_collision_detection *cd, unsigned bodyIndex, unsigned bodyType, int shape, const float *dims, float thickness,
                                const float *vertices, unsigned numVertices, int testMesh, int invertSDF) {
    if (!cd || !dims) return 1;
    std::vector<Vector3r> v(vertices ? numVertices : 0);
    for (size_t i = 0; i < v.size(); i++) v[i] = Vector3r(vertices[3 * i], vertices[3 * i + 1], vertices[3 * i + 2]);
    const Vector3r *vp = v.empty() ? nullptr : v.data();
    const unsigned nv = (unsigned)v.size();
    switch (shape) {
    case PBD_SHAPE_BOX: cd->cd.addCollisionBox(bodyIndex, bodyType, vp, nv, Vector3r(dims[0], dims[1], dims[2]), testMesh != 0, invertSDF != 0); break;
    case PBD_SHAPE_SPHERE: cd->cd.addCollisionSphere(bodyIndex, bodyType, vp, nv, dims[0], testMesh != 0, invertSDF != 0); break;
    case PBD_SHAPE_TORUS: cd->cd.addCollisionTorus(bodyIndex, bodyType, vp, nv, Vector2r{{dims[0], dims[1]}}, testMesh != 0, invertSDF != 0); break;
    case PBD_SHAPE_CYLINDER: cd->cd.addCollisionCylinder(bodyIndex, bodyType, vp, nv, Vector2r{{dims[0], dims[1]}}, testMesh != 0, invertSDF != 0); break;
    case PBD_SHAPE_HOLLOW_SPHERE: cd->cd.addCollisionHollowSphere(bodyIndex, bodyType, vp, nv, dims[0], thickness, testMesh != 0, invertSDF != 0); break;
    case PBD_SHAPE_HOLLOW_BOX: cd->cd.addCollisionHollowBox(bodyIndex, bodyType, vp, nv, Vector3r(dims[0], dims[1], dims[2]), thickness, testMesh != 0, invertSDF != 0); break;
    default: return 1;
    }
    return 0;
}
int pbdm_cd_add_collision_object_without_geometry(pbdm_collision_detection *cd, unsigned bodyIndex, unsigned bodyType, int testMesh) {
    if (!cd) return 1;
    cd->cd.addCollisionObjectWithoutGeometry(bodyIndex, bodyType, nullptr, 0, testMesh != 0);
    return 0;
}
unsigned pbdm_cd_num_collision_objects(pbdm_collision_detection *cd) { return (unsigned)cd->cd.getCollisionObjects().size(); }
int pbdm_set_contact_coefficients(pbdm_model *m, int kind, unsigned index, float restitution, float friction) {
    if (kind == 0) { auto &v = m->model.getRigidBodies(); if (index >= v.size()) return 1; v[index]->setRestitutionCoeff(restitution); v[index]->setFrictionCoeff(friction); return 0; }
    if (kind == 1) { auto &v = m->model.getTriangleModels(); if (index >= v.size()) return 1; v[index]->setRestitutionCoeff(restitution); v[index]->setFrictionCoeff(friction); return 0; }
    if (kind == 2) { auto &v = m->model.getTetModels(); if (index >= v.size()) return 1; v[index]->setRestitutionCoeff(restitution); v[index]->setFrictionCoeff(friction); return 0; }
    return 1;
}
int pbdm_set_rigid_body_geometry_frame(pbdm_model *m, unsigned i, const float *R9, const float *t3) {
    auto &v = m->model.getRigidBodies();
    if (i >= v.size() || !R9 || !t3) return 1;
    std::memcpy(v[i]->m_frameR.m, R9, sizeof(v[i]->m_frameR.m));
    v[i]->m_frameT = Vector3r(t3[0], t3[1], t3[2]);
    return 0;
}
void pbdm_set_contact_stiffness_particle_rigid_body(pbdm_model *m, float k) { m->model.setContactStiffnessParticleRigidBody(k); }
void pbdm_timestep_set_collision_detection(pbdm_timestep *ts, pbdm_model *m, pbdm_collision_detection *cd) {
    ts->ts.setCollisionDetection(m->model, cd ? &cd->cd : nullptr);
}

int pbdm_timestep_step(pbdm_timestep *ts, pbdm_model *m) { return ts->ts.step(m->model) ? 0 : 1; }
const char *pbdm_timestep_error(pbdm_timestep *ts) { return ts->ts.error().c_str(); }
pbd_engine *pbdm_timestep_engine(pbdm_timestep *ts) { return ts->ts.engine(); }

}  // extern "C"
