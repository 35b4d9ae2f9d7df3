/* include/pbd_b200_model.h -- C ABI of the host-side model mirror (libpbd_b200.so).
 *
 * For FFI users that do not link the reference: the same calls a pyPBD script makes
 * (pyPBD/SimulationModelModule.cpp:123-298, pyPBD/ParticleDataModule.cpp, pyPBD/TimeStepModule.cpp:15-31,
 * pyPBD/ParameterObjectModule.cpp:19-29), exposed with plain pointers.  Scene construction (mesh topology,
 * constraint initialisation, greedy colouring) is host code and works without a GPU; pbdm_timestep_* needs one.
 * Return convention: int status functions return 0 on success; "add" functions return the reference's bool
 * (1 = constraint added, 0 = degenerate rest state, nothing added; SimulationModel.cpp:565-575).
 */
#ifndef PBD_B200_MODEL_H
#define PBD_B200_MODEL_H
#include "pbd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pbdm_model pbdm_model;        /* PBD::SimulationModel   (Simulation/SimulationModel.h:134-327) */
typedef struct pbdm_timestep pbdm_timestep;  /* PBD::TimeStepController (Simulation/TimeStepController.h) + TimeManager */

pbdm_model *pbdm_model_create(void);
void pbdm_model_destroy(pbdm_model *m);
void pbdm_model_reset(pbdm_model *m);    /* SimulationModel::reset   (SimulationModel.cpp:270-304) */
void pbdm_model_cleanup(pbdm_model *m);  /* SimulationModel::cleanup (SimulationModel.cpp:105-126) */

/* meshes: translation[3], rotation[9] row-major, scale[2|3]  (SimulationModel.cpp:808-1005) */
void pbdm_add_regular_triangle_model(pbdm_model *m, int width, int height, const float *translation, const float *rotation, const float *scale);
void pbdm_add_regular_tet_model(pbdm_model *m, int width, int height, int depth, const float *translation, const float *rotation, const float *scale);
void pbdm_add_triangle_model(pbdm_model *m, unsigned nPoints, unsigned nFaces, const float *points, const unsigned *indices);
void pbdm_add_tet_model(pbdm_model *m, unsigned nPoints, unsigned nTets, const float *points, const unsigned *indices);

/* ParticleData (Simulation/ParticleData.h:86-311); attr = enum pbd_attr */
unsigned pbdm_num_particles(pbdm_model *m);
void pbdm_set_mass(pbdm_model *m, unsigned i, float mass);
float pbdm_get_mass(pbdm_model *m, unsigned i);
float pbdm_get_inv_mass(pbdm_model *m, unsigned i);
void pbdm_get_masses(pbdm_model *m, float *mass, float *invMass);  /* n floats each */
int pbdm_get_particle(pbdm_model *m, int attr, unsigned i, float *out3);
int pbdm_set_particle(pbdm_model *m, int attr, unsigned i, const float *in3);
int pbdm_get_particles(pbdm_model *m, int attr, float *out);       /* n*3 floats */
int pbdm_set_particles(pbdm_model *m, int attr, const float *in);  /* n*3 floats */
const float *pbdm_vertices(pbdm_model *m);  /* zero-copy view of m_x (pyPBD getVertices); pulls from the device first */

/* rigid bodies coupled to particles through ball joints: RigidBody::initBody(mass, x, inertiaTensor, rotation)
 * (Simulation/RigidBody.h:84-120, q = (w,x,y,z)), SimulationModel::addBallJoint / addRigidBodyParticleBallJoint
 * (SimulationModel.cpp:306-317, 449-460) via pbdm_add_constraint(PBD_BALLJOINT, {rb0, rb1}, pos[3]) and
 * pbdm_add_constraint(PBD_RB_PARTICLE_BALLJOINT, {rb, particle}, NULL). */
unsigned pbdm_add_rigid_body(pbdm_model *m, float mass, const float *x3, const float *inertia3, const float *q4);
unsigned pbdm_num_rigid_bodies(pbdm_model *m);
/* RigidBody::setMass (RigidBody.h:277-284): invMass = mass != 0 ? 1 / mass : 0; a body with mass 0 is static */
int pbdm_set_rigid_body_mass(pbdm_model *m, unsigned i, float mass);
float pbdm_get_rigid_body_mass(pbdm_model *m, unsigned i);
void pbdm_get_rigid_bodies(pbdm_model *m, float *out13);  /* per body: x(3) q(w,x,y,z) v(3) omega(3) */

/* constraints */
int pbdm_add_constraint(pbdm_model *m, int type, const unsigned *bodies, const float *args);  /* args = the add<X>Constraint arguments after the indices */
void pbdm_add_cloth_constraints(pbdm_model *m, unsigned triModel, unsigned clothMethod, float distanceStiffness, float xxStiffness, float yyStiffness,
                                float xyStiffness, float xyPoissonRatio, float yxPoissonRatio, int normalizeStretch, int normalizeShear);
void pbdm_add_bending_constraints(pbdm_model *m, unsigned triModel, unsigned bendingMethod, float stiffness);
void pbdm_add_solid_constraints(pbdm_model *m, unsigned tetModel, unsigned solidMethod, float stiffness, float poissonRatio, float volumeStiffness,
                                int normalizeStretch, int normalizeShear);
unsigned pbdm_num_constraints(pbdm_model *m);
int pbdm_get_constraint(pbdm_model *m, unsigned i, int *type, unsigned *bodies, float *params);
void pbdm_get_constraints(pbdm_model *m, int *types, unsigned *bodies /*4 per*/, float *params /*24 per*/);
void pbdm_init_constraint_groups(pbdm_model *m);  /* SimulationModel::initConstraintGroups (SimulationModel.cpp:1033-1094) */
unsigned pbdm_num_groups(pbdm_model *m);
void pbdm_get_groups(pbdm_model *m, unsigned *offsets, unsigned *ids);
/* global parameter setters (SimulationModel.cpp:1351-1485); which: see enum */
enum pbdm_param { PBDM_CLOTH_STIFFNESS = 0, PBDM_CLOTH_STIFFNESS_XX, PBDM_CLOTH_STIFFNESS_YY, PBDM_CLOTH_STIFFNESS_XY, PBDM_CLOTH_POISSON_XY,
                  PBDM_CLOTH_POISSON_YX, PBDM_CLOTH_BENDING_STIFFNESS, PBDM_CLOTH_NORMALIZE_STRETCH, PBDM_CLOTH_NORMALIZE_SHEAR,
                  PBDM_SOLID_STIFFNESS, PBDM_SOLID_POISSON, PBDM_SOLID_VOLUME_STIFFNESS, PBDM_SOLID_NORMALIZE_STRETCH, PBDM_SOLID_NORMALIZE_SHEAR };
int pbdm_set_model_param(pbdm_model *m, int which, float value);

/* mesh topology views */
unsigned pbdm_num_triangle_models(pbdm_model *m);
unsigned pbdm_tri_num_edges(pbdm_model *m, unsigned tm);
unsigned pbdm_tri_num_faces(pbdm_model *m, unsigned tm);
unsigned pbdm_tri_index_offset(pbdm_model *m, unsigned tm);
void pbdm_tri_get_edges(pbdm_model *m, unsigned tm, unsigned *out);  /* v0 v1 f0 f1 per edge */
void pbdm_tri_get_faces(pbdm_model *m, unsigned tm, unsigned *out);
unsigned pbdm_num_tet_models(pbdm_model *m);
unsigned pbdm_tet_num_edges(pbdm_model *m, unsigned tm);
unsigned pbdm_tet_num_tets(pbdm_model *m, unsigned tm);
unsigned pbdm_tet_index_offset(pbdm_model *m, unsigned tm);
void pbdm_tet_get_edges(pbdm_model *m, unsigned tm, unsigned *out);  /* v0 v1 per edge */
void pbdm_tet_get_tets(pbdm_model *m, unsigned tm, unsigned *out);

/* standalone colouring of a flat constraint list: constraint c uses bodies[bodyOff[c]..bodyOff[c+1]) */
unsigned pbdm_first_fit_colouring(unsigned numBodies, unsigned numConstraints, const unsigned *bodyOff, const unsigned *bodies, unsigned *colourOut);

/* TimeStepController + TimeManager + Simulation::GRAVITATION.  ids: 0 NUM_SUB_STEPS, 1 MAX_ITERATIONS, 2 MAX_ITERATIONS_V, 3 VELOCITY_UPDATE_METHOD */
/* ---- collision objects (Simulation/DistanceFieldCollisionDetection.h, Simulation/CollisionDetection.h:15-104): the registry with the
 * reference's add* arguments; the tests run on the GPU (include/pbd_b200.h "Contact path").  bodyType: 0 rigid body, 1 triangle model,
 * 2 tet model (CollisionObject::*CollisionObjectType).  `shape` = pbd_collider_shape, `dims` = what addCollisionBox (full extents) /
 * Sphere (radius) / Torus (radii) / Cylinder (radius, height) / HollowSphere / HollowBox take; `vertices` (local, 3 floats each, may be
 * NULL) feed the bounding box as the body's mesh does in the reference. */
typedef struct pbdm_collision_detection pbdm_collision_detection;
pbdm_collision_detection *pbdm_cd_create(void);
void pbdm_cd_destroy(pbdm_collision_detection *cd);
void pbdm_cd_set_tolerance(pbdm_collision_detection *cd, float tolerance);
float pbdm_cd_get_tolerance(pbdm_collision_detection *cd);
int pbdm_cd_add_collision_shape(pbdm_collision_detection *cd, unsigned bodyIndex, unsigned bodyType, int shape, const float *dims, float thickness,
                                const float *vertices, unsigned numVertices, int testMesh, int invertSDF);
int pbdm_cd_add_collision_object_without_geometry(pbdm_collision_detection *cd, unsigned bodyIndex, unsigned bodyType, int testMesh);
unsigned pbdm_cd_num_collision_objects(pbdm_collision_detection *cd);
/* restitution / friction coefficients of a rigid body (kind 0), triangle model (1) or tet model (2): set*Coeff of the three classes */
int pbdm_set_contact_coefficients(pbdm_model *m, int kind, unsigned index, float restitution, float friction);
void pbdm_set_contact_stiffness_particle_rigid_body(pbdm_model *m, float stiffness);  /* SimulationModel.h:253-254 */
/* Frame of a body's geometry relative to the body frame (the reference's m_q_mat / m_x0_mat, RigidBody.h:172-188): distance fields on
 * the body are evaluated at x_local = R * R(q)^T (x_world - x) + t.  R row-major 3x3 (principal-axes matrix), t = centre of mass in the
 * geometry's coordinates; identity / zero by default. */
int pbdm_set_rigid_body_geometry_frame(pbdm_model *m, unsigned i, const float *R9, const float *t3);

pbdm_timestep *pbdm_timestep_create(int device, void *stream);  /* NULL + pbd_last_error() when no CUDA device */
void pbdm_timestep_destroy(pbdm_timestep *ts);
int pbdm_timestep_set_uint(pbdm_timestep *ts, int id, unsigned value);
unsigned pbdm_timestep_get_uint(pbdm_timestep *ts, int id);
int pbdm_timestep_set_int(pbdm_timestep *ts, int id, int value);
int pbdm_timestep_get_int(pbdm_timestep *ts, int id);
void pbdm_timestep_set_time_step_size(pbdm_timestep *ts, float h);
float pbdm_timestep_get_time_step_size(pbdm_timestep *ts);
float pbdm_timestep_get_time(pbdm_timestep *ts);
void pbdm_timestep_set_time(pbdm_timestep *ts, float t);
void pbdm_timestep_set_gravitation(pbdm_timestep *ts, const float *g3);
void pbdm_timestep_set_mode(pbdm_timestep *ts, int mode);
/* TimeStep::setCollisionDetection (Simulation/TimeStep.cpp:63-68); cd = NULL detaches */
void pbdm_timestep_set_collision_detection(pbdm_timestep *ts, pbdm_model *m, pbdm_collision_detection *cd);
int pbdm_timestep_step(pbdm_timestep *ts, pbdm_model *m);  /* TimeStep::step(SimulationModel&) (Simulation/TimeStep.h:41) */
const char *pbdm_timestep_error(pbdm_timestep *ts);
pbd_engine *pbdm_timestep_engine(pbdm_timestep *ts);

#ifdef __cplusplus
}
#endif
#endif
