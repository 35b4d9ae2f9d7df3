/* include/pbd_b200.h -- C ABI of the B200-native PBD/XPBD constraint-projection engine (libpbd_b200.so).
 *
 * Drop-in seam: this is what a `PBD::TimeStep` subclass on the reference side binds to replace
 * `TimeStepController::step` (Simulation/TimeStepController.cpp:75-241; seam: Simulation/TimeStep.h:13-48,
 * installed via Simulation::setTimeStep, Simulation/Simulation.h:48-49).  The reference-side adapter that
 * flattens a `PBD::SimulationModel` into these calls is integration/GpuTimeStepController.h (compiled inside the
 * unmodified reference by oracle/Makefile and parity-tested).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success, non-zero on error
 * (the reference's own convention is bool/void with no exceptions, Simulation/SimulationModel.cpp:565-575);
 * pbd_last_error() returns the message of the last failure on the calling thread.  Nothing throws across
 * this boundary.  One engine <-> one CUDA device + one stream; engines are independent.
 * There is NO CPU fallback: without a CUDA device every entry point that needs one fails.
 */
#ifndef PBD_B200_H
#define PBD_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct pbd_engine pbd_engine;

/* Flat constraint types and their per-constraint parameter layout (floats, in this order).  The layout is the
 * reference's own member set per constraint class (Simulation/Constraints.h:255-457), matrices row-major.
 *  type                     bodies  params                                                   reference class
 *  PBD_DISTANCE             2       restLength, stiffness                                    DistanceConstraint           Constraints.cpp:1166-1206
 *  PBD_DISTANCE_XPBD        2       restLength, stiffness                         (+lambda)  DistanceConstraint_XPBD      Constraints.cpp:1211-1258
 *  PBD_DIHEDRAL             4       restAngle, stiffness                                     DihedralConstraint           Constraints.cpp:1264-1339
 *  PBD_ISOBENDING           4       stiffness, Q[16]                                         IsometricBendingConstraint   Constraints.cpp:1345-1402
 *  PBD_ISOBENDING_XPBD      4       stiffness, Q[16]                              (+lambda)  IsometricBendingConstraint_XPBD :1407-1471
 *  PBD_FEMTRIANGLE          3       area, invRestMat[4], Exx, Eyy, Exy, nu_xy, nu_yx         FEMTriangleConstraint        Constraints.cpp:1476-1538
 *  PBD_STRAINTRIANGLE       3       invRestMat[4], kxx, kyy, kxy, normStretch, normShear     StrainTriangleConstraint     Constraints.cpp:1544-1610
 *  PBD_VOLUME               4       restVolume, stiffness                                    VolumeConstraint             Constraints.cpp:1617-1677
 *  PBD_VOLUME_XPBD          4       restVolume, stiffness                         (+lambda)  VolumeConstraint_XPBD        Constraints.cpp:1683-1750
 *  PBD_FEMTET               4       volume, invRestMat[9], E, nu                             FEMTetConstraint             Constraints.cpp:1755-1825
 *  PBD_FEMTET_XPBD          4       volume, invRestMat[9], E, nu                  (+lambda)  XPBD_FEMTetConstraint        Constraints.cpp:1830-1906
 *  PBD_STRAINTET            4       invRestMat[9], stretchK, shearK, normStretch, normShear  StrainTetConstraint          Constraints.cpp:1912-1980
 *  PBD_SHAPEMATCHING        4       stiffness, restCm[3], x0[4][3], w[4], numClusters[4]     ShapeMatchingConstraint (4-particle clusters) :1985-2028
 *  PBD_BALLJOINT            2 (rigid body, rigid body)      jointInfo[3x4] column by column  BallJoint                    Constraints.cpp:54-125
 *  PBD_RB_PARTICLE_BALLJOINT 2 (rigid body, particle)       jointInfo[3x2] column by column  RigidBodyParticleBallJoint   Constraints.cpp:925-987
 *  (joints: only the local connector columns are read, the global ones are recomputed at every solve as
 *   updateConstraint does; rigid bodies must be uploaded with pbd_set_rigid_bodies before joints are added)
 */
enum pbd_constraint_type {
    PBD_DISTANCE = 0, PBD_DISTANCE_XPBD = 1, PBD_DIHEDRAL = 2, PBD_ISOBENDING = 3, PBD_ISOBENDING_XPBD = 4,
    PBD_FEMTRIANGLE = 5, PBD_STRAINTRIANGLE = 6, PBD_VOLUME = 7, PBD_VOLUME_XPBD = 8, PBD_FEMTET = 9,
    PBD_FEMTET_XPBD = 10, PBD_STRAINTET = 11, PBD_SHAPEMATCHING = 12,
    PBD_BALLJOINT = 13, PBD_RB_PARTICLE_BALLJOINT = 14, PBD_NUM_TYPES = 15
};

/* particle attributes (ParticleData, Simulation/ParticleData.h:91-100); host layout = packed 3 floats per particle,
 * exactly `std::vector<Vector3r>::data()` of the reference's fp32 build. */
enum pbd_attr { PBD_ATTR_X = 0, PBD_ATTR_V = 1, PBD_ATTR_X0 = 2, PBD_ATTR_OLDX = 3, PBD_ATTR_LASTX = 4 };

enum pbd_solver_mode {
    PBD_MODE_GRAPH = 0,      /* one kernel per (colour,type) bucket, whole step replayed as a CUDA graph */
    PBD_MODE_RESIDENT = 1,   /* one launch per step: positions resident in the distributed shared memory of thread-block clusters,
                                hardware cluster barrier between colours (csrc/resident.cuh); scenes up to ~1.5 M particles */
    PBD_MODE_LAUNCH = 2,     /* plain stream launches (debug / per-kernel profiling) */
    PBD_MODE_JACOBI = 3,     /* comparison path, NOT the reference's algorithm: colours ignored, one launch per type and sweep, corrections
                                accumulated with float4 atomicAdd and averaged per particle (rigid-body joints are not supported) */
    PBD_MODE_AUTO = 4        /* default: RESIDENT where it is the faster exact mode (cloth and FEM models that fit), else GRAPH; both produce
                                the same bits, so the choice is invisible in the results (pbd_get_mode reports it) */
};

typedef struct pbd_stats {
    unsigned long long projections;    /* solvePositionConstraint calls executed so far (reference counting: early-outs included) */
    unsigned long long kernel_launches; /* kernels (or graph kernel nodes) launched by pbd_step so far */
    unsigned long long steps;          /* TimeStepController::step equivalents executed */
    unsigned num_particles, num_constraints, num_groups, num_buckets;
    unsigned constraints_per_type[PBD_NUM_TYPES];
    double bytes_per_step;             /* algorithmic bytes of one step */
    float last_step_ms;                /* device time of the last pbd_step call (CUDA events on the engine stream), after pbd_sync */
} pbd_stats;

const char *pbd_last_error(void);
int pbd_device_count(int *count);

/* stream: a cudaStream_t owned by the caller (e.g. torch.cuda.current_stream().cuda_stream) or NULL to let the
 * engine create its own non-blocking stream. */
int pbd_create(int device, void *stream, pbd_engine **out);
int pbd_destroy(pbd_engine *e);

/* ParticleData upload.  mass: n floats (invMass derived as ParticleData::setMass does, ParticleData.h:239-246).
 * x0/v may be NULL (x0 = x, v = 0); oldX = lastX = x as ParticleData::addVertex does (ParticleData.h:127-137). */
int pbd_set_particles(pbd_engine *e, unsigned n, const float *x, const float *x0, const float *v, const float *mass);
int pbd_set_attr(pbd_engine *e, int attr, const float *src);  /* host -> device, n*3 floats */
int pbd_get_attr(pbd_engine *e, int attr, float *dst);        /* device -> host, n*3 floats; synchronises */
int pbd_set_masses(pbd_engine *e, const float *mass);

/* Rigid bodies taking part in the coloured sweep through BallJoint / RigidBodyParticleBallJoint (SURVEY.md 8f-1;
 * Simulation/RigidBody.h:84-120 initBody with explicit mass and principal inertia).  mass[n], x[3n], q[4n] as (w,x,y,z),
 * inertia[3n] (principal moments), v[3n] / omega[3n] may be NULL (= 0).  old/last state = current, like initBody.
 * pbd_get_rigid_bodies: any output may be NULL. */
int pbd_set_rigid_bodies(pbd_engine *e, unsigned n, const float *mass, const float *x, const float *q, const float *inertia,
                         const float *v, const float *omega);
int pbd_get_rigid_bodies(pbd_engine *e, float *x, float *q, float *v, float *omega);

/* Constraints.  `bodies`: count*numBodies(type) particle indices; `params`: count*numParams(type) floats in the
 * layout above; `ids`: the constraint's index in the reference's SimulationModel::m_constraints (insertion order),
 * or NULL for "append after everything added so far".  Insertion order defines colouring and hence the result. */
int pbd_clear_constraints(pbd_engine *e);
int pbd_add_constraints(pbd_engine *e, int type, unsigned count, const unsigned *bodies, const float *params,
                        const unsigned *ids);
int pbd_num_bodies(int type);
int pbd_num_params(int type);

/* Colour groups: either import the reference's SimulationModel::m_constraintGroups (offsets[nGroups+1], ids by
 * insertion index) or recompute them with the reference's greedy first-fit (SimulationModel.cpp:1033-1094). */
int pbd_set_groups(pbd_engine *e, unsigned nGroups, const unsigned *offsets, const unsigned *ids);
int pbd_color_first_fit(pbd_engine *e);
/* The same colouring computed on the GPU (exact: wavefronts over the insertion-order dependency DAG, csrc/colouring.cuh);
 * identical groups.  ms / wavefronts (may be NULL): device time of the colouring and number of wavefronts. */
int pbd_color_first_fit_device(pbd_engine *e, float *ms, unsigned *wavefronts);
int pbd_get_num_groups(pbd_engine *e, unsigned *nGroups);
int pbd_get_groups(pbd_engine *e, unsigned *offsets, unsigned *ids);

/* TimeStepController knobs (NUM_SUB_STEPS, MAX_ITERATIONS, VELOCITY_UPDATE_METHOD: TimeStepController.cpp:47-72),
 * TimeManager step size (TimeManager.h:26-27) and Simulation::GRAVITATION (Simulation.cpp:63). */
int pbd_set_params(pbd_engine *e, float dt, unsigned subSteps, unsigned maxIter, int velocityUpdateMethod,
                   const float gravity[3]);
int pbd_set_mode(pbd_engine *e, int mode);
int pbd_get_mode(pbd_engine *e, int *requested, int *active);  /* active = the mode the current image runs in (resolved at the first step) */
/* sort constraints inside each (colour,type) bucket by their lowest particle index (order inside a colour is free). */
int pbd_set_bucket_sort(pbd_engine *e, int enable);

/* nSteps x TimeStepController::step, asynchronous on the engine's stream. */
int pbd_step(pbd_engine *e, unsigned nSteps);
int pbd_sync(pbd_engine *e);
/* Page-lock caller-owned host arrays (e.g. the reference's std::vector<Vector3r> storage) so that pbd_step_host copies at full
 * PCIe speed straight from / into them; unpin before the memory is freed or reallocated.  Thin wrappers over cudaHostRegister,
 * exported so that a reference-side adapter needs no CUDA headers. */
int pbd_pin_host(void *ptr, size_t bytes);
int pbd_unpin_host(void *ptr);
/* End-to-end convenience used by host-buffer callers (what the TimeStep adapter calls once per step): upload x and v
 * (3 floats/particle each, pinned or pageable; NULL = keep the device state), run nSteps, download x (and v if
 * v_out != NULL), synchronise.  Input and output buffers may be the same arrays. */
int pbd_step_host(pbd_engine *e, unsigned nSteps, const float *x_in, const float *v_in, float *x_out, float *v_out);
/* Pipelined form for callers that stream frames (a renderer reading frame k while frame k+1 is simulated, a batch driver): same
 * arguments and the same per-call copies, but the call only enqueues -- the upload of the next call and the download of the
 * previous one overlap the projection kernels (three streams, two staging slots).  The input arrays must stay untouched and the
 * output arrays unread until pbd_step_host_wait covers the call: lag = 0 waits for every call issued so far, lag = 1 for all but
 * the newest (so a loop "async(k); wait(1);" keeps exactly one call in flight behind the host).  Host arrays should be pinned
 * (pbd_pin_host), otherwise the copies are staged by the driver and do not overlap. */
int pbd_step_host_async(pbd_engine *e, unsigned nSteps, const float *x_in, const float *v_in, float *x_out, float *v_out);
int pbd_step_host_wait(pbd_engine *e, unsigned lag);

/* ---- Contact path (its data-parallel part) ----------------------------------------------------
 * Particles of triangle / tet models against analytic distance fields carried by STATIC rigid bodies, and the velocity-level
 * contact solve: DistanceFieldCollisionDetection::collisionDetection + collisionDetectionRBSolid
 * (Simulation/DistanceFieldCollisionDetection.cpp:26-197, 290-357; distance functions :598-728), TimeStep::contactCallbackFunction ->
 * SimulationModel::addParticleRigidBodyContactConstraint (Simulation/SimulationModel.cpp:538-549),
 * ParticleRigidBodyContactConstraint (Simulation/Constraints.cpp:2115-2186) and TimeStepController::velocityConstraintProjection
 * (Simulation/TimeStepController.cpp:189-196, 298-357).  Runs once per pbd_step after the substeps, as in the reference.
 * A collider on a rigid body with mass != 0 is refused (its contacts couple through the body and are inherently sequential);
 * rigid-rigid and particle-tet contacts are not covered -- such models stay on the reference's CPU time step.
 *   pbd_particle_collider: one DistanceFieldCollisionObjectWithoutGeometry (particles [offset, offset+count) of a model, the
 *                          model's restitution / friction coefficients);
 *   pbd_rigid_collider:    one DistanceFieldCollision{Box,Sphere,Torus,Cylinder,HollowSphere,HollowBox} in the order of
 *                          CollisionDetection::getCollisionObjects(): `dim` = m_box (half extents) | m_radius | m_radii | m_dim
 *                          (radius, half height), m_thickness, m_invertSDF, the body's coefficients, RigidBody::getTransformationR
 *                          (row-major) / V1 / V2 and the object's m_aabb (after updateAABB, i.e. extended by the tolerance). */
enum pbd_collider_shape { PBD_SHAPE_BOX = 0, PBD_SHAPE_SPHERE = 1, PBD_SHAPE_TORUS = 2, PBD_SHAPE_CYLINDER = 3, PBD_SHAPE_HOLLOW_SPHERE = 4, PBD_SHAPE_HOLLOW_BOX = 5 };
typedef struct pbd_particle_collider { unsigned offset, count; float restitution, friction; } pbd_particle_collider;
typedef struct pbd_rigid_collider {
    int shape; unsigned body;
    float dim[3], thickness; int invert_sdf;
    float restitution, friction;
    float R[9], v1[3], v2[3];
    float aabb_min[3], aabb_max[3];
} pbd_rigid_collider;
typedef struct pbd_contact { unsigned particle, body; float cp0[3], cp1[3], normal[3], dist; } pbd_contact;
int pbd_set_colliders(pbd_engine *e, unsigned nParticleColliders, const pbd_particle_collider *pc, unsigned nRigidColliders, const pbd_rigid_collider *rc);
/* CollisionDetection::m_tolerance, SimulationModel::m_contactStiffnessParticleRigidBody, TimeStepController::m_maxIterationsV */
int pbd_set_contact_params(pbd_engine *e, float tolerance, float stiffness, unsigned maxIterationsV);
/* debug / rendering: keep the contacts of every step (up to `capacity`); pbd_get_contacts returns those of the last step sorted by
 * (particle, body) and the number found (which may exceed what was kept) */
int pbd_record_contacts(pbd_engine *e, unsigned capacity);
int pbd_get_contacts(pbd_engine *e, pbd_contact *out, unsigned capacity, unsigned *count);

int pbd_get_lambdas(pbd_engine *e, int type, float *dst, unsigned *ids); /* debug: per-type XPBD multipliers + insertion ids */
int pbd_get_stats(pbd_engine *e, pbd_stats *out);
/* per-type device time of one profiled step (plain launches bracketed by CUDA events; ms per type + prologue/epilogue) */
int pbd_profile_step(pbd_engine *e, float *ms_per_type /*PBD_NUM_TYPES*/, float *ms_integrate, float *ms_velocity,
                     unsigned *launches_per_type /*PBD_NUM_TYPES*/);

#ifdef __cplusplus
}
#endif
#endif
