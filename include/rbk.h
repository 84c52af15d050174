/* rbk.h - C ABI of librbk.so: the B200-native RigidBodyIntegrator step.
 *
 * This is the drop-in boundary for ONE hot path of craabreu/openmm_rigidbody_plugin: what its
 * IntegrateRigidBodyStepKernel implementations do (openmmapi/include/RigidBodyKernels.h:47-99).
 * Plain C, plain pointers and sizes; no C++ or torch types cross it.  Every entry point returns
 * 0 on success or a non-zero RBK_E* code, with a thread-local message in rbk_last_error()
 * (the reference throws OpenMM::OpenMMException; no exception crosses this boundary).
 *
 * Each function cites the reference interface it replaces (file:line relative to the reference
 * tree).  INTEGRATION.md shows the C++ glue (a KernelImpl subclass + kernel factory) that a
 * maintainer of the reference would add to call it.
 *
 * Threading/streams: no internal threads, no hidden synchronisation in rbk_part1/rbk_part2;
 * all device work is enqueued on the cudaStream_t passed as `stream` (void*, NULL = default
 * stream).  There is NO CPU fallback: without a CUDA device every device entry point fails.
 */
#ifndef RBK_H_
#define RBK_H_

#ifdef __cplusplus
extern "C" {
#endif

#define RBK_VERSION 1

/* error codes */
#define RBK_OK            0
#define RBK_EINVAL        1   /* bad argument (also: negative rotation mode)                       */
#define RBK_ECONSTRAINT   2   /* "Constraints involving rigid-body atoms are not allowed"          */
#define RBK_ECUDA         3   /* CUDA runtime error / no device                                     */
#define RBK_ESTATE        4   /* call order violated (e.g. part1 before upload)                    */
#define RBK_ENOMEM        5

/* atom-array layouts accepted by the device entry points (element type: double) */
#define RBK_LAYOUT_VEC3   0   /* xyzxyz...  = std::vector<OpenMM::Vec3>, the Reference platform's  */
#define RBK_LAYOUT_SOA    1   /* x[stride] y[stride] z[stride] planes                              */

typedef struct rbk_system rbk_system;

int         rbk_version(void);
const char* rbk_last_error(void);
/* Diagnostics: host<->device copies issued by librbk in this process so far - out[4] = H2D calls, H2D bytes, D2H calls,
 * D2H bytes.  The device entry points (rbk_part1/2*, rbk_part2_part1*, rbk_free_*) must not move any of them. */
int         rbk_debug_copy_counters(const rbk_system* sys, long long* out);
/* Diagnostics: the Taylor order the mode-0 kernels will use at their next launch: 11, 13 or 16 for systems of bodies of <= 4
 * atoms, 12, 13 or 16 for systems of large bodies ("series ladder": chosen on the device from the previous launch's
 * convergence statistics), the fixed order 12 otherwise. */
int         rbk_debug_series_order(rbk_system* sys, int* out, void* stream);
/* Diagnostics: kernel launches one call makes on the current system with fp64 arrays - out[3] = rbk_part1, rbk_part2,
 * rbk_part2_part1 (launch structure: DESIGN.md section 4; benchmarks report their launch counts from this). */
int         rbk_debug_launches_per_call(const rbk_system* sys, int* out);

/* ---- host model -------------------------------------------------------------------------- */

/* RigidBodySystem::initialize (openmmapi/src/RigidBodySystem.cpp:55-114) incl. cleanBodyIndices
 * (:28-49): compacts bodyIndices (<=0 -> free; distinct positive labels -> 1..nB in ascending
 * label order), counts free atoms / body sizes, lays out atomIndex = [free..., body 1..., ...]
 * and rejects constraints that touch body atoms.  isVirtual may be NULL (no virtual sites);
 * constraintAtoms holds 2*numConstraints atom indices.  rotationMode: 0 = exact, n>0 = NO-SQUISH
 * with n sub-steps (RigidBodyIntegrator::setRotationMode, openmmapi/src/RigidBodyIntegrator.cpp:26-32). */
int rbk_create(int numAtoms, const int* bodyIndices, const double* masses, const unsigned char* isVirtual,
               int numConstraints, const int* constraintAtoms, int rotationMode, rbk_system** out);
void rbk_destroy(rbk_system* sys);

/* out[5] = numBodies, numFree, numActualAtoms, numBodyAtoms, numDOF
 * (RigidBodySystem::getNum*, openmmapi/include/RigidBodySystem.h:37-42; numDOF is valid after
 * the first rbk_update with geometry != 0, RigidBodySystem.cpp:130-134). */
int rbk_get_counts(const rbk_system* sys, int* out);
int rbk_get_body_index(const rbk_system* sys, int* out);   /* [numAtoms]       cleaned index       */
int rbk_get_atom_index(const rbk_system* sys, int* out);   /* [numActualAtoms] RigidBodySystem::getAtomIndex */

/* RigidBodySystem::update (RigidBodySystem.cpp:120-142) = what RigidBodyIntegrator::stateChanged
 * triggers (RigidBodyIntegrator.cpp:63-74): rebuild body geometry (buildGeometry, RigidBody.cpp:65-116)
 * and/or dynamics (buildDynamics, :123-142) on the HOST model from host arrays R,V,F in
 * RBK_LAYOUT_VEC3 (any of them may be NULL when not needed by the requested parts). */
int rbk_update(rbk_system* sys, const double* R, const double* V, const double* F, int geometry, int velocities);

/* Host model dump (RigidBodySystem::getRigidBody / getBodyFixedPosition).  Arrays are per body,
 * row-major [nB][3|4]; any pointer may be NULL.  twoK = [nB][2] (2Kt, 2Kr). */
int rbk_get_host_bodies(const rbk_system* sys, int* N, int* dof, int* loc, double* mass, double* I, double* invI,
                        double* rcm, double* pcm, double* q, double* pi, double* force, double* torque, double* twoK);
int rbk_get_body_fixed(const rbk_system* sys, double* d);  /* [numBodyAtoms][3] */

/* ---- device ------------------------------------------------------------------------------- */

/* IntegrateRigidBodyStepKernel::uploadBodySystem (RigidBodyKernels.h:66; CUDA platform:
 * platforms/cuda/src/CudaRigidBodyKernels.cpp:293-372): (re)allocate device state on first use and
 * copy the whole host body model to the device in SoA form. */
int rbk_upload(rbk_system* sys, void* stream);

/* rbk_update + rbk_upload without leaving the device: RigidBodySystem::update (RigidBodySystem.cpp:120-142) evaluated by
 * one thread per body directly on the caller's DEVICE arrays, writing the body state and body-frame coordinates in
 * place (no host rebuild, no upload).  Same rounding as the host model except for libm's acos/cos (agreement ~1e-15).
 * The dynamics build SETS p = sum m v (the reference accumulates into its stale host copy when velocities are set
 * twice).  Afterwards the host copy of the bodies is stale: rbk_get_host_bodies fails, use rbk_download_bodies.
 * If the caller's atoms are reordered, call rbk_upload once (or this function) and rbk_set_atom_location first. */
int rbk_update_device(rbk_system* sys, const double* pos, const double* vel, const double* force, int layout,
                      long long stride, int geometry, int velocities, void* stream);

/* Replace the plugin-order -> caller-order atom map (CUDA platform's atomLocation,
 * CudaRigidBodyKernels.cpp:277-284 and ReorderListener :69-113).  location[i] is the index in the
 * caller's pos/vel/force arrays of plugin atom i (i indexes rbk_get_atom_index order).
 * NULL restores the default location[i] = atomIndex[i].  Allocates the handle's device state when called before any
 * rbk_upload / rbk_update_device (callers that build the bodies on the device). */
int rbk_set_atom_location(rbk_system* sys, const int* location, void* stream);

/* RigidBodySystem::integratePart1 (RigidBodySystem.cpp:170-187): free atoms half-kick + drift,
 * bodies half-kick, drift, rotation (exactRotation RigidBody.cpp:241-308 | noSquishRotation :220-231)
 * and atom-position reconstruction (updateAtomicPositions :148-153).
 * pos/vel/force are DEVICE pointers in `layout` (stride = plane stride in elements for SOA). */
int rbk_part1(rbk_system* sys, double dt, double* pos, double* vel, const double* force,
              int layout, long long stride, void* stream);

/* RigidBodySystem::integratePart2 (RigidBodySystem.cpp:193-204): free atoms second half-kick
 * (+ constraint displacement term), bodies forceAndTorque (RigidBody.cpp:174-183), second half
 * kick and atom-velocity reconstruction (updateAtomicVelocities :159-168). */
int rbk_part2(rbk_system* sys, double dt, const double* pos, double* vel, const double* force,
              int layout, long long stride, void* stream);

/* rbk_part2 of step k immediately followed by rbk_part1 of step k+1, as ONE pass over the data - legal because nothing
 * happens between the two in RigidBodyIntegrator::step (openmmapi/src/RigidBodyIntegrator.cpp:96-101: execute, execute,
 * ...; the force evaluation sits between Part 1 and Part 2 of the SAME step).  step(n) becomes
 *     part1, forces, [part2_part1, forces] x (n-1), part2.
 * Positions, velocities and r, p, q, pi are bit-identical to calling rbk_part2 then rbk_part1 with the same arguments
 * (water-size bodies: up to the summation order of a body's atom forces); the body state makes one round trip through
 * HBM instead of two and the body-frame coordinates are read once.  The stored body force / torque are NOT refreshed by
 * this call (the next rbk_part2 or rbk_part2_part1 recomputes them before anything reads them).  On return `vel` holds the velocities at the
 * end of step k and `pos` the positions after Part 1 of step k+1 (exactly the state the reference is in when it
 * evaluates forces).  Stream semantics: everything is ordered after the work already queued on `stream`, and `stream`
 * continues only when the whole call's work is done; for systems of large bodies WITH free atoms the free atoms either
 * ride along in the bodies' Part 2 kernel (here and in rbk_part2, when their list splits into at most 62 per atom tile) or are
 * integrated on a second stream owned by the handle (rbk_part1; all three calls otherwise), forked from and joined back to
 * `stream` with events (no host synchronisation; the pattern is legal under stream capture). */
int rbk_part2_part1(rbk_system* sys, double dt, double* pos, double* vel, const double* force,
                    int layout, long long stride, void* stream);

/* RigidBodySystem::computeKineticEnergies (RigidBodySystem.cpp:210-220) =
 * IntegrateRigidBodyStepKernel::getKineticEnergies (RigidBodyKernels.h:86): out[0] = translational,
 * out[1] = rotational kinetic energy.  Synchronises `stream` once to return the two doubles. */
int rbk_kinetic(rbk_system* sys, const double* vel, int layout, long long stride, double* out, void* stream);

/* Device body state -> host arrays [nB][3|4] (any pointer may be NULL); synchronises `stream`.
 * torque is the quaternion-frame 4-vector C(q)tau as the reference stores it (RigidBody.h:40). */
int rbk_download_bodies(rbk_system* sys, double* rcm, double* pcm, double* q, double* pi,
                        double* force, double* torque, void* stream);
/* (force / torque: RBK_ESTATE between rbk_part2_part1 and the closing rbk_part2 - interior steps keep them in registers) */

/* ---- host-buffer step (the reference-facing call measured as `e2e`) ------------------------ */

/* Force provider for rbk_execute_host: given positions R (host, VEC3) fill F (host, VEC3). */
typedef void (*rbk_force_fn)(const double* R, double* F, int numAtoms, void* user);

/* ReferenceIntegrateRigidBodyStepKernel::execute (platforms/reference/src/ReferenceRigidBodyKernels.cpp:82-108)
 * with HOST buffers, `steps` times: Part 1 on the device, positions copied to R, forces obtained
 * from `forces` (NULL = keep F as is) and copied to the device, Part 2, velocities copied to V.
 * R,V,F: host arrays in RBK_LAYOUT_VEC3 (pinned memory makes the copies asynchronous); on the first
 * call R,V,F are uploaded in full.  Uses device mirrors owned by the handle.
 *   forces != NULL: per step Part 1, R to the host, forces(R, F), F to the device, then Part 2 - for steps > 1 fused with
 *                   the next step's Part 1 (rbk_part2_part1).  F at entry = the forces at the current positions.
 *   forces == NULL: F = the forces at the NEW positions, known up front and used for every step of the call; its upload
 *                   runs on a copy stream underneath Part 1 and the download of R; R is downloaded after the last Part 1.
 * V is written once, after the last step of the call (nothing on the host reads velocities between the steps of
 * RigidBodyIntegrator::step(n)); V == NULL leaves the velocities on the device until a later call passes V. */
int rbk_execute_host(rbk_system* sys, double dt, int steps, double* R, double* V, double* F,
                     rbk_force_fn forces, void* user, void* stream);

/* The same step with the two host-side hooks the reference's Reference-platform kernel runs when the system
 * has free atoms (ReferenceRigidBodyKernels.cpp:92-104):
 *   constrainPositions(oldR, R, ...)  after Part 1 and before the forces - the place of
 *       ReferenceConstraints::apply(oldPos, R, invMass, tol) and ReferenceVirtualSites::computePositions;
 *       oldR holds the positions before the step OF THE FREE ATOMS (the only ones a solver may move; the entries of
 *       rigid-body atoms and virtual sites are unspecified), R the unconstrained new ones; return nonzero if R was changed
 *       (it is then copied back to the device, where Part 2 turns R - savedPos into the velocity correction of
 *       RigidBodySystem.cpp:196-197);
 *   constrainVelocities(R, V, ...)    after Part 2 - ReferenceConstraints::applyToVelocities(R, V, invMass, tol);
 *       return nonzero if V was changed (copied back to the device for the next step's free-atom kick).
 * Either hook may be NULL.  Only free atoms and virtual sites may be modified: the positions and velocities of
 * body atoms are outputs of the rigid-body state. */
typedef int (*rbk_positions_fn)(const double* oldR, double* R, int numAtoms, void* user);
typedef int (*rbk_velocities_fn)(const double* R, double* V, int numAtoms, void* user);
int rbk_execute_host_hooks(rbk_system* sys, double dt, int steps, double* R, double* V, double* F,
                           rbk_force_fn forces, rbk_positions_fn constrainPositions,
                           rbk_velocities_fn constrainVelocities, void* user, void* stream);

/* ---- OpenMM-CUDA boundary formats ------------------------------------------------------------
 * The device arrays OpenMM's CUDA platform hands to an integrator kernel (what the reference's CUDA kernels
 * read and write, platforms/cuda/src/kernels/rigidbodyintegrator.cu:30-64,270-279 and
 * platforms/cuda/src/CudaRigidBodyKernels.cpp:389-403): posq = real4 (xyz + charge), in mixed precision with a
 * second float4 array posqCorrection holding the low-order part; velm = mixed4 (xyz + inverse mass);
 * force = long long[3*paddedNumAtoms], one plane per component, fixed point with scale 2^32.  The kernels read
 * and write these formats directly (no conversion pass); .w components are never modified.  Atom order is
 * OpenMM's reordered order: pass the mapping with rbk_set_atom_location after every reorder. */
#define RBK_OPENMM_SINGLE 0   /* posq float4,            velm float4  */
#define RBK_OPENMM_MIXED  1   /* posq float4 + correction, velm double4 */
#define RBK_OPENMM_DOUBLE 2   /* posq double4,           velm double4 */
int rbk_part1_openmm(rbk_system* sys, double dt, void* posq, void* posqCorrection, void* velm, const long long* force,
                     int paddedNumAtoms, int precision, void* stream);
int rbk_part2_openmm(rbk_system* sys, double dt, void* posq, void* posqCorrection, void* velm, const long long* force,
                     int paddedNumAtoms, int precision, void* stream);
int rbk_kinetic_openmm(rbk_system* sys, const void* velm, int precision, double* out, void* stream);
/* rbk_part2_part1 on the OpenMM formats: Part 2 of step k + Part 1 of step k+1 in one pass, for callers that have nothing
 * between two execute() calls (CudaIntegrateRigidBodyStepKernel::execute, CudaRigidBodyKernels.cpp:377-444, when
 * context.updateContextState() is a no-op; cu.reorderAtoms() moves to the point after this call - the old forces are not
 * needed any more once both half kicks are done). */
int rbk_part2_part1_openmm(rbk_system* sys, double dt, void* posq, void* posqCorrection, void* velm, const long long* force,
                           int paddedNumAtoms, int precision, void* stream);
/* CudaIntegrateRigidBodyStepKernel::ReorderListener::execute (CudaRigidBodyKernels.cpp:78-108), on the device: move the
 * force of every atom the integrator owns from its old index to its new one (the next Part 1 kicks the free atoms with
 * them; force == NULL skips this) and install `location` like rbk_set_atom_location. */
int rbk_reorder_openmm(rbk_system* sys, const int* location, long long* force, int paddedNumAtoms, void* stream);

/* Free-atom constraint hooks of the CUDA flow (CudaRigidBodyKernels.cpp:405-421; kernels freeAtomsDelta and the
 * free-atom loop of integrateRigidBodyPart1, rigidbodyintegrator.cu:276-285, 303-312):
 *   rbk_free_delta_openmm      posDelta[i].xyz = (v + f invMass dt/2) dt for every free atom (mixed4 array:
 *                              float4 in single precision, double4 otherwise; .w untouched);
 *   <caller>                   integration.applyConstraints(tol) corrects posDelta;
 *   rbk_part1_delta_openmm     Part 1 in which free atoms move by posDelta instead of v dt (the difference between the
 *                              two, i.e. what the solver did, reaches the velocities in Part 2 as (x - savedPos)/dt -
 *                              the Reference platform's arithmetic; the reference's CUDA kernel drops that term);
 *   <caller>                   computeVirtualSites, forces, rbk_part2_openmm, applyVelocityConstraints. */
int rbk_free_delta_openmm(rbk_system* sys, double dt, const void* velm, const long long* force, int paddedNumAtoms,
                          int precision, void* posDelta, void* stream);
int rbk_part1_delta_openmm(rbk_system* sys, double dt, void* posq, void* posqCorrection, void* velm,
                           const long long* force, int paddedNumAtoms, int precision, const void* posDelta, void* stream);
int rbk_update_device_openmm(rbk_system* sys, void* posq, void* posqCorrection, void* velm, const long long* force,
                             int paddedNumAtoms, int precision, int geometry, int velocities, void* stream);

/* ---- refined ("shadow") energy diagnostics ---------------------------------------------------
 * RigidBodyIntegrator::setComputeRefinedEnergies / getRefinedKineticEnergies / getPotentialEnergyRefinement
 * (openmmapi/include/RigidBodyIntegrator.h:94-137, RigidBodyKernels.h:92-98), which the reference implements only in
 * its CUDA platform (COMPMOD paths of platforms/cuda/src/kernels/rigidbodyintegrator.cu:238-243,276-296,318-321,380-384,
 * 433-469; host side CudaRigidBodyKernels.cpp:118-194,405-438,481-494).  Once enabled, every Part 1 is preceded by a
 * virtual backward step and every Part 2 followed by a virtual forward step per body (two extra rotations per
 * body-step), accumulating third-order estimates of dr/dt and dq/dt; rbk_part2_part1 then runs the two kernels.
 *   RBK_REFINED_ALL    bodies and free atoms (flows without free-atom constraints);
 *   RBK_REFINED_BODIES bodies only - the caller accumulates the free-atom term itself with rbk_free_delta_openmm
 *                      / rbk_free_dot_openmm around its constraint solver, like CudaRigidBodyKernels.cpp:405-438
 *                      (factors -1, 5, 2; see DESIGN.md on the reference's -1).
 * rbk_refined_kinetic: out[0] = translational, out[1] = rotational refined kinetic energy (1/(6 dt) applied);
 * rbk_potential_refinement: out[0] = -(dt^2/24) [sum F.F/M + tau.(tau/I) + sum_free f.f/m].  Both synchronise. */
#define RBK_REFINED_OFF    0
#define RBK_REFINED_ALL    1
#define RBK_REFINED_BODIES 2
int rbk_set_refined_energies(rbk_system* sys, int mode, void* stream);
int rbk_refined_kinetic(rbk_system* sys, double dt, const double* vel, int layout, long long stride, double* out, void* stream);
int rbk_potential_refinement(rbk_system* sys, double dt, const double* force, int layout, long long stride, double* out,
                             void* stream);
int rbk_refined_kinetic_openmm(rbk_system* sys, double dt, const void* velm, int precision, double* out, void* stream);
int rbk_potential_refinement_openmm(rbk_system* sys, double dt, const long long* force, int paddedNumAtoms, double* out,
                                    void* stream);
/* freeAtomsDot (rigidbodyintegrator.cu:291-297): posDot = (restart ? 0 : posDot) + factor * posDelta.xyz for free atoms */
int rbk_free_dot_openmm(rbk_system* sys, const void* posDelta, int precision, double factor, int restart, void* stream);
/* host-buffer variants for the rbk_execute_host flow (V / F: host VEC3 arrays as returned by the last step) */
int rbk_refined_kinetic_host(rbk_system* sys, double dt, const double* V, double* out, void* stream);
int rbk_potential_refinement_host(rbk_system* sys, double dt, const double* F, double* out, void* stream);

/* rbk_kinetic for callers that hold velocities on the HOST (Reference-platform data): copies V into the
 * handle's device mirror (only free atoms need it) and runs the same device reduction. */
int rbk_kinetic_host(rbk_system* sys, const double* V, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RBK_H_ */
