/* rb_oracle.h - flat C interface of the CPU oracle (oracle/rb_oracle.c).
 *
 * TEST INFRASTRUCTURE ONLY: a plain-C, fp64, single-threaded restatement of the reference
 * plugin's Reference-platform arithmetic for the RigidBodyIntegrator step.  It exists to CHECK the
 * CUDA product (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs)
 * and is never linked, loaded or called by the product library librbk.so.
 *
 * The interface is deliberately identical (prefix orc_ instead of ref_) to oracle/ref_driver.cpp,
 * which wraps the unmodified reference sources, so the two can be diffed call by call.
 * Parity status: PINNED - tests/test_oracle.py checks this restatement against outputs of the
 * true reference (oracle/_ref, when built) and against the committed fixtures in tests/golden/
 * that were generated from the true reference by tests/golden/make_golden.py.
 */
#ifndef RB_ORACLE_H_
#define RB_ORACLE_H_
#ifdef __cplusplus
extern "C" {
#endif

const char* orc_last_error(void);
void* orc_create(int numAtoms, const int* bodyIndices, const double* masses, const unsigned char* isVirtual,
                 int numConstraints, const int* constraintAtoms, int rotationMode);
void orc_destroy(void* h);
void orc_counts(void* h, int* out6);          /* numBodies,numFree,numActualAtoms,numBodyAtoms,numDOF,numAtoms */
void orc_body_index(void* h, int* out);       /* cleaned body index per atom                                 */
void orc_atom_index(void* h, int* out);       /* [free atoms..., body 1 atoms..., body 2 atoms...]           */
void orc_set_state(void* h, const double* R, const double* V, const double* F);
void orc_get_state(void* h, double* R, double* V, double* F);
void orc_set_tether(void* h, double k, const double* E, const double* charge, const double* x0);
double orc_compute_forces(void* h);
void orc_update(void* h, int geometry, int velocities);
void orc_part1(void* h, double dt);
void orc_part2(void* h, double dt);
void orc_step(void* h, double dt, int steps);
void orc_set_alternate(void* h, int flag);     /* flip the sign of F before every Part 2 (benchmark workload) */
void orc_kinetic(void* h, double* out2);
void orc_get_bodies(void* h, int* N, int* dof, int* loc, double* mass, double* I, double* invI, double* rcm,
                    double* pcm, double* q, double* pi, double* force, double* torque, double* twoK);
void orc_get_body_fixed(void* h, double* d);

/* Refined ("shadow") energies of the reference's CUDA platform, restated in fp64 (PARITY UNPINNED, see rb_oracle.c). */
void   orc_set_refined(void* h, int flag);
void   orc_refined_kinetic(void* h, double dt, double* out2);
double orc_potential_refinement(void* h, double dt);

/* Scalar special functions, exported for known-answer tests against mpmath/scipy. */
void   orc_jacobi(double u, double m, double* sn, double* cn, double* dn);
double orc_carlson_rc(double x, double y);
double orc_carlson_rf(double x, double y, double z);
double orc_carlson_rj(double x, double y, double z, double p);
void   orc_exact_rotation(double dt, const double* I, double* q, double* pi);
void   orc_nosquish_rotation(double dt, int n, int dof, const double* invI, double* q, double* pi);

#ifdef __cplusplus
}
#endif
#endif
