/* oracle/_ref driver: a flat C interface over the UNMODIFIED reference sources.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is compiled together with the reference plugin's own
 *   /root/reference/openmmapi/src/{MatVec,RigidBody,eigenDecomposition,RigidBodySystem}.cpp
 * (in place, never copied) against the header shim in oracle/shim/, producing
 * oracle/_ref/librb_ref.so.  It drives RigidBodySystem exactly the way
 * ReferenceIntegrateRigidBodyStepKernel::execute does when there are no constraints or virtual
 * sites (platforms/reference/src/ReferenceRigidBodyKernels.cpp:82-108):
 *      integratePart1 -> (force evaluation) -> integratePart2
 * and RigidBodyIntegrator::stateChanged (openmmapi/src/RigidBodyIntegrator.cpp:63-74):
 *      update(geometry, velocities).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the resulting library.  The product library (librbk.so) never links or calls it.
 */
#include "RigidBodySystem.h"
#include "openmm/OpenMMException.h"
#include "openmm/System.h"
#include <cstring>
#include <string>
#include <vector>

using namespace RigidBodyPlugin;
using namespace OpenMM;
using std::vector;

// Defined (with external linkage) in openmmapi/src/RigidBodySystem.cpp:28-49.
vector<int> cleanBodyIndices(const vector<int>& bodyIndices);

namespace {

thread_local std::string lastError;

struct RefHandle {
    System system;
    ContextImpl context;
    RigidBodySystem bodies;
    vector<int> cleaned;
    int numAtoms;
    // analytic test potential:  U = sum_i 1/2 k |x_i - x0_i|^2 - c_i E.x_i
    bool tether, alternate;
    double k, E[3];
    vector<double> charge;
    vector<Vec3> x0;
    RefHandle() : context(system), numAtoms(0), tether(false), alternate(false), k(0.0) { E[0] = E[1] = E[2] = 0.0; }
};

void toVec(const double* src, vector<Vec3>& dst, int n) {
    dst.resize(n);
    if (n) std::memcpy((void*) &dst[0], src, sizeof(double)*3*n);
}

void fromVec(const vector<Vec3>& src, double* dst) {
    if (!src.empty()) std::memcpy(dst, (const void*) &src[0], sizeof(double)*3*src.size());
}

double tetherForces(RefHandle* h) {
    double U = 0.0;
    for (int i = 0; i < h->numAtoms; i++) {
        Vec3 dx = h->context.R[i] - h->x0[i];
        Vec3 Ev(h->E[0], h->E[1], h->E[2]);
        h->context.F[i] = dx*(-h->k) + Ev*h->charge[i];
        U += 0.5*h->k*dx.dot(dx) - h->charge[i]*Ev.dot(h->context.R[i]);
    }
    return U;
}

} // namespace

extern "C" {

const char* ref_last_error() { return lastError.c_str(); }

void* ref_create(int numAtoms, const int* bodyIndices, const double* masses, const unsigned char* isVirtual,
                 int numConstraints, const int* constraintAtoms, int rotationMode) {
    RefHandle* h = new RefHandle();
    try {
        h->numAtoms = numAtoms;
        for (int i = 0; i < numAtoms; i++) {
            h->system.addParticle(masses[i]);
            if (isVirtual != NULL && isVirtual[i]) h->system.setVirtualSite(i, true);
        }
        for (int i = 0; i < numConstraints; i++)
            h->system.addConstraint(constraintAtoms[2*i], constraintAtoms[2*i+1], 0.1);
        vector<int> idx(bodyIndices, bodyIndices + numAtoms);
        h->cleaned = cleanBodyIndices(idx);
        h->context.R.resize(numAtoms);
        h->context.V.resize(numAtoms);
        h->context.F.resize(numAtoms);
        h->bodies.initialize(h->context, idx, rotationMode);
    }
    catch (const std::exception& e) {
        lastError = e.what();
        delete h;
        return NULL;
    }
    return h;
}

void ref_destroy(void* p) { delete (RefHandle*) p; }

void ref_counts(void* p, int* out) {
    RefHandle* h = (RefHandle*) p;
    out[0] = h->bodies.getNumBodies();
    out[1] = h->bodies.getNumFree();
    out[2] = h->bodies.getNumActualAtoms();
    out[3] = h->bodies.getNumBodyAtoms();
    out[4] = h->bodies.getNumDOF();
    out[5] = h->numAtoms;
}

void ref_body_index(void* p, int* out) {
    RefHandle* h = (RefHandle*) p;
    for (int i = 0; i < h->numAtoms; i++) out[i] = h->cleaned[i];
}

void ref_atom_index(void* p, int* out) {
    RefHandle* h = (RefHandle*) p;
    int n = h->bodies.getNumFree() + h->bodies.getNumBodyAtoms();
    // NB: atomIndex has numActualAtoms slots, of which numFree + (sum of body N) are filled.
    int m = h->bodies.getNumActualAtoms();
    for (int i = 0; i < m; i++) out[i] = h->bodies.getAtomIndex(i);
    (void) n;
}

void ref_set_state(void* p, const double* R, const double* V, const double* F) {
    RefHandle* h = (RefHandle*) p;
    if (R) toVec(R, h->context.R, h->numAtoms);
    if (V) toVec(V, h->context.V, h->numAtoms);
    if (F) toVec(F, h->context.F, h->numAtoms);
}

void ref_get_state(void* p, double* R, double* V, double* F) {
    RefHandle* h = (RefHandle*) p;
    if (R) fromVec(h->context.R, R);
    if (V) fromVec(h->context.V, V);
    if (F) fromVec(h->context.F, F);
}

void ref_set_tether(void* p, double k, const double* E, const double* charge, const double* x0) {
    RefHandle* h = (RefHandle*) p;
    h->tether = true;
    h->k = k;
    for (int c = 0; c < 3; c++) h->E[c] = E[c];
    h->charge.assign(charge, charge + h->numAtoms);
    toVec(x0, h->x0, h->numAtoms);
}

/* Evaluate the analytic test potential at the current positions (fills F); returns U. */
double ref_compute_forces(void* p) {
    RefHandle* h = (RefHandle*) p;
    return h->tether ? tetherForces(h) : 0.0;
}

/* RigidBodyIntegrator::stateChanged -> RigidBodySystem::update(context, geometry, velocities). */
void ref_update(void* p, int geometry, int velocities) {
    RefHandle* h = (RefHandle*) p;
    h->bodies.update(h->context, geometry != 0, velocities != 0);
}

void ref_part1(void* p, double dt) {
    RefHandle* h = (RefHandle*) p;
    h->bodies.integratePart1(dt, h->context.F, h->context.V, h->context.R);
}

void ref_part2(void* p, double dt) {
    RefHandle* h = (RefHandle*) p;
    h->bodies.integratePart2(dt, h->context.R, h->context.F, h->context.V);
}

/* n steps of the Reference-platform execute() loop (no constraints / virtual sites). */
void ref_step(void* p, double dt, int steps) {
    RefHandle* h = (RefHandle*) p;
    for (int s = 0; s < steps; s++) {
        h->bodies.integratePart1(dt, h->context.F, h->context.V, h->context.R);
        if (h->tether) tetherForces(h);
        if (h->alternate)                       // benchmark workload: fixed forces whose sign flips every step
            for (int i = 0; i < h->numAtoms; i++) h->context.F[i] = -h->context.F[i];
        h->bodies.integratePart2(dt, h->context.R, h->context.F, h->context.V);
    }
}

void ref_set_alternate(void* p, int flag) { ((RefHandle*) p)->alternate = flag != 0; }

void ref_kinetic(void* p, double* out) {
    RefHandle* h = (RefHandle*) p;
    h->bodies.computeKineticEnergies(h->context.V);
    out[0] = h->bodies.getTranslationalEnergy();
    out[1] = h->bodies.getRotationalEnergy();
}

/* Dump per-body state (arrays sized numBodies x {1,3,4}); any pointer may be NULL. */
void ref_get_bodies(void* p, int* N, int* dof, int* loc, double* mass, double* I, double* invI, double* rcm,
                    double* pcm, double* q, double* pi, double* force, double* torque, double* twoK) {
    RefHandle* h = (RefHandle*) p;
    int nb = h->bodies.getNumBodies();
    for (int b = 0; b < nb; b++) {
        RigidBody body = h->bodies.getRigidBody(b);
        if (N) N[b] = body.N;
        if (dof) dof[b] = body.dof;
        if (loc) loc[b] = body.loc;
        if (mass) mass[b] = body.mass;
        for (int c = 0; c < 3; c++) {
            if (I) I[3*b+c] = body.I[c];
            if (invI) invI[3*b+c] = body.invI[c];
            if (rcm) rcm[3*b+c] = body.rcm[c];
            if (pcm) pcm[3*b+c] = body.pcm[c];
            if (force) force[3*b+c] = body.force[c];
        }
        for (int c = 0; c < 4; c++) {
            if (q) q[4*b+c] = body.q[c];
            if (pi) pi[4*b+c] = body.pi[c];
            if (torque) torque[4*b+c] = body.torque[c];
        }
        if (twoK) { twoK[2*b] = body.twoKt; twoK[2*b+1] = body.twoKr; }
    }
}

void ref_get_body_fixed(void* p, double* d) {
    RefHandle* h = (RefHandle*) p;
    int n = h->bodies.getNumBodyAtoms();
    for (int i = 0; i < n; i++) {
        Vec3 x = h->bodies.getBodyFixedPosition(i);
        d[3*i] = x[0]; d[3*i+1] = x[1]; d[3*i+2] = x[2];
    }
}

} // extern "C"
