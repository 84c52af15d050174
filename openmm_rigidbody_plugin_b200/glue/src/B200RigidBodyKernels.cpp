// B200 kernel: every IntegrateRigidBodyStepKernel method forwards to the C ABI (include/rbk.h).
// Step protocol = ReferenceIntegrateRigidBodyStepKernel::execute (platforms/reference/src/ReferenceRigidBodyKernels.cpp:82-108)
// Part 1 -> [constraints among free atoms, virtual sites] -> forces at the new positions -> Part 2 -> [velocity constraints].
#include "B200RigidBodyKernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/ReferenceVirtualSites.h"

using namespace RigidBodyPlugin;
using namespace OpenMM;
using std::vector;

static void check(int rc) {
    if (rc != RBK_OK) throw OpenMMException(rbk_last_error());
}

void B200IntegrateRigidBodyStepKernel::initialize(ContextImpl& contextRef, const RigidBodyIntegrator& integrator) {
    context = &contextRef;
    const System& sys = contextRef.getSystem();
    const int numAtoms = sys.getNumParticles();
    invMass.resize(numAtoms);                                   // ReferenceRigidBodyKernels.cpp:68-75
    bool virtualSites = false;
    for (int i = 0; i < numAtoms; i++) {
        const double mass = sys.getParticleMass(i);
        invMass[i] = mass == 0.0 ? 0.0 : 1.0/mass;
        virtualSites = virtualSites || sys.isVirtualSite(i);
    }
    oldPos.resize(numAtoms);
    refined = integrator.getComputeRefinedEnergies();
    // the reference calls the constraint solver whenever there are free atoms (a no-op without constraints) and
    // ReferenceVirtualSites::computePositions always; the hooks are only installed when they have work to do
    velocityHook = sys.getNumConstraints() != 0;
    positionHook = velocityHook || virtualSites;
}

void B200IntegrateRigidBodyStepKernel::uploadBodySystem(RigidBodySystem& bodySystem) {
    bodies = &bodySystem;
    system = bodySystem.getHandle();
    check(rbk_upload(system, NULL));
    if (refined) check(rbk_set_refined_energies(system, RBK_REFINED_ALL, NULL));
}

void B200IntegrateRigidBodyStepKernel::evaluateForces(const double*, double*, int, void* self) {
    // positions have been copied back into the platform data; let OpenMM fill data.forces
    static_cast<B200IntegrateRigidBodyStepKernel*>(self)->context->calcForcesAndEnergy(true, false);
}

// After Part 1: ReferenceConstraints::apply(oldPos, R, invMass, tol) when there are free atoms, then
// ReferenceVirtualSites::computePositions (ReferenceRigidBodyKernels.cpp:98-100).  R aliases data.positions.
int B200IntegrateRigidBodyStepKernel::constrainPositions(const double* oldR, double*, int numAtoms, void* selfPtr) {
    B200IntegrateRigidBodyStepKernel* self = static_cast<B200IntegrateRigidBodyStepKernel*>(selfPtr);
    const System& sys = self->context->getSystem();
    if (self->velocityHook) {
        for (int i = 0; i < numAtoms; i++) self->oldPos[i] = Vec3(oldR[3*i], oldR[3*i+1], oldR[3*i+2]);
        self->data.constraints->apply(self->oldPos, *self->data.positions, self->invMass, self->tolerance);
    }
    ReferenceVirtualSites::computePositions(sys, *self->data.positions);
    return 1;
}

// After Part 2: ReferenceConstraints::applyToVelocities(R, V, invMass, tol) (ReferenceRigidBodyKernels.cpp:103-104).
int B200IntegrateRigidBodyStepKernel::constrainVelocities(const double*, double*, int, void* selfPtr) {
    B200IntegrateRigidBodyStepKernel* self = static_cast<B200IntegrateRigidBodyStepKernel*>(selfPtr);
    self->data.constraints->applyToVelocities(*self->data.positions, *self->data.velocities, self->invMass, self->tolerance);
    return 1;
}

void B200IntegrateRigidBodyStepKernel::execute(ContextImpl& contextRef, const RigidBodyIntegrator& integrator) {
    executeSteps(contextRef, integrator, 1);
}

// step(n) as one call of the host-buffer entry point: the forces go up and the positions come down every step (the
// Reference platform evaluates forces on the host), Part 2 of a step and Part 1 of the next run as one pass where no
// velocity hook sits between them, and the velocities come back once, after the last step.
void B200IntegrateRigidBodyStepKernel::executeSteps(ContextImpl& contextRef, const RigidBodyIntegrator& integrator, int steps) {
    if (steps <= 0) return;
    if (system == NULL) throw OpenMMException("B200 rigid-body kernel: positions have not been set");
    vector<Vec3>& R = *data.positions;
    vector<Vec3>& V = *data.velocities;
    vector<Vec3>& F = *data.forces;
    const double dt = integrator.getStepSize();
    context = &contextRef;
    tolerance = integrator.getConstraintTolerance();
    const bool freeAtoms = bodies != NULL && bodies->getNumFree() != 0;
    check(rbk_execute_host_hooks(system, dt, steps, &R[0][0], &V[0][0], &F[0][0], &evaluateForces,
                                 positionHook ? &constrainPositions : NULL,
                                 velocityHook && freeAtoms ? &constrainVelocities : NULL, this, NULL));
    data.time += dt*steps;
    data.stepCount += steps;
}

double B200IntegrateRigidBodyStepKernel::computeKineticEnergy(ContextImpl&, const RigidBodyIntegrator& integrator) {
    vector<double> ke = getKineticEnergies(integrator);
    return ke[0] + ke[1];
}

vector<double> B200IntegrateRigidBodyStepKernel::getKineticEnergies(const RigidBodyIntegrator&) {
    vector<double> ke(2, 0.0);
    if (system == NULL) return ke;
    check(rbk_kinetic_host(system, &(*data.velocities)[0][0], ke.data(), NULL));
    if (bodies != NULL) bodies->setKineticEnergies(ke[0], ke[1]);
    return ke;
}

// Refined ("shadow") energies: the diagnostics the reference has on its CUDA platform only
// (CudaRigidBodyKernels.cpp:469-494); its Reference platform returns the plain energies / 0
// (ReferenceRigidBodyKernels.cpp:123-132), which is also what is returned here until they are switched on
// with RigidBodyIntegrator::setComputeRefinedEnergies(true) and a step has been taken.
vector<double> B200IntegrateRigidBodyStepKernel::getRefinedKineticEnergies(const RigidBodyIntegrator& integrator) {
    if (!integrator.getComputeRefinedEnergies() || system == NULL || data.stepCount == 0) return getKineticEnergies(integrator);
    vector<double> ke(2, 0.0);
    check(rbk_refined_kinetic_host(system, integrator.getStepSize(), &(*data.velocities)[0][0], ke.data(), NULL));
    return ke;
}

double B200IntegrateRigidBodyStepKernel::getPotentialEnergyRefinement(const RigidBodyIntegrator& integrator) {
    if (!integrator.getComputeRefinedEnergies() || system == NULL || data.stepCount == 0) return 0.0;
    double out[2] = {0.0, 0.0};
    check(rbk_potential_refinement_host(system, integrator.getStepSize(), &(*data.forces)[0][0], out, NULL));
    return out[0];
}
