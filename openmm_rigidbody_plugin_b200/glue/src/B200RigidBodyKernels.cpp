// B200 kernel: every IntegrateRigidBodyStepKernel method forwards to the C ABI (include/rbk.h).
// Step protocol = ReferenceIntegrateRigidBodyStepKernel::execute (platforms/reference/src/ReferenceRigidBodyKernels.cpp:82-108)
// for systems without free-atom constraints / virtual sites: Part 1 -> forces at the new positions -> Part 2.
#include "B200RigidBodyKernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"

using namespace RigidBodyPlugin;
using namespace OpenMM;
using std::vector;

static void check(int rc) {
    if (rc != RBK_OK) throw OpenMMException(rbk_last_error());
}

void B200IntegrateRigidBodyStepKernel::initialize(ContextImpl& contextRef, const RigidBodyIntegrator& integrator) {
    context = &contextRef;
    if (contextRef.getSystem().getNumConstraints() != 0 && integrator.getRigidBodySystem().getNumFree() != 0)
        throw OpenMMException("B200 rigid-body kernel: constraints between free atoms need OpenMM's constraint kernels, "
                              "which are outside this implementation");
}

void B200IntegrateRigidBodyStepKernel::uploadBodySystem(RigidBodySystem& bodySystem) {
    bodies = &bodySystem;
    system = bodySystem.getHandle();
    check(rbk_upload(system, NULL));
}

void B200IntegrateRigidBodyStepKernel::evaluateForces(const double*, double*, int, void* self) {
    // positions have been copied back into the platform data; let OpenMM fill data.forces
    static_cast<B200IntegrateRigidBodyStepKernel*>(self)->context->calcForcesAndEnergy(true, false);
}

void B200IntegrateRigidBodyStepKernel::execute(ContextImpl& contextRef, const RigidBodyIntegrator& integrator) {
    if (system == NULL) throw OpenMMException("B200 rigid-body kernel: positions have not been set");
    vector<Vec3>& R = *data.positions;
    vector<Vec3>& V = *data.velocities;
    vector<Vec3>& F = *data.forces;
    const double dt = integrator.getStepSize();
    context = &contextRef;
    check(rbk_execute_host(system, dt, 1, &R[0][0], &V[0][0], &F[0][0], &evaluateForces, this, NULL));
    data.time += dt;
    data.stepCount++;
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

// Refined ("shadow") energies are diagnostics of the reference's CUDA platform only; like its Reference platform
// (ReferenceRigidBodyKernels.cpp:123-132) this implementation reports the plain kinetic energies and no refinement.
vector<double> B200IntegrateRigidBodyStepKernel::getRefinedKineticEnergies(const RigidBodyIntegrator& integrator) {
    return getKineticEnergies(integrator);
}

double B200IntegrateRigidBodyStepKernel::getPotentialEnergyRefinement(const RigidBodyIntegrator&) { return 0.0; }
