// RigidBodyIntegrator: binds to a Context, owns the body system and one kernel, loops execute() in step().
// Behaviour follows openmmapi/src/RigidBodyIntegrator.cpp (constructor defaults, exception messages, the
// stateChanged protocol); the arithmetic is entirely behind IntegrateRigidBodyStepKernel.
#include "RigidBodyIntegrator.h"
#include "RigidBodyKernels.h"
#include "openmm/OpenMMException.h"

using namespace RigidBodyPlugin;
using namespace OpenMM;
using std::string;
using std::vector;

RigidBodyIntegrator::RigidBodyIntegrator(double stepSize, const vector<int>& bodyIndices)
    : bodyIndices(bodyIndices), rotationMode(0), computeRefinedEnergies(false) {
    setStepSize(stepSize);
    setConstraintTolerance(1e-5);
}

void RigidBodyIntegrator::setRotationMode(int mode) {
    if (mode < 0) throw OpenMMException("Rotation mode cannot be negative");
    if (owner != NULL) throw OpenMMException("Cannot set rotation mode: integrator already bound to a context");
    rotationMode = mode;
}

void RigidBodyIntegrator::setComputeRefinedEnergies(bool compute) {
    if (owner != NULL) throw OpenMMException("Cannot set refined energy computation: integrator already bound to a context");
    computeRefinedEnergies = compute;
}

void RigidBodyIntegrator::initialize(ContextImpl& contextRef) {
    if (owner != NULL && &contextRef.getOwner() != owner) throw OpenMMException("This Integrator is already bound to a context");
    if (contextRef.getSystem().getNumParticles() != (int) bodyIndices.size())
        throw OpenMMException("Number of body indices differs from that of atoms in Context");
    context = &contextRef;
    owner = &contextRef.getOwner();
    bodySystem.initialize(contextRef, bodyIndices, rotationMode);
    kernel = context->getPlatform().createKernel(IntegrateRigidBodyStepKernel::Name(), contextRef);
    kernel.getAs<IntegrateRigidBodyStepKernel>().initialize(contextRef, *this);
}

void RigidBodyIntegrator::cleanup() { kernel = Kernel(); }

vector<string> RigidBodyIntegrator::getKernelNames() { return vector<string>(1, IntegrateRigidBodyStepKernel::Name()); }

void RigidBodyIntegrator::stateChanged(State::DataType changed) {
    if (changed != State::Positions && changed != State::Velocities) return;
    const bool positions = changed == State::Positions;
    if (positions) {
        context->updateContextState();
        context->calcForcesAndEnergy(true, false);           // the body build needs F and tau at the new positions
    }
    IntegrateRigidBodyStepKernel& impl = kernel.getAs<IntegrateRigidBodyStepKernel>();
    if (impl.updateBodySystem(*context, bodySystem, positions, true)) return;      // built where the kernel keeps the bodies
    bodySystem.update(*context, positions, true);
    impl.uploadBodySystem(bodySystem);
}

double RigidBodyIntegrator::computeKineticEnergy() {
    return kernel.getAs<IntegrateRigidBodyStepKernel>().computeKineticEnergy(*context, *this);
}

vector<double> RigidBodyIntegrator::getKineticEnergies() {
    return kernel.getAs<IntegrateRigidBodyStepKernel>().getKineticEnergies(*this);
}

vector<double> RigidBodyIntegrator::getRefinedKineticEnergies() {
    return kernel.getAs<IntegrateRigidBodyStepKernel>().getRefinedKineticEnergies(*this);
}

double RigidBodyIntegrator::getPotentialEnergyRefinement() {
    return kernel.getAs<IntegrateRigidBodyStepKernel>().getPotentialEnergyRefinement(*this);
}

void RigidBodyIntegrator::step(int steps) {
    if (context == NULL) throw OpenMMException("This Integrator is not bound to a context!");
    kernel.getAs<IntegrateRigidBodyStepKernel>().executeSteps(*context, *this, steps);
}
