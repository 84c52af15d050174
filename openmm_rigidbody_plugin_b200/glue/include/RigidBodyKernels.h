#ifndef RBK_GLUE_RIGIDBODYKERNELS_H_
#define RBK_GLUE_RIGIDBODYKERNELS_H_
// The kernel interface the integrator dispatches through - the drop-in boundary of the reference
// (openmmapi/include/RigidBodyKernels.h:47-99): same kernel name, same seven pure virtuals.
#include "RigidBodyIntegrator.h"
#include "openmm/KernelImpl.h"
#include "openmm/Platform.h"
#include <string>
#include <vector>

namespace RigidBodyPlugin {

class IntegrateRigidBodyStepKernel : public OpenMM::KernelImpl {
public:
    static std::string Name() { return "IntegrateRigidBodyStep"; }
    IntegrateRigidBodyStepKernel(std::string name, const OpenMM::Platform& platform) : OpenMM::KernelImpl(name, platform) {}
    virtual void initialize(OpenMM::ContextImpl& context, const RigidBodyIntegrator& integrator) = 0;
    virtual void uploadBodySystem(RigidBodySystem& bodySystem) = 0;
    virtual void execute(OpenMM::ContextImpl& context, const RigidBodyIntegrator& integrator) = 0;
    virtual double computeKineticEnergy(OpenMM::ContextImpl& context, const RigidBodyIntegrator& integrator) = 0;
    virtual std::vector<double> getKineticEnergies(const RigidBodyIntegrator& integrator) = 0;
    virtual std::vector<double> getRefinedKineticEnergies(const RigidBodyIntegrator& integrator) = 0;
    virtual double getPotentialEnergyRefinement(const RigidBodyIntegrator& integrator) = 0;
    // ---- two optional hooks on top of the reference's interface (defaults = the reference's behaviour) -------------------
    // RigidBodyIntegrator::step(n) (openmmapi/src/RigidBodyIntegrator.cpp:96-101) as ONE call, so that an implementation
    // may fuse Part 2 of a step with Part 1 of the next and return host data once per call.
    virtual void executeSteps(OpenMM::ContextImpl& context, const RigidBodyIntegrator& integrator, int steps) {
        for (int i = 0; i < steps; ++i) execute(context, integrator);
    }
    // RigidBodyIntegrator::stateChanged (:63-74): rebuild the bodies where the implementation keeps them.  Return false
    // to get the reference's path: RigidBodySystem::update on the host followed by uploadBodySystem.
    virtual bool updateBodySystem(OpenMM::ContextImpl& context, RigidBodySystem& bodySystem, bool geometry, bool velocities) {
        return false;
    }
};

} // namespace RigidBodyPlugin
#endif
