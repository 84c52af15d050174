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
};

} // namespace RigidBodyPlugin
#endif
