#ifndef RBK_GLUE_B200_KERNELS_H_
#define RBK_GLUE_B200_KERNELS_H_
// IntegrateRigidBodyStepKernel implemented on librbk's C ABI, operating on the Reference platform's host data
// (std::vector<Vec3> positions / velocities / forces): the counterpart of
// platforms/reference/src/ReferenceRigidBodyKernels.h, with the arithmetic on the GPU.
#include "RigidBodyKernels.h"
#include "openmm/reference/ReferencePlatform.h"
#include "rbk.h"

namespace RigidBodyPlugin {

class B200IntegrateRigidBodyStepKernel : public IntegrateRigidBodyStepKernel {
public:
    B200IntegrateRigidBodyStepKernel(std::string name, const OpenMM::Platform& platform, OpenMM::ReferencePlatform::PlatformData& data)
        : IntegrateRigidBodyStepKernel(name, platform), data(data), system(NULL), context(NULL), bodies(NULL), tolerance(1e-5),
          positionHook(false), velocityHook(false), refined(false) {}
    void initialize(OpenMM::ContextImpl& context, const RigidBodyIntegrator& integrator);
    void uploadBodySystem(RigidBodySystem& bodySystem);
    void execute(OpenMM::ContextImpl& context, const RigidBodyIntegrator& integrator);
    void executeSteps(OpenMM::ContextImpl& context, const RigidBodyIntegrator& integrator, int steps);
    double computeKineticEnergy(OpenMM::ContextImpl& context, const RigidBodyIntegrator& integrator);
    std::vector<double> getKineticEnergies(const RigidBodyIntegrator& integrator);
    std::vector<double> getRefinedKineticEnergies(const RigidBodyIntegrator& integrator);
    double getPotentialEnergyRefinement(const RigidBodyIntegrator& integrator);
private:
    static void evaluateForces(const double* R, double* F, int numAtoms, void* self);
    static int constrainPositions(const double* oldR, double* R, int numAtoms, void* self);
    static int constrainVelocities(const double* R, double* V, int numAtoms, void* self);
    OpenMM::ReferencePlatform::PlatformData& data;
    rbk_system* system;                      // borrowed from the integrator's RigidBodySystem
    OpenMM::ContextImpl* context;
    RigidBodySystem* bodies;
    std::vector<double> invMass;
    std::vector<OpenMM::Vec3> oldPos;
    double tolerance;
    bool positionHook, velocityHook, refined;
};

} // namespace RigidBodyPlugin
#endif
