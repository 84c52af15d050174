#ifndef RBK_GLUE_B200_CUDA_KERNELS_H_
#define RBK_GLUE_B200_CUDA_KERNELS_H_
// IntegrateRigidBodyStepKernel on OpenMM's CUDA platform, implemented on librbk's C ABI: the counterpart of the
// reference's CudaIntegrateRigidBodyStepKernel (platforms/cuda/src/CudaRigidBodyKernels.h:45-137).  Positions,
// velocities and forces are BORROWED from the CudaContext (posq [+ posqCorrection], velm, force) and never leave the
// device: every step is rbk_*_openmm calls on those arrays.
#include "RigidBodyKernels.h"
#include "openmm/cuda/CudaContext.h"
#include "rbk.h"
#include <vector>

namespace RigidBodyPlugin {

class B200CudaIntegrateRigidBodyStepKernel : public IntegrateRigidBodyStepKernel {
public:
    B200CudaIntegrateRigidBodyStepKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu)
        : IntegrateRigidBodyStepKernel(name, platform), cu(cu), system(NULL), bodies(NULL), precision(RBK_OPENMM_SINGLE),
          numFree(0), numBodies(0), refined(false), constrained(false), statelessForces(false), stepsTaken(0) {}
    void initialize(OpenMM::ContextImpl& context, const RigidBodyIntegrator& integrator);
    void uploadBodySystem(RigidBodySystem& bodySystem);
    bool updateBodySystem(OpenMM::ContextImpl& context, RigidBodySystem& bodySystem, bool geometry, bool velocities);
    void execute(OpenMM::ContextImpl& context, const RigidBodyIntegrator& integrator);
    void executeSteps(OpenMM::ContextImpl& context, const RigidBodyIntegrator& integrator, int steps);
    double computeKineticEnergy(OpenMM::ContextImpl& context, const RigidBodyIntegrator& integrator);
    std::vector<double> getKineticEnergies(const RigidBodyIntegrator& integrator);
    std::vector<double> getRefinedKineticEnergies(const RigidBodyIntegrator& integrator);
    double getPotentialEnergyRefinement(const RigidBodyIntegrator& integrator);
    void atomsReordered();                          // called by the reorder listener
private:
    class ReorderListener;
    void* posq() { return (void*) cu.getPosq().getDevicePointer(); }
    void* posqCorrection() { return cu.getUseMixedPrecision() ? (void*) cu.getPosqCorrection().getDevicePointer() : NULL; }
    void* velm() { return (void*) cu.getVelm().getDevicePointer(); }
    long long* force() { return (long long*) cu.getForce().getDevicePointer(); }
    void* stream() { return (void*) cu.getCurrentStream(); }
    std::vector<int> currentLocation() const;
    void firstHalf(const RigidBodyIntegrator& integrator);
    void afterSecondHalf(const RigidBodyIntegrator& integrator);
    void endOfStep(const RigidBodyIntegrator& integrator);
    OpenMM::CudaContext& cu;
    rbk_system* system;                             // borrowed from the integrator's RigidBodySystem
    RigidBodySystem* bodies;
    std::vector<int> atomIndex;                     // plugin order -> particle (RigidBodySystem::getAtomIndex, cached)
    int precision, numFree, numBodies;
    bool refined, constrained, statelessForces;
    long long stepsTaken;
};

} // namespace RigidBodyPlugin
#endif
