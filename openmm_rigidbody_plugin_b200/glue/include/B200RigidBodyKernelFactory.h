#ifndef RBK_GLUE_B200_KERNEL_FACTORY_H_
#define RBK_GLUE_B200_KERNEL_FACTORY_H_
// Kernel factory of the B200 implementation (counterpart of platforms/reference/include/ReferenceRigidBodyKernelFactory.h
// and platforms/cuda/include/CudaRigidBodyKernelFactory.h).
#include "openmm/KernelFactory.h"

namespace RigidBodyPlugin {

class B200RigidBodyKernelFactory : public OpenMM::KernelFactory {
public:
    OpenMM::KernelImpl* createKernelImpl(std::string name, const OpenMM::Platform& platform, OpenMM::ContextImpl& context) const;
};

} // namespace RigidBodyPlugin

extern "C" void registerPlatforms();
extern "C" void registerKernelFactories();
extern "C" void registerRigidBodyB200KernelFactories();
extern "C" void registerRigidBodyCudaKernelFactories();      // the name the reference's CUDA plugin exports
#endif
