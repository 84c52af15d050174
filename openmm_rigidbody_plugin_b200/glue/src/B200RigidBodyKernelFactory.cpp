// Plugin entry points and kernel factory (the same extern "C" symbols every OpenMM platform plugin exports; reference:
// platforms/cuda/src/CudaRigidBodyKernelFactory.cpp:43-72 and platforms/reference/src/ReferenceRigidBodyKernelFactory.cpp:41-63).
// One library serves both platforms: on "CUDA" the kernel works on the CudaContext's device arrays
// (B200CudaIntegrateRigidBodyStepKernel), on "Reference" on the platform's host vectors (B200IntegrateRigidBodyStepKernel).
#include "B200RigidBodyKernelFactory.h"
#include "B200CudaRigidBodyKernels.h"
#include "B200RigidBodyKernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/cuda/CudaPlatform.h"
#include "openmm/reference/ReferencePlatform.h"

using namespace RigidBodyPlugin;
using namespace OpenMM;

extern "C" OPENMM_EXPORT void registerPlatforms() {}

extern "C" OPENMM_EXPORT void registerKernelFactories() {
    for (int i = 0; i < Platform::getNumPlatforms(); i++) {
        Platform& platform = Platform::getPlatform(i);
        if (platform.getName() == "CUDA" || dynamic_cast<ReferencePlatform*>(&platform) != NULL)
            platform.registerKernelFactory(IntegrateRigidBodyStepKernel::Name(), new B200RigidBodyKernelFactory());
    }
}

// for programs that link the plugin instead of loading it from lib/plugins (the reference's tests do:
// registerRigidBodyCudaKernelFactories, CudaRigidBodyKernelFactory.cpp:56-65)
extern "C" OPENMM_EXPORT void registerRigidBodyB200KernelFactories() {
    try {
        Platform::getPlatformByName("CUDA");
    }
    catch (...) {
        Platform::registerPlatform(new CudaPlatform());
    }
    registerKernelFactories();
}

extern "C" OPENMM_EXPORT void registerRigidBodyCudaKernelFactories() { registerRigidBodyB200KernelFactories(); }

KernelImpl* B200RigidBodyKernelFactory::createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
    if (name != IntegrateRigidBodyStepKernel::Name())
        throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '") + name + "'").c_str());
    if (platform.getName() == "CUDA") {
        CudaContext& cu = *static_cast<CudaPlatform::PlatformData*>(context.getPlatformData())->contexts[0];
        return new B200CudaIntegrateRigidBodyStepKernel(name, platform, cu);
    }
    ReferencePlatform::PlatformData& data = *static_cast<ReferencePlatform::PlatformData*>(context.getPlatformData());
    return new B200IntegrateRigidBodyStepKernel(name, platform, data);
}
