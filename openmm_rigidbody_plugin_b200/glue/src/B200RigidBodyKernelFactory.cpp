// Plugin entry points and kernel factory (same three extern "C" symbols every OpenMM platform plugin exports;
// reference: platforms/reference/src/ReferenceRigidBodyKernelFactory.cpp:41-63).
#include "B200RigidBodyKernelFactory.h"
#include "B200RigidBodyKernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/ReferencePlatform.h"

using namespace RigidBodyPlugin;
using namespace OpenMM;

extern "C" OPENMM_EXPORT void registerPlatforms() {}

extern "C" OPENMM_EXPORT void registerKernelFactories() {
    for (int i = 0; i < Platform::getNumPlatforms(); i++) {
        Platform& platform = Platform::getPlatform(i);
        if (dynamic_cast<ReferencePlatform*>(&platform) != NULL)
            platform.registerKernelFactory(IntegrateRigidBodyStepKernel::Name(), new B200RigidBodyKernelFactory());
    }
}

extern "C" OPENMM_EXPORT void registerRigidBodyB200KernelFactories() { registerKernelFactories(); }

KernelImpl* B200RigidBodyKernelFactory::createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
    ReferencePlatform::PlatformData& data = *static_cast<ReferencePlatform::PlatformData*>(context.getPlatformData());
    if (name == IntegrateRigidBodyStepKernel::Name()) return new B200IntegrateRigidBodyStepKernel(name, platform, data);
    throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '") + name + "'").c_str());
}
