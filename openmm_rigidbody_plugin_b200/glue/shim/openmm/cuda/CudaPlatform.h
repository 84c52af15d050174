#ifndef OPENMM_CUDAPLATFORM_H_
#define OPENMM_CUDAPLATFORM_H_
// shim, see ../Vec3.h: Platform "CUDA" with its PlatformData (one CudaContext per device; the plugin uses contexts[0],
// platforms/cuda/src/CudaRigidBodyKernelFactory.cpp:67-72 of the reference) and the CudaPrecision property.
#include "openmm/Platform.h"
#include "openmm/cuda/CudaContext.h"
namespace OpenMM {
class CudaPlatform : public Platform {
public:
    class PlatformData {
    public:
        PlatformData(const System& system, const std::string& precision) { contexts.push_back(new CudaContext(system, precision)); }
        ~PlatformData() { for (size_t i = 0; i < contexts.size(); i++) delete contexts[i]; }
        void initializeContexts(const System&) {}
        std::vector<CudaContext*> contexts;
    };
    CudaPlatform() : name("CUDA") { setPropertyDefaultValue(CudaPrecision(), "single"); }
    const std::string& getName() const { return name; }
    static const std::string& CudaPrecision() { static const std::string key = "CudaPrecision"; return key; }
private:
    std::string name;
};
}
#endif
