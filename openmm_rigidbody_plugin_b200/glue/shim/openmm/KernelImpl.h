#ifndef OPENMM_KERNELIMPL_H_
#define OPENMM_KERNELIMPL_H_
// shim, see Vec3.h
#include <string>
namespace OpenMM {
class Platform;
class KernelImpl {
public:
    KernelImpl(std::string name, const Platform& platform) : name(name), platform(&platform), referenceCount(1) {}
    virtual ~KernelImpl() {}
    std::string getName() const { return name; }
    const Platform& getPlatform() { return *platform; }
private:
    friend class Kernel;
    std::string name;
    const Platform* platform;
    int referenceCount;
};
}
#endif
