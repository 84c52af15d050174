#ifndef OPENMM_KERNEL_H_
#define OPENMM_KERNEL_H_
// shim, see Vec3.h: reference-counted handle to a KernelImpl, like OpenMM::Kernel
#include "KernelImpl.h"
namespace OpenMM {
class Kernel {
public:
    Kernel() : impl(0) {}
    Kernel(KernelImpl* impl) : impl(impl) {}
    Kernel(const Kernel& copy) : impl(copy.impl) { if (impl) impl->referenceCount++; }
    ~Kernel() { release(); }
    Kernel& operator=(const Kernel& copy) {
        if (copy.impl) copy.impl->referenceCount++;
        release();
        impl = copy.impl;
        return *this;
    }
    std::string getName() const { return impl->getName(); }
    const KernelImpl& getImpl() const { return *impl; }
    KernelImpl& getImpl() { return *impl; }
    template <class T> T& getAs() { return dynamic_cast<T&>(*impl); }
    template <class T> const T& getAs() const { return dynamic_cast<const T&>(*impl); }
private:
    void release() { if (impl && --impl->referenceCount == 0) delete impl; impl = 0; }
    KernelImpl* impl;
};
}
#endif
