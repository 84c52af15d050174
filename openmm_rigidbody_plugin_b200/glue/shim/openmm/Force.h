#ifndef OPENMM_FORCE_H_
#define OPENMM_FORCE_H_
// shim, see Vec3.h.  Real OpenMM separates Force / ForceImpl / platform kernels; for exercising the integrator
// glue a Force only needs to add its forces and return its energy.
#include "Vec3.h"
#include <vector>
namespace OpenMM {
class ContextImpl;
class Force {
public:
    virtual ~Force() {}
    virtual double calcForcesAndEnergy(const std::vector<Vec3>& positions, std::vector<Vec3>& forces) const = 0;
    // ForceImpl::updateContextState: thermostats, barostats, CMMotionRemover change the state here, once per step
    virtual void updateContextState(ContextImpl& context) const {}
};
}
#endif
