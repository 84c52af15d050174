/* Minimal stand-in for openmm/internal/ContextImpl.h: a bag holding a System and the
 * R, V, F vectors that RigidBodySystem::update pulls out of the context.  Test infrastructure. */
#ifndef RBK_SHIM_OPENMM_CONTEXTIMPL_H_
#define RBK_SHIM_OPENMM_CONTEXTIMPL_H_
#include "openmm/System.h"
#include "openmm/Vec3.h"
#include <vector>
namespace OpenMM {
class ContextImpl {
public:
    explicit ContextImpl(System& system) : system(system) {}
    const System& getSystem() const { return system; }
    void getPositions(std::vector<Vec3>& out) const { out = R; }
    void getVelocities(std::vector<Vec3>& out) const { out = V; }
    void getForces(std::vector<Vec3>& out) const { out = F; }
    std::vector<Vec3> R, V, F;
private:
    System& system;
};
}
#endif
