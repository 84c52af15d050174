#ifndef OPENMM_CONTEXTIMPL_H_
#define OPENMM_CONTEXTIMPL_H_
// shim, see Vec3.h
#include "openmm/Integrator.h"
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/cuda/CudaPlatform.h"
#include "openmm/reference/ReferencePlatform.h"
#include "openmm/reference/ReferenceVirtualSites.h"
#include <vector>
namespace OpenMM {
class Context;
class ContextImpl {
public:
    ContextImpl(Context& owner, const System& system, Integrator& integrator, Platform* platform, void* platformData, bool cuda = false)
        : owner(owner), system(system), integrator(integrator), platform(platform), platformData(platformData), lastEnergy(0.0), cuda(cuda),
          numStateUpdates(0) {}
    Context& getOwner() { return owner; }
    const System& getSystem() const { return system; }
    Integrator& getIntegrator() { return integrator; }
    Platform& getPlatform() { return *platform; }
    void* getPlatformData() { return platformData; }
    void getPositions(std::vector<Vec3>& positions);
    void getVelocities(std::vector<Vec3>& velocities);
    void getForces(std::vector<Vec3>& forces);
    bool updateContextState() {
        numStateUpdates++;
        for (int i = 0; i < system.getNumForces(); i++) system.getForce(i).updateContextState(*this);
        return false;
    }
    int getNumStateUpdates() const { return numStateUpdates; }      // shim-only: how often an integrator asked
    double calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups = 0xFFFFFFFF);
    double getLastEnergy() const { return lastEnergy; }
private:
    ReferencePlatform::PlatformData* ref() { return (ReferencePlatform::PlatformData*) platformData; }
    CudaContext& cu() { return *((CudaPlatform::PlatformData*) platformData)->contexts[0]; }
    Context& owner;
    const System& system;
    Integrator& integrator;
    Platform* platform;
    void* platformData;
    double lastEnergy;
    bool cuda;
    int numStateUpdates;
};

inline void ContextImpl::getPositions(std::vector<Vec3>& positions) { if (cuda) cu().downloadPositions(positions); else positions = *ref()->positions; }
inline void ContextImpl::getVelocities(std::vector<Vec3>& velocities) { if (cuda) cu().downloadVelocities(velocities); else velocities = *ref()->velocities; }
inline void ContextImpl::getForces(std::vector<Vec3>& forces) { if (cuda) cu().downloadForces(forces); else forces = *ref()->forces; }
// Reference platform: forces are evaluated in place on the host data.  CUDA platform (shim): the real platform runs its
// force kernels on the device arrays; here the positions come down, the shim Forces are evaluated on the host and the
// result goes up into the fixed-point force array - a stand-in for the force field, not part of the integrator.
// A System without Forces leaves the device force array as it is (integrator-only runs with prescribed forces).
inline double ContextImpl::calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups) {
    const int n = system.getNumParticles();
    if (cuda && system.getNumForces() == 0) return lastEnergy = 0.0;
    std::vector<Vec3> scratch(n), downloaded;
    if (cuda) cu().downloadPositions(downloaded);
    const std::vector<Vec3>& pos = cuda ? downloaded : *ref()->positions;
    std::vector<Vec3>& f = includeForces && !cuda ? *ref()->forces : scratch;
    for (size_t i = 0; i < f.size(); i++) f[i] = Vec3();
    double energy = 0.0;
    for (int i = 0; i < system.getNumForces(); i++) energy += system.getForce(i).calcForcesAndEnergy(pos, f);
    if (includeForces) {
        ReferenceVirtualSites::distributeForces(system, pos, f);
        if (cuda) cu().uploadForces(f);
    }
    lastEnergy = energy;
    return energy;
}
}
#endif
