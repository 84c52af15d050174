#ifndef OPENMM_CONTEXTIMPL_H_
#define OPENMM_CONTEXTIMPL_H_
// shim, see Vec3.h
#include "openmm/Integrator.h"
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/reference/ReferencePlatform.h"
#include "openmm/reference/ReferenceVirtualSites.h"
#include <vector>
namespace OpenMM {
class Context;
class ContextImpl {
public:
    ContextImpl(Context& owner, const System& system, Integrator& integrator, Platform* platform, void* platformData)
        : owner(owner), system(system), integrator(integrator), platform(platform), platformData(platformData), lastEnergy(0.0) {}
    Context& getOwner() { return owner; }
    const System& getSystem() const { return system; }
    Integrator& getIntegrator() { return integrator; }
    Platform& getPlatform() { return *platform; }
    void* getPlatformData() { return platformData; }
    void getPositions(std::vector<Vec3>& positions);
    void getVelocities(std::vector<Vec3>& velocities);
    void getForces(std::vector<Vec3>& forces);
    bool updateContextState() { return false; }
    double calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups = 0xFFFFFFFF);
    double getLastEnergy() const { return lastEnergy; }
private:
    Context& owner;
    const System& system;
    Integrator& integrator;
    Platform* platform;
    void* platformData;
    double lastEnergy;
};

inline void ContextImpl::getPositions(std::vector<Vec3>& positions) { positions = *((ReferencePlatform::PlatformData*) platformData)->positions; }
inline void ContextImpl::getVelocities(std::vector<Vec3>& velocities) { velocities = *((ReferencePlatform::PlatformData*) platformData)->velocities; }
inline void ContextImpl::getForces(std::vector<Vec3>& forces) { forces = *((ReferencePlatform::PlatformData*) platformData)->forces; }
inline double ContextImpl::calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups) {
    ReferencePlatform::PlatformData* d = (ReferencePlatform::PlatformData*) platformData;
    std::vector<Vec3> scratch(d->numParticles);
    std::vector<Vec3>& f = includeForces ? *d->forces : scratch;
    for (size_t i = 0; i < f.size(); i++) f[i] = Vec3();
    double energy = 0.0;
    for (int i = 0; i < system.getNumForces(); i++) energy += system.getForce(i).calcForcesAndEnergy(*d->positions, f);
    if (includeForces) ReferenceVirtualSites::distributeForces(system, *d->positions, f);
    lastEnergy = energy;
    return energy;
}
}
#endif
