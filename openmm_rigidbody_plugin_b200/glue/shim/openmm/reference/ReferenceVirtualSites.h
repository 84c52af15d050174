#ifndef OPENMM_REFERENCEVIRTUALSITES_H_
#define OPENMM_REFERENCEVIRTUALSITES_H_
// shim, see Vec3.h
#include "openmm/System.h"
#include "openmm/VirtualSite.h"
#include "openmm/Vec3.h"
#include <vector>
namespace OpenMM {
class ReferenceVirtualSites {
public:
    static void computePositions(const System& system, std::vector<Vec3>& atomCoordinates) {
        for (int i = 0; i < system.getNumParticles(); i++) {
            if (!system.isVirtualSite(i)) continue;
            const TwoParticleAverageSite* site = dynamic_cast<const TwoParticleAverageSite*>(system.getVirtualSite(i));
            if (site != NULL)
                atomCoordinates[i] = atomCoordinates[site->getParticle(0)]*site->getWeight(0) + atomCoordinates[site->getParticle(1)]*site->getWeight(1);
        }
    }
    // forces that act on a virtual site are handed to the particles that define it
    static void distributeForces(const System& system, const std::vector<Vec3>&, std::vector<Vec3>& forces) {
        for (int i = 0; i < system.getNumParticles(); i++) {
            if (!system.isVirtualSite(i)) continue;
            const TwoParticleAverageSite* site = dynamic_cast<const TwoParticleAverageSite*>(system.getVirtualSite(i));
            if (site == NULL) continue;
            forces[site->getParticle(0)] += forces[i]*site->getWeight(0);
            forces[site->getParticle(1)] += forces[i]*site->getWeight(1);
            forces[i] = Vec3();
        }
    }
};
}
#endif
