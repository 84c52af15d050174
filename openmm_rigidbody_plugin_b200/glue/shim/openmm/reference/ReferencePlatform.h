#ifndef OPENMM_REFERENCEPLATFORM_H_
#define OPENMM_REFERENCEPLATFORM_H_
// shim, see Vec3.h: the Reference platform keeps positions/velocities/forces as std::vector<Vec3> in its PlatformData
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/reference/ReferenceConstraints.h"
namespace OpenMM {
class ReferencePlatform : public Platform {
public:
    class PlatformData {
    public:
        PlatformData(const System& system) : time(0.0), stepCount(0), numParticles(system.getNumParticles()) {
            positions = new std::vector<Vec3>(numParticles);
            velocities = new std::vector<Vec3>(numParticles);
            forces = new std::vector<Vec3>(numParticles);
            constraints = new ReferenceConstraints(system);
        }
        ~PlatformData() { delete positions; delete velocities; delete forces; delete constraints; }
        double time;
        int stepCount, numParticles;
        std::vector<Vec3>* positions;
        std::vector<Vec3>* velocities;
        std::vector<Vec3>* forces;
        ReferenceConstraints* constraints;
    };
    ReferencePlatform() : name("Reference") {}
    const std::string& getName() const { return name; }
private:
    std::string name;
};
}
#endif
