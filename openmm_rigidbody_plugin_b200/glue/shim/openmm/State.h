#ifndef OPENMM_STATE_H_
#define OPENMM_STATE_H_
// shim, see Vec3.h
#include "Vec3.h"
#include <vector>
namespace OpenMM {
class State {
public:
    enum DataType { Positions = 1, Velocities = 2, Forces = 4, Energy = 8, Parameters = 16, ParameterDerivatives = 32 };
    State() : time(0.0), ke(0.0), pe(0.0) {}
    double getTime() const { return time; }
    const std::vector<Vec3>& getPositions() const { return positions; }
    const std::vector<Vec3>& getVelocities() const { return velocities; }
    const std::vector<Vec3>& getForces() const { return forces; }
    double getKineticEnergy() const { return ke; }
    double getPotentialEnergy() const { return pe; }
private:
    friend class Context;
    double time, ke, pe;
    std::vector<Vec3> positions, velocities, forces;
};
}
#endif
