#ifndef OPENMM_HARMONICBONDFORCE_H_
#define OPENMM_HARMONICBONDFORCE_H_
// shim, see Vec3.h
#include "Force.h"
#include <cmath>
namespace OpenMM {
class HarmonicBondForce : public Force {
public:
    int addBond(int particle1, int particle2, double length, double k) {
        Bond b = {particle1, particle2, length, k};
        bonds.push_back(b);
        return (int) bonds.size() - 1;
    }
    double calcForcesAndEnergy(const std::vector<Vec3>& pos, std::vector<Vec3>& forces) const {
        double energy = 0.0;
        for (size_t i = 0; i < bonds.size(); i++) {
            const Bond& b = bonds[i];
            Vec3 d = pos[b.p2] - pos[b.p1];
            double r = std::sqrt(d.dot(d));
            energy += 0.5*b.k*(r - b.length)*(r - b.length);
            Vec3 f = d*(b.k*(r - b.length)/r);
            forces[b.p1] += f;
            forces[b.p2] -= f;
        }
        return energy;
    }
private:
    struct Bond { int p1, p2; double length, k; };
    std::vector<Bond> bonds;
};
}
#endif
