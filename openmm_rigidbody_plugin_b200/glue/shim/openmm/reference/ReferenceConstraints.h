#ifndef OPENMM_REFERENCECONSTRAINTS_H_
#define OPENMM_REFERENCECONSTRAINTS_H_
// shim, see Vec3.h: OpenMM's ReferenceConstraints picks SETTLE/CCMA; this stand-in with the same interface runs plain
// SHAKE / RATTLE sweeps over the System's distance constraints until every one is within `tolerance`.
#include "openmm/OpenMMException.h"
#include "openmm/System.h"
#include "openmm/Vec3.h"
#include <cmath>
#include <vector>
namespace OpenMM {
class ReferenceConstraints {
public:
    ReferenceConstraints(const System& system) {
        for (int i = 0; i < system.getNumConstraints(); i++) {
            int a, b; double d;
            system.getConstraintParameters(i, a, b, d);
            atom1.push_back(a); atom2.push_back(b); distance.push_back(d);
        }
    }
    // atomCoordinates: positions before the step; atomCoordinatesP: unconstrained new positions, corrected in place
    void apply(std::vector<Vec3>& atomCoordinates, std::vector<Vec3>& atomCoordinatesP, std::vector<double>& inverseMasses, double tolerance) {
        for (int sweep = 0; sweep < 500; sweep++) {
            bool done = true;
            for (size_t k = 0; k < atom1.size(); k++) {
                const int a = atom1[k], b = atom2[k];
                const Vec3 r0 = atomCoordinates[a] - atomCoordinates[b], r = atomCoordinatesP[a] - atomCoordinatesP[b];
                const double d2 = distance[k]*distance[k], diff = d2 - r.dot(r);
                if (std::fabs(diff) <= 2.0*tolerance*d2) continue;
                done = false;
                const double g = diff/(2.0*(inverseMasses[a] + inverseMasses[b])*r0.dot(r));
                atomCoordinatesP[a] += r0*(g*inverseMasses[a]);
                atomCoordinatesP[b] -= r0*(g*inverseMasses[b]);
            }
            if (done) return;
        }
        throw OpenMMException("ReferenceConstraints (shim): SHAKE did not converge");
    }
    void applyToVelocities(std::vector<Vec3>& atomCoordinates, std::vector<Vec3>& velocities, std::vector<double>& inverseMasses, double tolerance) {
        for (int sweep = 0; sweep < 500; sweep++) {
            bool done = true;
            for (size_t k = 0; k < atom1.size(); k++) {
                const int a = atom1[k], b = atom2[k];
                const Vec3 r = atomCoordinates[a] - atomCoordinates[b], v = velocities[a] - velocities[b];
                const double d2 = distance[k]*distance[k], rv = r.dot(v);
                if (std::fabs(rv) <= tolerance*d2) continue;
                done = false;
                const double g = rv/((inverseMasses[a] + inverseMasses[b])*d2);
                velocities[a] -= r*(g*inverseMasses[a]);
                velocities[b] += r*(g*inverseMasses[b]);
            }
            if (done) return;
        }
        throw OpenMMException("ReferenceConstraints (shim): RATTLE did not converge");
    }
private:
    std::vector<int> atom1, atom2;
    std::vector<double> distance;
};
}
#endif
