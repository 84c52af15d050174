#ifndef OPENMM_CUSTOMEXTERNALFORCE_H_
#define OPENMM_CUSTOMEXTERNALFORCE_H_
// shim, see Vec3.h.  OpenMM's CustomExternalForce evaluates a user expression per particle; this stand-in of the same
// name (and the same role: a stateless per-particle potential) has ONE fixed form, the analytic test potential of the
// parity tests: E_i = k/2 |x_i - x0_i|^2 - q_i (field . x_i)  (harmonic tether to a site + a charge in a uniform field).
#include "Force.h"
namespace OpenMM {
class CustomExternalForce : public Force {
public:
    CustomExternalForce(double k, const Vec3& field) : k(k), field(field) {}
    int addParticle(int particle, const Vec3& site, double charge) {
        Term t = {particle, site, charge};
        terms.push_back(t);
        return (int) terms.size() - 1;
    }
    double calcForcesAndEnergy(const std::vector<Vec3>& pos, std::vector<Vec3>& forces) const {
        double energy = 0.0;
        for (size_t i = 0; i < terms.size(); i++) {
            const Term& t = terms[i];
            const Vec3 dx = pos[t.particle] - t.site;
            forces[t.particle] += dx*(-k) + field*t.charge;
            energy += 0.5*k*dx.dot(dx) - t.charge*pos[t.particle].dot(field);
        }
        return energy;
    }
private:
    struct Term { int particle; Vec3 site; double charge; };
    double k;
    Vec3 field;
    std::vector<Term> terms;
};
}
#endif
