#ifndef OPENMM_SYSTEM_H_
#define OPENMM_SYSTEM_H_
// shim, see Vec3.h
#include "Force.h"
#include "VirtualSite.h"
#include <vector>
namespace OpenMM {
class System {
public:
    System() {}
    ~System() {
        for (size_t i = 0; i < forces.size(); i++) delete forces[i];
        for (size_t i = 0; i < sites.size(); i++) delete sites[i];
    }
    int addParticle(double mass) { masses.push_back(mass); virtualSite.push_back(false); sites.push_back(NULL); return (int) masses.size() - 1; }
    int getNumParticles() const { return (int) masses.size(); }
    double getParticleMass(int index) const { return masses[index]; }
    void setVirtualSite(int index, bool flag) { virtualSite[index] = flag; }     // shim: a bare flag (site without geometry)
    void setVirtualSite(int index, VirtualSite* site) { delete sites[index]; sites[index] = site; virtualSite[index] = site != NULL; }   // takes ownership
    const VirtualSite* getVirtualSite(int index) const { return sites[index]; }
    bool isVirtualSite(int index) const { return virtualSite[index]; }
    int addConstraint(int particle1, int particle2, double distance) {
        Constraint c = {particle1, particle2, distance};
        constraints.push_back(c);
        return (int) constraints.size() - 1;
    }
    int getNumConstraints() const { return (int) constraints.size(); }
    void getConstraintParameters(int index, int& particle1, int& particle2, double& distance) const {
        particle1 = constraints[index].p1; particle2 = constraints[index].p2; distance = constraints[index].d;
    }
    int addForce(Force* force) { forces.push_back(force); return (int) forces.size() - 1; }     // takes ownership
    int getNumForces() const { return (int) forces.size(); }
    const Force& getForce(int index) const { return *forces[index]; }
private:
    System(const System&);
    System& operator=(const System&);
    struct Constraint { int p1, p2; double d; };
    std::vector<double> masses;
    std::vector<bool> virtualSite;
    std::vector<VirtualSite*> sites;
    std::vector<Constraint> constraints;
    std::vector<Force*> forces;
};
}
#endif
