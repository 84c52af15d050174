/* Minimal stand-in for openmm/System.h: only what RigidBodySystem.cpp consumes
 * (particle masses, virtual-site flags, constraints).  Test infrastructure. */
#ifndef RBK_SHIM_OPENMM_SYSTEM_H_
#define RBK_SHIM_OPENMM_SYSTEM_H_
#include <vector>
namespace OpenMM {
class System {
public:
    int addParticle(double mass) { masses.push_back(mass); virtualSite.push_back(false); return (int) masses.size() - 1; }
    void setVirtualSite(int i, bool flag) { virtualSite[i] = flag; }
    int addConstraint(int a, int b, double d) { c1.push_back(a); c2.push_back(b); cd.push_back(d); return (int) c1.size() - 1; }
    int getNumParticles() const { return (int) masses.size(); }
    double getParticleMass(int i) const { return masses[i]; }
    bool isVirtualSite(int i) const { return virtualSite[i]; }
    int getNumConstraints() const { return (int) c1.size(); }
    void getConstraintParameters(int i, int& a, int& b, double& d) const { a = c1[i]; b = c2[i]; d = cd[i]; }
private:
    std::vector<double> masses;
    std::vector<bool> virtualSite;
    std::vector<int> c1, c2;
    std::vector<double> cd;
};
}
#endif
