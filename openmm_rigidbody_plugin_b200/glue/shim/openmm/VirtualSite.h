#ifndef OPENMM_VIRTUALSITE_H_
#define OPENMM_VIRTUALSITE_H_
// shim, see Vec3.h: the one virtual-site class the glue tests need
namespace OpenMM {
class VirtualSite {
public:
    virtual ~VirtualSite() {}
    virtual int getNumParticles() const = 0;
    virtual int getParticle(int particle) const = 0;
};
class TwoParticleAverageSite : public VirtualSite {
public:
    TwoParticleAverageSite(int particle1, int particle2, double weight1, double weight2) : w1(weight1), w2(weight2) { p[0] = particle1; p[1] = particle2; }
    int getNumParticles() const { return 2; }
    int getParticle(int particle) const { return p[particle]; }
    double getWeight(int particle) const { return particle == 0 ? w1 : w2; }
private:
    int p[2];
    double w1, w2;
};
}
#endif
