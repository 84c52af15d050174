#ifndef RBK_GLUE_RIGIDBODYSYSTEM_H_
#define RBK_GLUE_RIGIDBODYSYSTEM_H_
// RigidBodyPlugin::RigidBodySystem as the rest of the plugin sees it (reference: openmmapi/include/RigidBodySystem.h),
// backed by a librbk handle instead of a host-side stepper.  Host-side work (index mapping, body build) happens
// inside librbk (rbk_create / rbk_update); this class owns the handle and exposes the reference's getters.
#include "openmm/internal/ContextImpl.h"
#include "rbk.h"
#include <vector>

namespace RigidBodyPlugin {

class RigidBodySystem {
public:
    RigidBodySystem() : handle(NULL), transKE(0.0), rotKE(0.0) {}
    ~RigidBodySystem() { rbk_destroy(handle); }
    void initialize(OpenMM::ContextImpl& context, const std::vector<int>& bodyIndices, int rotationMode);
    void update(OpenMM::ContextImpl& context, bool geometry, bool velocities);
    int getNumDOF() const { return count(4); }
    int getNumFree() const { return count(1); }
    int getNumBodies() const { return count(0); }
    int getNumActualAtoms() const { return count(2); }
    int getNumBodyAtoms() const { return count(3); }
    int getAtomIndex(int i) const;
    const std::vector<int>& getAtomIndices() const { return atomIndex; }      // plugin order -> particle, all of it
    double getTranslationalEnergy() const { return transKE; }
    double getRotationalEnergy() const { return rotKE; }
    double getKineticEnergy() const { return transKE + rotKE; }
    void setKineticEnergies(double trans, double rot) { transKE = trans; rotKE = rot; }
    rbk_system* getHandle() const { return handle; }
private:
    RigidBodySystem(const RigidBodySystem&);
    RigidBodySystem& operator=(const RigidBodySystem&);
    int count(int which) const;
    rbk_system* handle;
    std::vector<int> atomIndex;              // filled once in initialize (the index mapping never changes afterwards)
    double transKE, rotKE;
};

} // namespace RigidBodyPlugin
#endif
