#ifndef RBK_GLUE_RIGIDBODYINTEGRATOR_H_
#define RBK_GLUE_RIGIDBODYINTEGRATOR_H_
// RigidBodyPlugin::RigidBodyIntegrator with the reference's public interface
// (openmmapi/include/RigidBodyIntegrator.h:49-137): same constructor, same methods, same exceptions.
#include "RigidBodySystem.h"
#include "openmm/Context.h"
#include "openmm/Integrator.h"
#include "openmm/Kernel.h"
#include <string>
#include <vector>

namespace RigidBodyPlugin {

class RigidBodyIntegrator : public OpenMM::Integrator {
public:
    // stepSize in ps; bodyIndices[i] = rigid body of atom i (<= 0: free atom)
    explicit RigidBodyIntegrator(double stepSize, const std::vector<int>& bodyIndices);
    void setRotationMode(int mode);                  // 0 = exact (default), n = NO-SQUISH with n sub-steps
    int getRotationMode() const { return rotationMode; }
    void setComputeRefinedEnergies(bool compute);
    bool getComputeRefinedEnergies() const { return computeRefinedEnergies; }
    void step(int steps);
    std::vector<int> getBodyIndices() const { return bodyIndices; }
    const RigidBodySystem& getRigidBodySystem() const { return bodySystem; }
    OpenMM::ContextImpl& getContextImpl() { return *context; }      // (tests) the ContextImpl this integrator is bound to
    std::vector<double> getKineticEnergies();        // {translational, rotational}
    std::vector<double> getRefinedKineticEnergies();
    double getPotentialEnergyRefinement();
protected:
    void initialize(OpenMM::ContextImpl& context);
    void cleanup();
    std::vector<std::string> getKernelNames();
    void stateChanged(OpenMM::State::DataType changed);
    double computeKineticEnergy();
private:
    std::vector<int> bodyIndices;
    RigidBodySystem bodySystem;
    OpenMM::Kernel kernel;
    int rotationMode;
    bool computeRefinedEnergies;
};

} // namespace RigidBodyPlugin
#endif
