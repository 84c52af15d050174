// B200 kernel on OpenMM's CUDA platform.  Step protocol = CudaIntegrateRigidBodyStepKernel::execute
// (platforms/cuda/src/CudaRigidBodyKernels.cpp:377-444):
//   updateContextState -> [free atoms: posDelta, applyConstraints] -> Part 1 -> computeVirtualSites -> forces ->
//   Part 2 -> [applyVelocityConstraints] -> time / step count -> reorderAtoms,
// every arrow a librbk call on the CudaContext's own device arrays; nothing is copied per step.
// step(n) with n > 1 (executeSteps) runs Part 2 of one step and Part 1 of the next as ONE pass (rbk_part2_part1_openmm)
// whenever nothing can sit between them: no free-atom constraints, no diagnostics, and a System whose Forces are all of
// classes known not to touch the state in updateContextState.
#include "B200CudaRigidBodyKernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/cuda/CudaIntegrationUtilities.h"
#include <cstring>
#include <typeinfo>

using namespace RigidBodyPlugin;
using namespace OpenMM;
using std::vector;

static void check(int rc) {
    if (rc != RBK_OK) throw OpenMMException(rbk_last_error());
}

// ReorderListener (CudaRigidBodyKernels.cpp:69-113): after cu.reorderAtoms() moved posq / velm, rebuild
// atomLocation[i] = invOrder[atomIndex(i)] and move the forces of the owned atoms to their new slots.
class B200CudaIntegrateRigidBodyStepKernel::ReorderListener : public CudaContext::ReorderListener {
public:
    ReorderListener(B200CudaIntegrateRigidBodyStepKernel& kernel) : kernel(kernel) {}
    void execute() { kernel.atomsReordered(); }
private:
    B200CudaIntegrateRigidBodyStepKernel& kernel;
};

vector<int> B200CudaIntegrateRigidBodyStepKernel::currentLocation() const {
    const vector<int>& order = cu.getAtomIndex();
    vector<int> invOrder(order.size());
    for (size_t i = 0; i < order.size(); i++) invOrder[order[i]] = (int) i;
    vector<int> location(atomIndex.size());
    for (size_t i = 0; i < atomIndex.size(); i++) location[i] = invOrder[atomIndex[i]];
    return location;
}

void B200CudaIntegrateRigidBodyStepKernel::atomsReordered() {
    if (system == NULL) return;
    const vector<int> location = currentLocation();
    check(rbk_reorder_openmm(system, location.empty() ? NULL : location.data(), force(), cu.getPaddedNumAtoms(), stream()));
}

// Force classes whose ForceImpl::updateContextState is the base-class no-op (OpenMM 7.x): with only these in the
// System nothing happens between two execute() calls and step(n) may fuse across the step boundary.
static bool isStateless(const Force& force) {
    static const char* const known[] = {
        "HarmonicBondForce", "HarmonicAngleForce", "PeriodicTorsionForce", "RBTorsionForce", "CMAPTorsionForce", "NonbondedForce",
        "GBSAOBCForce", "CustomBondForce", "CustomAngleForce", "CustomTorsionForce", "CustomExternalForce", "CustomNonbondedForce",
        "CustomGBForce", "CustomHbondForce", "CustomCompoundBondForce", "CustomCentroidBondForce", "CustomManyParticleForce"};
    const char* name = typeid(force).name();              // Itanium ABI: N6OpenMM<len><class>E
    for (size_t i = 0; i < sizeof(known)/sizeof(known[0]); i++) {
        const char* hit = std::strstr(name, known[i]);
        if (hit != NULL && hit[std::strlen(known[i])] == 'E') return true;
    }
    return false;
}

void B200CudaIntegrateRigidBodyStepKernel::initialize(ContextImpl& context, const RigidBodyIntegrator& integrator) {
    cu.setAsCurrent();
    precision = cu.getUseDoublePrecision() ? RBK_OPENMM_DOUBLE : (cu.getUseMixedPrecision() ? RBK_OPENMM_MIXED : RBK_OPENMM_SINGLE);
    refined = integrator.getComputeRefinedEnergies();
    const System& sys = context.getSystem();
    constrained = sys.getNumConstraints() != 0;
    statelessForces = true;
    for (int i = 0; i < sys.getNumForces(); i++) statelessForces = statelessForces && isStateless(sys.getForce(i));
    const RigidBodySystem& bodySystem = integrator.getRigidBodySystem();
    numBodies = bodySystem.getNumBodies();
    numFree = bodySystem.getNumFree();
    system = bodySystem.getHandle();
    atomIndex = bodySystem.getAtomIndices();
    // atomLocation (CudaRigidBodyKernels.cpp:277-284); allocates the handle's device state on first use
    const vector<int> location = currentLocation();
    check(rbk_set_atom_location(system, location.empty() ? NULL : location.data(), stream()));
    cu.addReorderListener(new ReorderListener(*this));      // owned by the CudaContext
}

// stateChanged (RigidBodyIntegrator.cpp:63-74) without leaving the device: RigidBodySystem::update evaluated by
// rbk_update_device_openmm straight from posq / velm / force - no download, no host rebuild, no upload.
bool B200CudaIntegrateRigidBodyStepKernel::updateBodySystem(ContextImpl&, RigidBodySystem& bodySystem, bool geometry, bool velocities) {
    cu.setAsCurrent();
    bodies = &bodySystem;
    system = bodySystem.getHandle();
    check(rbk_update_device_openmm(system, posq(), posqCorrection(), velm(), force(), cu.getPaddedNumAtoms(), precision,
                                   geometry ? 1 : 0, velocities ? 1 : 0, stream()));
    if (refined) check(rbk_set_refined_energies(system, numFree != 0 && constrained ? RBK_REFINED_BODIES : RBK_REFINED_ALL, stream()));
    return true;
}

// The reference's path: a body system built on the host (RigidBodySystem::update) goes to the device
// (CudaRigidBodyKernels.cpp:293-372).  Kept for callers that build on the host; stateChanged uses updateBodySystem.
void B200CudaIntegrateRigidBodyStepKernel::uploadBodySystem(RigidBodySystem& bodySystem) {
    cu.setAsCurrent();
    bodies = &bodySystem;
    system = bodySystem.getHandle();
    check(rbk_upload(system, stream()));
    const vector<int> location = currentLocation();
    check(rbk_set_atom_location(system, location.empty() ? NULL : location.data(), stream()));
    if (refined) check(rbk_set_refined_energies(system, numFree != 0 && constrained ? RBK_REFINED_BODIES : RBK_REFINED_ALL, stream()));
}

// Part 1 with its free-atom constraint hooks (CudaRigidBodyKernels.cpp:405-423)
void B200CudaIntegrateRigidBodyStepKernel::firstHalf(const RigidBodyIntegrator& integrator) {
    const double dt = integrator.getStepSize(), tol = integrator.getConstraintTolerance();
    const int padded = cu.getPaddedNumAtoms();
    if (numFree == 0 || !constrained) {                     // nothing for applyConstraints to do: free atoms move by v dt
        check(rbk_part1_openmm(system, dt, posq(), posqCorrection(), velm(), force(), padded, precision, stream()));
        return;
    }
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    void* posDelta = (void*) integration.getPosDelta().getDevicePointer();
    if (refined) {                                          // virtual backward step of the free atoms (:406-416)
        integration.applyConstraints(tol);
        check(rbk_free_delta_openmm(system, -dt, velm(), force(), padded, precision, posDelta, stream()));
        check(rbk_free_dot_openmm(system, posDelta, precision, -1.0, 1, stream()));
    }
    check(rbk_free_delta_openmm(system, dt, velm(), force(), padded, precision, posDelta, stream()));
    integration.applyConstraints(tol);
    check(rbk_part1_delta_openmm(system, dt, posq(), posqCorrection(), velm(), force(), padded, precision, posDelta, stream()));
}

// what follows Part 2 (:427-438)
void B200CudaIntegrateRigidBodyStepKernel::afterSecondHalf(const RigidBodyIntegrator& integrator) {
    if (numFree == 0 || !constrained) return;
    const double dt = integrator.getStepSize(), tol = integrator.getConstraintTolerance();
    CudaIntegrationUtilities& integration = cu.getIntegrationUtilities();
    integration.applyVelocityConstraints(tol);
    if (refined) {
        const int padded = cu.getPaddedNumAtoms();
        void* posDelta = (void*) integration.getPosDelta().getDevicePointer();
        check(rbk_free_dot_openmm(system, posDelta, precision, 5.0, 0, stream()));
        check(rbk_free_delta_openmm(system, dt, velm(), force(), padded, precision, posDelta, stream()));
        integration.applyConstraints(tol);
        check(rbk_free_dot_openmm(system, posDelta, precision, 2.0, 0, stream()));
    }
}

void B200CudaIntegrateRigidBodyStepKernel::endOfStep(const RigidBodyIntegrator& integrator) {
    cu.setTime(cu.getTime() + integrator.getStepSize());
    cu.setStepCount(cu.getStepCount() + 1);
    stepsTaken++;
    cu.reorderAtoms();
}

void B200CudaIntegrateRigidBodyStepKernel::execute(ContextImpl& context, const RigidBodyIntegrator& integrator) {
    if (system == NULL || bodies == NULL) throw OpenMMException("B200 rigid-body kernel: positions have not been set");
    context.updateContextState();
    cu.setAsCurrent();
    firstHalf(integrator);
    cu.getIntegrationUtilities().computeVirtualSites();
    context.calcForcesAndEnergy(true, false);
    check(rbk_part2_openmm(system, integrator.getStepSize(), posq(), posqCorrection(), velm(), force(), cu.getPaddedNumAtoms(), precision,
                           stream()));
    afterSecondHalf(integrator);
    endOfStep(integrator);
}

void B200CudaIntegrateRigidBodyStepKernel::executeSteps(ContextImpl& context, const RigidBodyIntegrator& integrator, int steps) {
    const bool fuse = steps > 1 && !refined && !(numFree != 0 && constrained) && statelessForces;
    if (!fuse) {
        for (int i = 0; i < steps; i++) execute(context, integrator);
        return;
    }
    if (system == NULL || bodies == NULL) throw OpenMMException("B200 rigid-body kernel: positions have not been set");
    const double dt = integrator.getStepSize();
    const int padded = cu.getPaddedNumAtoms();
    cu.setAsCurrent();
    context.updateContextState();                           // a no-op for every Force of this System (isStateless)
    check(rbk_part1_openmm(system, dt, posq(), posqCorrection(), velm(), force(), padded, precision, stream()));
    for (int i = 0; i < steps; i++) {
        cu.getIntegrationUtilities().computeVirtualSites();
        context.calcForcesAndEnergy(true, false);
        if (i < steps - 1) {
            // Part 2 of step i + Part 1 of step i+1 in one pass.  cu.reorderAtoms() of step i moves behind it: once both
            // half kicks are done nothing needs the old forces any more, and the new order is in place before the next
            // force evaluation exactly as after the reference's execute().
            context.updateContextState();
            check(rbk_part2_part1_openmm(system, dt, posq(), posqCorrection(), velm(), force(), padded, precision, stream()));
        }
        else check(rbk_part2_openmm(system, dt, posq(), posqCorrection(), velm(), force(), padded, precision, stream()));
        endOfStep(integrator);
    }
}

double B200CudaIntegrateRigidBodyStepKernel::computeKineticEnergy(ContextImpl&, const RigidBodyIntegrator& integrator) {
    vector<double> ke = getKineticEnergies(integrator);
    return ke[0] + ke[1];
}

vector<double> B200CudaIntegrateRigidBodyStepKernel::getKineticEnergies(const RigidBodyIntegrator&) {
    vector<double> ke(2, 0.0);
    if (system == NULL || bodies == NULL) return ke;
    cu.setAsCurrent();
    check(rbk_kinetic_openmm(system, velm(), precision, ke.data(), stream()));
    bodies->setKineticEnergies(ke[0], ke[1]);
    return ke;
}

// CudaRigidBodyKernels.cpp:469-476: the refined estimate when it was switched on, else the plain energies
vector<double> B200CudaIntegrateRigidBodyStepKernel::getRefinedKineticEnergies(const RigidBodyIntegrator& integrator) {
    if (!refined || system == NULL || stepsTaken == 0) return getKineticEnergies(integrator);
    vector<double> ke(2, 0.0);
    cu.setAsCurrent();
    check(rbk_refined_kinetic_openmm(system, integrator.getStepSize(), velm(), precision, ke.data(), stream()));
    return ke;
}

// CudaRigidBodyKernels.cpp:481-494
double B200CudaIntegrateRigidBodyStepKernel::getPotentialEnergyRefinement(const RigidBodyIntegrator& integrator) {
    if (!refined || system == NULL || stepsTaken == 0) return 0.0;
    double out[2] = {0.0, 0.0};
    cu.setAsCurrent();
    check(rbk_potential_refinement_openmm(system, integrator.getStepSize(), force(), cu.getPaddedNumAtoms(), out, stream()));
    return out[0];
}
