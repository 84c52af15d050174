// rbk_host.hpp - host-side rigid-body model of librbk (index mapping, body build).
//
// Product code.  Everything here runs once per Context creation / setPositions / setVelocities
// (never per step); the per-step work lives in rbk_kernels.cu.  The arithmetic follows the
// reference's host model so that body constants (I, q, d) come out identical:
//   openmmapi/src/RigidBodySystem.cpp:28-142, openmmapi/src/RigidBody.cpp:26-142,
//   openmmapi/src/eigenDecomposition.cpp:36-176, openmmapi/src/MatVec.cpp:344-373.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace rbk {

struct HostBody {
    int N = 0, dof = 0, loc = 0;
    double mass = 0, invMass = 0;
    double I[3] = {0, 0, 0}, invI[3] = {0, 0, 0};
    double rcm[3] = {0, 0, 0}, pcm[3] = {0, 0, 0};
    double q[4] = {0, 0, 0, 0}, pi[4] = {0, 0, 0, 0};
    double force[3] = {0, 0, 0}, torque[4] = {0, 0, 0, 0}, tau[3] = {0, 0, 0};
    double twoKt = 0, twoKr = 0;
};

struct HostModel {
    int numAtoms = 0, numBodies = 0, numFree = 0, numActualAtoms = 0, numBodyAtoms = 0, numDOF = 0;
    int numConstraints = 0, rotationMode = 0;
    std::vector<int> bodyIndex;        // cleaned label per atom (0 = free)
    std::vector<int> atomIndex;        // [free atoms..., body 1 atoms..., body 2 atoms...]
    std::vector<double> mass;          // per atom
    std::vector<uint8_t> isVirtual;
    std::vector<HostBody> body;
    std::vector<double> d;             // body-frame coordinates, 3 per body atom
    std::vector<double> delta;         // space-frame displacements from the centre of mass
    std::vector<double> freeInvMass;

    // Returns an empty string on success, else the error message.
    std::string initialize(int numAtoms, const int* bodyIndices, const double* masses, const unsigned char* isVirtual,
                           int numConstraints, const int* constraintAtoms, int rotationMode);
    void update(const double* R, const double* V, const double* F, bool geometry, bool velocities);

private:
    void buildGeometry(HostBody& b, const double* R, const double* F);
    void buildDynamics(HostBody& b, const double* V);
};

} // namespace rbk
