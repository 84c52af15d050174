// RigidBodySystem facade over a librbk handle (reference: openmmapi/src/RigidBodySystem.cpp:55-142).
#include "RigidBodySystem.h"
#include "openmm/OpenMMException.h"

using namespace RigidBodyPlugin;
using namespace OpenMM;
using std::vector;

static void check(int rc) {
    if (rc != RBK_OK) throw OpenMMException(rbk_last_error());
}

void RigidBodySystem::initialize(ContextImpl& context, const vector<int>& bodyIndices, int rotationMode) {
    const System& system = context.getSystem();
    const int n = system.getNumParticles();
    vector<double> mass(n);
    vector<unsigned char> isVirtual(n);
    for (int i = 0; i < n; i++) {
        mass[i] = system.getParticleMass(i);
        isVirtual[i] = system.isVirtualSite(i) ? 1 : 0;
    }
    vector<int> constraintAtoms;
    for (int i = 0; i < system.getNumConstraints(); i++) {
        int a, b;
        double d;
        system.getConstraintParameters(i, a, b, d);
        constraintAtoms.push_back(a);
        constraintAtoms.push_back(b);
    }
    rbk_destroy(handle);
    handle = NULL;
    check(rbk_create(n, bodyIndices.data(), mass.data(), isVirtual.data(), system.getNumConstraints(),
                     constraintAtoms.empty() ? NULL : constraintAtoms.data(), rotationMode, &handle));
    atomIndex.assign(getNumActualAtoms(), 0);
    if (!atomIndex.empty()) check(rbk_get_atom_index(handle, atomIndex.data()));
}

void RigidBodySystem::update(ContextImpl& context, bool geometry, bool velocities) {
    vector<Vec3> R, V, F;
    context.getPositions(R);
    context.getVelocities(V);
    context.getForces(F);
    // OpenMM::Vec3 is three contiguous doubles, i.e. RBK_LAYOUT_VEC3
    check(rbk_update(handle, R.empty() ? NULL : &R[0][0], V.empty() ? NULL : &V[0][0], F.empty() ? NULL : &F[0][0],
                     geometry ? 1 : 0, velocities ? 1 : 0));
}

int RigidBodySystem::count(int which) const {
    if (handle == NULL) return 0;
    int c[5];
    check(rbk_get_counts(handle, c));
    return c[which];
}

int RigidBodySystem::getAtomIndex(int i) const { return atomIndex.at(i); }
