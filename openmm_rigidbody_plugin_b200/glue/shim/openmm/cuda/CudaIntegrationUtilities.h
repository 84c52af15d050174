#ifndef OPENMM_CUDAINTEGRATIONUTILITIES_H_
#define OPENMM_CUDAINTEGRATIONUTILITIES_H_
// shim, see ../Vec3.h: getPosDelta, applyConstraints, applyVelocityConstraints, computeVirtualSites with the semantics of
// OpenMM's CudaIntegrationUtilities (constraints act on posDelta = the displacement about to be added to posq), computed
// on the host with the shim's ReferenceConstraints / ReferenceVirtualSites.
#include "openmm/cuda/CudaContext.h"
#include "openmm/reference/ReferenceConstraints.h"
#include "openmm/reference/ReferenceVirtualSites.h"
namespace OpenMM {
class CudaIntegrationUtilities {
public:
    CudaIntegrationUtilities(CudaContext& context, const System& system) : context(context), system(system), constraints(system) {
        const bool mixed = context.getUseDoublePrecision() || context.getUseMixedPrecision();
        posDelta = new CudaArray(context, context.getPaddedNumAtoms(), mixed ? sizeof(double4) : sizeof(float4), "posDelta");
        hasVirtualSites = false;
        for (int i = 0; i < system.getNumParticles(); i++) hasVirtualSites = hasVirtualSites || system.isVirtualSite(i);
    }
    ~CudaIntegrationUtilities() { delete posDelta; }
    CudaArray& getPosDelta() { return *posDelta; }
    void applyConstraints(double tol) {
        if (system.getNumConstraints() == 0) return;
        std::vector<Vec3> pos, delta, moved;
        context.downloadPositions(pos);
        readDelta(delta);
        moved.resize(pos.size());
        for (size_t i = 0; i < pos.size(); i++) moved[i] = pos[i] + delta[i];
        std::vector<double> w = context.inverseMasses();
        constraints.apply(pos, moved, w, tol);
        for (size_t i = 0; i < pos.size(); i++) delta[i] = moved[i] - pos[i];
        writeDelta(delta);
    }
    void applyVelocityConstraints(double tol) {
        if (system.getNumConstraints() == 0) return;
        std::vector<Vec3> pos, vel;
        context.downloadPositions(pos);
        context.downloadVelocities(vel);
        std::vector<double> w = context.inverseMasses();
        constraints.applyToVelocities(pos, vel, w, tol);
        context.uploadVelocities(vel);
    }
    void computeVirtualSites() {
        if (!hasVirtualSites) return;
        std::vector<Vec3> pos;
        context.downloadPositions(pos);
        ReferenceVirtualSites::computePositions(system, pos);
        context.uploadPositions(pos);
    }
private:
    template <class T4> void rw(std::vector<Vec3>& delta, bool write) {
        std::vector<T4> host;
        posDelta->download(host);
        const std::vector<int>& index = context.getAtomIndex();
        const int n = context.getNumAtoms();
        if (!write) delta.assign(n, Vec3());
        for (int s = 0; s < context.getPaddedNumAtoms(); s++) {
            const int p = index[s];
            if (p >= n) continue;
            if (write) { host[s].x = delta[p][0]; host[s].y = delta[p][1]; host[s].z = delta[p][2]; }
            else delta[p] = Vec3(host[s].x, host[s].y, host[s].z);
        }
        if (write) posDelta->upload(host);
    }
    void readDelta(std::vector<Vec3>& d) { if (posDelta->getElementSize() == (int) sizeof(double4)) rw<double4>(d, false); else rw<float4>(d, false); }
    void writeDelta(std::vector<Vec3>& d) { if (posDelta->getElementSize() == (int) sizeof(double4)) rw<double4>(d, true); else rw<float4>(d, true); }
    CudaContext& context;
    const System& system;
    ReferenceConstraints constraints;
    CudaArray* posDelta;
    bool hasVirtualSites;
};
}
#endif
