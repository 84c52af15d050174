#ifndef OPENMM_CUDACONTEXT_H_
#define OPENMM_CUDACONTEXT_H_
// shim, see ../Vec3.h: the part of OpenMM's CudaContext an integrator plugin touches
// (platforms/cuda/src/CudaRigidBodyKernels.cpp of the reference: setAsCurrent, getPosq, getPosqCorrection, getVelm,
// getForce, getAtomIndex, getNumAtoms, getPaddedNumAtoms, getUseMixedPrecision, getUseDoublePrecision,
// getIntegrationUtilities, addReorderListener, reorderAtoms, set/getTime, set/getStepCount, getPinnedBuffer, TileSize,
// intToString).  Same device data layout as the real one: posq real4 (+ posqCorrection in mixed precision), velm
// mixed4 with the inverse mass in .w, force = long long[3*paddedNumAtoms] fixed point (scale 2^32), atoms stored in a
// reordered order (atomIndex[slot] = particle).  Functional, not fast: everything the real platform does with its own
// kernels (force evaluation, constraints, virtual sites, reordering) is done here on the host through download/upload.
#include "openmm/System.h"
#include "openmm/Vec3.h"
#include "openmm/cuda/CudaArray.h"
#include <cmath>
#include <sstream>
#include <string>
#include <vector>
namespace OpenMM {
class CudaIntegrationUtilities;
class CudaContext {
public:
    static const int TileSize = 32;
    class ReorderListener {
    public:
        virtual void execute() = 0;
        virtual ~ReorderListener() {}
    };
    CudaContext(const System& system, const std::string& precision);
    ~CudaContext();
    void setAsCurrent() {}
    const System& getSystem() const { return system; }
    int getNumAtoms() const { return numAtoms; }
    int getPaddedNumAtoms() const { return paddedNumAtoms; }
    bool getUseDoublePrecision() const { return useDouble; }
    bool getUseMixedPrecision() const { return useMixed; }
    CudaArray& getPosq() { return *posq; }
    CudaArray& getPosqCorrection() { return *posqCorrection; }
    CudaArray& getVelm() { return *velm; }
    CudaArray& getForce() { return *force; }
    const std::vector<int>& getAtomIndex() const { return atomIndex; }
    CudaIntegrationUtilities& getIntegrationUtilities() { return *integration; }
    CUstream getCurrentStream() { return 0; }
    void addReorderListener(ReorderListener* listener) { listeners.push_back(listener); }       // takes ownership
    void reorderAtoms();
    double getTime() const { return time; }
    void setTime(double t) { time = t; }
    int getStepCount() const { return stepCount; }
    void setStepCount(int steps) { stepCount = steps; }
    void* getPinnedBuffer() { return pinned; }
    static std::string intToString(int value) { std::stringstream s; s << value; return s.str(); }
    // ---- shim-only helpers (what Context / ContextImpl / the platform's own kernels do in real OpenMM)
    void setReorderInterval(int steps) { reorderInterval = steps; }       // 0 = atoms are never reordered
    void setReorderUnit(int atoms) { reorderUnit = atoms; }               // permute blocks of this many consecutive particles ("molecules")
    int getNumReorders() const { return numReorders; }
    void uploadPositions(const std::vector<Vec3>& positions);
    void uploadVelocities(const std::vector<Vec3>& velocities);
    void uploadForces(const std::vector<Vec3>& forces);
    void downloadPositions(std::vector<Vec3>& positions);
    void downloadVelocities(std::vector<Vec3>& velocities);
    void downloadForces(std::vector<Vec3>& forces);
    std::vector<double> inverseMasses() const;
private:
    template <class T4> void put(CudaArray& array, const std::vector<Vec3>& values, const std::vector<double>* w);
    template <class T4> void get(CudaArray& array, std::vector<Vec3>& values);
    template <class T4> void permute(CudaArray& array, const std::vector<int>& oldSlotOfNew);
    const System& system;
    int numAtoms, paddedNumAtoms, stepCount, reorderInterval, sinceReorder, numReorders, reorderUnit;
    bool useDouble, useMixed;
    double time;
    CudaArray *posq, *posqCorrection, *velm, *force;
    CudaIntegrationUtilities* integration;
    std::vector<int> atomIndex;
    std::vector<ReorderListener*> listeners;
    void* pinned;
    unsigned long long rngState;
};
}
#include "openmm/cuda/CudaIntegrationUtilities.h"
namespace OpenMM {
inline CudaContext::CudaContext(const System& system, const std::string& precision)
    : system(system), numAtoms(system.getNumParticles()), stepCount(0), reorderInterval(0), sinceReorder(0), numReorders(0), reorderUnit(1),
      useDouble(precision == "double"), useMixed(precision == "mixed"), time(0.0), posqCorrection(NULL), pinned(NULL), rngState(88172645463325252ULL) {
    if (!useDouble && !useMixed && precision != "single") throw OpenMMException("Illegal value for CudaPrecision: " + precision);
    paddedNumAtoms = TileSize*((numAtoms + TileSize - 1)/TileSize);
    const int realSize = useDouble ? sizeof(double4) : sizeof(float4), mixedSize = useDouble || useMixed ? sizeof(double4) : sizeof(float4);
    posq = new CudaArray(*this, paddedNumAtoms, realSize, "posq");
    if (useMixed) posqCorrection = new CudaArray(*this, paddedNumAtoms, sizeof(float4), "posqCorrection");
    velm = new CudaArray(*this, paddedNumAtoms, mixedSize, "velm");
    force = new CudaArray(*this, paddedNumAtoms*3, sizeof(long long), "force");
    atomIndex.resize(paddedNumAtoms);
    for (int i = 0; i < paddedNumAtoms; i++) atomIndex[i] = i;
    cudaMallocHost(&pinned, (size_t) paddedNumAtoms*3*sizeof(long long) + 4096);
    integration = new CudaIntegrationUtilities(*this, system);
    uploadVelocities(std::vector<Vec3>(numAtoms));            // velm.w = inverse mass from the start
}
inline CudaContext::~CudaContext() {
    for (size_t i = 0; i < listeners.size(); i++) delete listeners[i];
    delete integration; delete posq; delete posqCorrection; delete velm; delete force;
    cudaFreeHost(pinned);
}
inline std::vector<double> CudaContext::inverseMasses() const {
    std::vector<double> w(numAtoms);
    for (int i = 0; i < numAtoms; i++) { const double m = system.getParticleMass(i); w[i] = m == 0.0 ? 0.0 : 1.0/m; }
    return w;
}
// values are in PARTICLE order, the device arrays in slot order: slot s holds particle atomIndex[s]; .w is kept unless given
template <class T4> void CudaContext::put(CudaArray& array, const std::vector<Vec3>& values, const std::vector<double>* w) {
    std::vector<T4> host;
    array.download(host);
    for (int s = 0; s < paddedNumAtoms; s++) {
        const int p = atomIndex[s];
        if (p >= numAtoms) continue;
        host[s].x = values[p][0]; host[s].y = values[p][1]; host[s].z = values[p][2];
        if (w) host[s].w = (*w)[p];
    }
    array.upload(host);
}
template <class T4> void CudaContext::get(CudaArray& array, std::vector<Vec3>& values) {
    std::vector<T4> host;
    array.download(host);
    values.resize(numAtoms);
    for (int s = 0; s < paddedNumAtoms; s++) if (atomIndex[s] < numAtoms) values[atomIndex[s]] = Vec3(host[s].x, host[s].y, host[s].z);
}
inline void CudaContext::uploadPositions(const std::vector<Vec3>& positions) {
    if (useDouble) { put<double4>(*posq, positions, NULL); return; }
    put<float4>(*posq, positions, NULL);
    if (useMixed) {                                           // value = (float) hi + (float) lo
        std::vector<float4> hi, lo;
        posq->download(hi); posqCorrection->download(lo);
        for (int s = 0; s < paddedNumAtoms; s++) {
            const int p = atomIndex[s];
            if (p >= numAtoms) continue;
            lo[s].x = (float) (positions[p][0] - (double) hi[s].x); lo[s].y = (float) (positions[p][1] - (double) hi[s].y);
            lo[s].z = (float) (positions[p][2] - (double) hi[s].z);
        }
        posqCorrection->upload(lo);
    }
}
inline void CudaContext::downloadPositions(std::vector<Vec3>& positions) {
    if (useDouble) { get<double4>(*posq, positions); return; }
    get<float4>(*posq, positions);
    if (useMixed) {
        std::vector<Vec3> lo;
        get<float4>(*posqCorrection, lo);
        for (int i = 0; i < numAtoms; i++) positions[i] += lo[i];
    }
}
inline void CudaContext::uploadVelocities(const std::vector<Vec3>& velocities) {
    const std::vector<double> w = inverseMasses();
    if (useDouble || useMixed) put<double4>(*velm, velocities, &w); else put<float4>(*velm, velocities, &w);
}
inline void CudaContext::downloadVelocities(std::vector<Vec3>& velocities) {
    if (useDouble || useMixed) get<double4>(*velm, velocities); else get<float4>(*velm, velocities);
}
inline void CudaContext::uploadForces(const std::vector<Vec3>& forces) {
    std::vector<long long> host((size_t) paddedNumAtoms*3, 0);
    for (int s = 0; s < paddedNumAtoms; s++) {
        const int p = atomIndex[s];
        if (p >= numAtoms) continue;
        for (int c = 0; c < 3; c++) host[s + (size_t) c*paddedNumAtoms] = (long long) std::llrint(forces[p][c]*4294967296.0);
    }
    force->upload(host);
}
inline void CudaContext::downloadForces(std::vector<Vec3>& forces) {
    std::vector<long long> host;
    force->download(host);
    forces.resize(numAtoms);
    for (int s = 0; s < paddedNumAtoms; s++) {
        const int p = atomIndex[s];
        if (p >= numAtoms) continue;
        forces[p] = Vec3(host[s]/4294967296.0, host[s + (size_t) paddedNumAtoms]/4294967296.0, host[s + (size_t) 2*paddedNumAtoms]/4294967296.0);
    }
}
template <class T4> void CudaContext::permute(CudaArray& array, const std::vector<int>& oldSlotOfNew) {
    std::vector<T4> host, moved;
    array.download(host);
    moved.resize(host.size());
    for (int s = 0; s < paddedNumAtoms; s++) moved[s] = host[oldSlotOfNew[s]];
    array.upload(moved);
}
// Real OpenMM sorts molecules along a space-filling curve every few hundred steps when a cutoff is in use; the shim
// applies a pseudo-random permutation of the atoms every `reorderInterval` calls - the most general case a listener
// can meet.  posq, posqCorrection and velm move; the forces are NOT moved (they are recomputed before OpenMM's own
// integrators read them - a plugin that reads them earlier must move them itself, which is what its listener is for).
inline void CudaContext::reorderAtoms() {
    if (reorderInterval <= 0 || ++sinceReorder < reorderInterval) return;
    sinceReorder = 0;
    numReorders++;
    std::vector<int> oldSlotOfNew(paddedNumAtoms);
    for (int s = 0; s < paddedNumAtoms; s++) oldSlotOfNew[s] = s;
    const int unit = reorderUnit > 0 && numAtoms % reorderUnit == 0 ? reorderUnit : 1, units = numAtoms/unit;
    std::vector<int> unitOrder(units);
    for (int u = 0; u < units; u++) unitOrder[u] = u;
    for (int u = units - 1; u > 0; u--) {                    // Fisher-Yates over the units (padding stays at the end)
        rngState ^= rngState << 13; rngState ^= rngState >> 7; rngState ^= rngState << 17;
        std::swap(unitOrder[u], unitOrder[(int) (rngState % (unsigned long long) (u + 1))]);
    }
    for (int u = 0; u < units; u++)
        for (int j = 0; j < unit; j++) oldSlotOfNew[u*unit + j] = unitOrder[u]*unit + j;
    if (useDouble) permute<double4>(*posq, oldSlotOfNew); else permute<float4>(*posq, oldSlotOfNew);
    if (useMixed) permute<float4>(*posqCorrection, oldSlotOfNew);
    if (useDouble || useMixed) permute<double4>(*velm, oldSlotOfNew); else permute<float4>(*velm, oldSlotOfNew);
    std::vector<int> newIndex(paddedNumAtoms);
    for (int s = 0; s < paddedNumAtoms; s++) newIndex[s] = atomIndex[oldSlotOfNew[s]];
    atomIndex = newIndex;
    for (size_t i = 0; i < listeners.size(); i++) listeners[i]->execute();
}
}
#endif
