// rbk_api.cu - the C ABI declared in include/rbk.h: handle, device memory, error plumbing.
// There is no CPU fallback: every device entry point fails with RBK_ECUDA when no GPU is usable.
#include "../../include/rbk.h"
#include "rbk_device.hpp"
#include "rbk_host.hpp"
#include "rbk_math.cuh"

#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using rbk::AtomView;
using rbk::DeviceSystem;
using rbk::HostBody;
using rbk::HostModel;

namespace {

thread_local std::string g_error;

int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}

#define RBK_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (expr);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(RBK_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));         \
    } while (0)

template <class T> cudaError_t devAlloc(T*& p, size_t count) {
    p = nullptr;
    return count ? cudaMalloc((void**) &p, count*sizeof(T)) : cudaSuccess;
}

size_t padTo(size_t n, size_t m) { return (n + m - 1)/m*m; }

// Every host<->device copy librbk issues goes through these two wrappers and is counted (rbk_debug_copy_counters): the
// claim "a step moves nothing between host and device" is then something a test can assert.
std::atomic<long long> g_copies[4];                 // H2D calls, H2D bytes, D2H calls, D2H bytes
void countCopy(size_t bytes, cudaMemcpyKind kind) {
    const int base = kind == cudaMemcpyHostToDevice ? 0 : (kind == cudaMemcpyDeviceToHost ? 2 : -1);
    if (base < 0) return;
    g_copies[base]++;
    g_copies[base + 1] += (long long) bytes;
}
cudaError_t copyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
    countCopy(bytes, kind);
    return cudaMemcpyAsync(dst, src, bytes, kind, st);
}
cudaError_t copySync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    countCopy(bytes, kind);
    return cudaMemcpy(dst, src, bytes, kind);
}

} // namespace

struct rbk_system {
    HostModel host;
    DeviceSystem dev{};
    bool allocated = false, uploaded = false;
    // device allocations
    double* dState = nullptr;
    double* dDxyz = nullptr;
    uint8_t* dLocalBody = nullptr;
    int* dLoc = nullptr;
    int4* dTileMeta = nullptr;
    int4* dBodyTileMeta = nullptr;
    int4* dWarpTileMeta = nullptr;
    int* dTileCounter = nullptr;
    rbk::SeriesControl* dSeriesCtl = nullptr;
    int* hRung = nullptr;            // mapped pinned: the kernels' published rung, read by the launchers as a hint
    rbk::TileMaps tileMaps{};        // TMA descriptors of the one-warp-tile pipeline (valid when dev.tileMaps != NULL)
    int* dAtomLoc = nullptr;
    int* dBodyRun = nullptr;             // caller slot of each body's first atom (storage order), valid when dev.bodyRun != NULL
    int* dPluginLoc = nullptr;           // plugin-order atom -> caller slot, as last set (rbk_reorder_openmm moves forces with it)
    long long* dForcePacked = nullptr;   // scratch of rbk_reorder_openmm
    // Storage order of the bodies / free atoms on the device.  When every body has the same size and the caller keeps each
    // body's atoms together (what OpenMM's reorderAtoms does: whole molecules move), the device arrays are kept SORTED by
    // the caller's slots, so that every kernel streams through the caller's arrays front to back however the caller has
    // permuted its molecules (setLocation re-sorts after each reorder).  Identity otherwise.
    std::vector<int> bodyOrder, bodyPos; // storage position -> plugin body, and its inverse
    std::vector<int> freeOrder, freePos; // the same for the free atoms
    int uniformBodySize = 0;             // > 0: every body has this many atoms
    bool sorted = false;                 // storage order differs from plugin order
    int* dGather = nullptr;              // scratch: gather indices of a re-sort
    double* dScratch = nullptr;          // scratch: destination of a re-sort
    size_t scratchDoubles = 0;
    double* dFreeInvMass = nullptr;
    double* dSavedPos = nullptr;
    double* dAtomMass = nullptr;     // body atoms, storage order (GPU-side body build)
    double* dAtomInvMass = nullptr;  // 1/mass of the body atoms, storage order (full-sector velm stores)
    int* dDofSum = nullptr;
    bool hostStale = false;          // device build used: the host copy of the bodies is not current
    bool deviceAhead = false;        // a step has run since the last upload: r, q, p, pi, F, tau live on the device only
    bool forceTorqueStale = false;   // last step call was rbk_part2_part1: the F / tau planes are one step old
    double* dKinPartial = nullptr;
    unsigned* dKinCounter = nullptr;
    double* dKinOut = nullptr;
    double* hKinOut = nullptr;       // pinned
    int refinedMode = RBK_REFINED_OFF;
    rbk::RefinedState refined{nullptr, nullptr, nullptr};
    // device mirrors for rbk_execute_host
    double* mPos = nullptr;
    double* mVel = nullptr;
    double* mForce = nullptr;
    double* mForce2 = nullptr;       // second force mirror: new forces land here while part 1 still reads the old ones
    rbk::SideStream side{nullptr, nullptr, nullptr};       // free atoms of large-body steps (created on first use)
    cudaStream_t h2dStream = nullptr, d2hStream = nullptr;
    cudaEvent_t evStart = nullptr, evForces = nullptr, evPart1 = nullptr, evPositions = nullptr;
    bool mirrorsLoaded = false;
    bool hostVelStale = false;       // the last rbk_execute_host call left the velocities on the device (V == NULL)
    std::vector<double> staging, oldPositions;
    std::vector<int> locationTable;

    ~rbk_system() {
        cudaFree(dState); cudaFree(dDxyz); cudaFree(dLocalBody); cudaFree(dLoc); cudaFree(dTileMeta); cudaFree(dBodyTileMeta); cudaFree(dWarpTileMeta); cudaFree(dTileCounter); cudaFree(dSeriesCtl);
        cudaFree(dAtomLoc); cudaFree(dBodyRun); cudaFree(dPluginLoc); cudaFree(dForcePacked); cudaFree(dGather); cudaFree(dScratch); cudaFree(dFreeInvMass); cudaFree(dSavedPos); cudaFree(dAtomMass); cudaFree(dAtomInvMass); cudaFree(dDofSum); cudaFree(dKinPartial);
        cudaFree(dKinCounter); cudaFree(dKinOut); cudaFree(refined.rdot); cudaFree(refined.qdot); cudaFree(refined.posDot); cudaFree(mPos); cudaFree(mVel); cudaFree(mForce); cudaFree(mForce2);
        if (side.stream) cudaStreamDestroy(side.stream);
        if (side.fork) cudaEventDestroy(side.fork);
        if (side.join) cudaEventDestroy(side.join);
        if (h2dStream) cudaStreamDestroy(h2dStream);
        if (d2hStream) cudaStreamDestroy(d2hStream);
        for (cudaEvent_t e : {evStart, evForces, evPart1, evPositions}) if (e) cudaEventDestroy(e);
        if (hKinOut) cudaFreeHost(hKinOut);
        if (hRung) cudaFreeHost(hRung);
    }
};

namespace {

// TMA descriptors (CUtensorMap) for the one-warp-tile pipeline: the state planes as a 2-D fp64 tensor [NPLANES][bodyStride]
// with boxes of 32 bodies x 18 / 24 planes, the body-frame coordinates as [3][atomStride] with a box of kWarpTileAtoms x 3.
// cuTensorMapEncodeTiled is a driver-API call; it is looked up at run time so that librbk does not link libcuda.
// Returns false when the driver does not offer it - the kernels then use 1-D bulk copies.
bool encodeTileMaps(rbk_system* sys) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static_assert(sizeof(rbk::TensorMapBlob) == sizeof(CUtensorMap), "TensorMapBlob must match CUtensorMap");
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult found = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &found) != cudaSuccess ||
        found != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return false;
    }
    const EncodeFn encode = (EncodeFn) fn;
    const DeviceSystem& d = sys->dev;
    const cuuint32_t ones[2] = {1, 1};
    auto make = [&](rbk::TensorMapBlob& out, void* base, size_t inner, int rows, int boxInner, int boxRows) {
        const cuuint64_t dims[2] = {(cuuint64_t) inner, (cuuint64_t) rows};
        const cuuint64_t strides[1] = {(cuuint64_t) inner*sizeof(double)};
        const cuuint32_t box[2] = {(cuuint32_t) boxInner, (cuuint32_t) boxRows};
        return encode(reinterpret_cast<CUtensorMap*>(&out), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, ones,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    return make(sys->tileMaps.state18, sys->dState, d.bodyStride, rbk::NPLANES, 32, 18) &&
           make(sys->tileMaps.state24, sys->dState, d.bodyStride, rbk::NPLANES, 32, 24) &&
           make(sys->tileMaps.dxyz, sys->dDxyz, d.atomStride, 3, rbk::kWarpTileAtoms, 3);
}

// Cut the body list into tiles (<= kBlock bodies, <= kMaxTileAtoms atoms unless one body is larger) and
// allocate + fill everything that does not change between uploads.
int allocateDevice(rbk_system* sys, cudaStream_t st) {
    HostModel& h = sys->host;
    DeviceSystem& d = sys->dev;
    int count = 0;
    cudaError_t ce = cudaGetDeviceCount(&count);
    if (ce != cudaSuccess || count == 0)
        return fail(RBK_ECUDA, "librbk needs a CUDA device (there is no CPU fallback)");

    const int nB = h.numBodies, nA = h.numBodyAtoms, nF = h.numFree;
    std::vector<int> loc((size_t) nB + 1, 0);
    std::vector<uint8_t> local(padTo((size_t) std::max(nA, 1) + 4, 16), 0);   // +4: staged in 4-byte granules
    int maxSize = 0;
    for (int b = 0; b < nB; b++) {
        loc[b] = h.body[b].loc;
        maxSize = std::max(maxSize, h.body[b].N);
    }
    loc[nB] = nB ? h.body[nB-1].loc + h.body[nB-1].N : 0;
    // atom tiles (define the per-atom body byte) and body tiles
    auto cut = [&](int atomCap, int bodyCap, bool fillLocal) {
        std::vector<int4> out;
        int first = 0, inTile = 0, atomsInTile = 0;
        for (int b = 0; b < nB; b++) {
            const int n = h.body[b].N;
            if (inTile > 0 && (inTile == bodyCap || atomsInTile + n > atomCap)) {
                out.push_back(make_int4(first, inTile, loc[first], atomsInTile));
                first = b;
                inTile = 0;
                atomsInTile = 0;
            }
            if (fillLocal) for (int j = 0; j < n; j++) local[(size_t) loc[b] + j] = (uint8_t) inTile;
            inTile++;
            atomsInTile += n;
        }
        if (inTile > 0) out.push_back(make_int4(first, inTile, loc[first], atomsInTile));
        return out;
    };
    // large bodies: smaller atom tiles keep a tile's coordinates L1-resident between the two atom phases of part 2
    const bool large = nB > 0 && (long long) nA > (long long) rbk::kSplitAtomsPerBody*nB;
    const int atomCap = large ? rbk::kLargeBodyTileAtoms : rbk::kTileAtoms;
    // (large bodies: at most kBlock - 8 bodies per atom tile - part2LargeKernel keeps a few threads for the tile's plane copies)
    std::vector<int4> meta = cut(atomCap, large ? rbk::kBlock - 8 : rbk::kBlock, true), bodyMeta = cut(rbk::kMaxTileAtoms, rbk::kBlock, false);

    d.numBodies = nB;
    d.numFree = nF;
    d.numBodyAtoms = nA;
    d.numTiles = (int) meta.size();
    d.numBodyTiles = (int) bodyMeta.size();
    d.splitPart1 = nB > 0 && (long long) nA > (long long) rbk::kSplitAtomsPerBody*nB;
    d.fusable = !d.splitPart1;
    // interior steps of step(n) keep F and tau in registers; RBK_EAGER_FORCE_TORQUE=1 restores the stores (A/B measurements)
    const char* eager = std::getenv("RBK_EAGER_FORCE_TORQUE");
    d.lazyForceTorque = !(eager && eager[0] == '1');
    for (const int4& t : meta) if (t.w > rbk::kTileAtoms) d.fusable = 0;
    // one-warp tiles for the step-fused kernel: bodies of <= 4 atoms (water) - then the atom tiles above are exactly the
    // consecutive groups of 128 bodies and these are their 32-body quarters (localBody & 31 = index in the quarter)
    std::vector<int4> warpMeta;
    if (d.fusable && nB > 0 && maxSize <= rbk::kWarpTileAtoms/32)
        for (int b = 0; b < nB; b += 32) {
            const int n = std::min(32, nB - b);
            warpMeta.push_back(make_int4(b, n, loc[b], loc[b + n] - loc[b]));
        }
    d.numWarpTiles = (int) warpMeta.size();
    d.stageBodies = 4;
    if (large)
        for (const int4& t : meta) if (t.w <= rbk::kLargeBodyTileAtoms) d.stageBodies = std::max(d.stageBodies, (t.y + 3) & ~3);
    d.numFreeBlocks = (nF + rbk::kFreePerBlock - 1)/rbk::kFreePerBlock;
    d.rotationMode = h.rotationMode;
    d.maxBodySize = maxSize;
    int device = 0;
    RBK_CUDA(cudaGetDevice(&device));
    RBK_CUDA(cudaDeviceGetAttribute(&d.numSMs, cudaDevAttrMultiProcessorCount, device));
    d.bodyStride = padTo((size_t) std::max(nB, 1), 32);
    d.atomStride = padTo((size_t) std::max(nA, 1), 32);
    d.freeStride = padTo((size_t) std::max(nF, 1), 32);

    RBK_CUDA(devAlloc(sys->dState, d.bodyStride*rbk::NPLANES));
    RBK_CUDA(devAlloc(sys->dDxyz, d.atomStride*3));
    RBK_CUDA(devAlloc(sys->dLocalBody, local.size()));
    loc.resize(loc.size() + 32, loc.back());                  // one-warp tiles copy 32 offsets whatever the tile holds
    RBK_CUDA(devAlloc(sys->dLoc, loc.size()));
    if (meta.empty()) meta.push_back(make_int4(0, 0, 0, 0));
    if (bodyMeta.empty()) bodyMeta.push_back(make_int4(0, 0, 0, 0));
    RBK_CUDA(devAlloc(sys->dTileMeta, meta.size()));
    RBK_CUDA(devAlloc(sys->dBodyTileMeta, bodyMeta.size()));
    if (warpMeta.empty()) warpMeta.push_back(make_int4(0, 0, 0, 0));
    RBK_CUDA(devAlloc(sys->dWarpTileMeta, warpMeta.size()));
    RBK_CUDA(copyAsync(sys->dWarpTileMeta, warpMeta.data(), warpMeta.size()*sizeof(int4), cudaMemcpyHostToDevice, st));
    RBK_CUDA(copyAsync(sys->dTileMeta, meta.data(), meta.size()*sizeof(int4), cudaMemcpyHostToDevice, st));
    RBK_CUDA(copyAsync(sys->dBodyTileMeta, bodyMeta.data(), bodyMeta.size()*sizeof(int4), cudaMemcpyHostToDevice, st));
    RBK_CUDA(devAlloc(sys->dAtomLoc, (size_t) std::max(h.numActualAtoms, 1)));
    RBK_CUDA(devAlloc(sys->dPluginLoc, (size_t) std::max(h.numActualAtoms, 1)));
    RBK_CUDA(devAlloc(sys->dBodyRun, (size_t) std::max(nB, 1)));
    sys->uniformBodySize = maxSize;
    for (int b = 0; b < nB; b++) if (h.body[b].N != maxSize) sys->uniformBodySize = 0;
    sys->bodyOrder.resize(nB); sys->bodyPos.resize(nB); sys->freeOrder.resize(nF); sys->freePos.resize(nF);
    for (int b = 0; b < nB; b++) sys->bodyOrder[b] = sys->bodyPos[b] = b;
    for (int k = 0; k < nF; k++) sys->freeOrder[k] = sys->freePos[k] = k;
    sys->sorted = false;
    RBK_CUDA(devAlloc(sys->dFreeInvMass, d.freeStride));
    RBK_CUDA(devAlloc(sys->dSavedPos, d.freeStride*3));
    std::vector<double> atomMass(d.atomStride, 0.0);
    for (int a = 0; a < nA; a++) atomMass[a] = h.mass[h.atomIndex[(size_t) nF + a]];
    RBK_CUDA(devAlloc(sys->dAtomMass, atomMass.size()));
    RBK_CUDA(devAlloc(sys->dDofSum, 1));
    RBK_CUDA(copyAsync(sys->dAtomMass, atomMass.data(), atomMass.size()*sizeof(double), cudaMemcpyHostToDevice, st));
    std::vector<double> atomInvMass(d.atomStride, 0.0);
    for (int a = 0; a < nA; a++) atomInvMass[a] = atomMass[a] == 0.0 ? 0.0 : 1.0/atomMass[a];     // as CudaContext fills velm.w
    RBK_CUDA(devAlloc(sys->dAtomInvMass, atomInvMass.size()));
    RBK_CUDA(copyAsync(sys->dAtomInvMass, atomInvMass.data(), atomInvMass.size()*sizeof(double), cudaMemcpyHostToDevice, st));
    RBK_CUDA(cudaMemsetAsync(sys->dState, 0, d.bodyStride*rbk::NPLANES*sizeof(double), st));
    RBK_CUDA(devAlloc(sys->dKinPartial, (size_t) 2*rbk::kKineticBlocks));
    RBK_CUDA(devAlloc(sys->dKinCounter, 1));
    RBK_CUDA(devAlloc(sys->dKinOut, 2));
    RBK_CUDA(cudaMallocHost((void**) &sys->hKinOut, 2*sizeof(double)));
    RBK_CUDA(cudaMemsetAsync(sys->dKinCounter, 0, sizeof(unsigned), st));
    RBK_CUDA(cudaMemsetAsync(sys->dSavedPos, 0, d.freeStride*3*sizeof(double), st));
    RBK_CUDA(copyAsync(sys->dLocalBody, local.data(), local.size(), cudaMemcpyHostToDevice, st));
    RBK_CUDA(copyAsync(sys->dLoc, loc.data(), loc.size()*sizeof(int), cudaMemcpyHostToDevice, st));
    if (nF) RBK_CUDA(copyAsync(sys->dFreeInvMass, h.freeInvMass.data(), (size_t) nF*sizeof(double), cudaMemcpyHostToDevice, st));
    RBK_CUDA(cudaStreamSynchronize(st));       // the host vectors above die at scope exit

    d.state = sys->dState;
    d.dxyz = sys->dDxyz;
    d.localBody = sys->dLocalBody;
    d.loc = sys->dLoc;
    d.tileMeta = sys->dTileMeta;
    d.bodyTileMeta = sys->dBodyTileMeta;
    d.warpTileMeta = sys->dWarpTileMeta;
    RBK_CUDA(devAlloc(sys->dTileCounter, 1));
    RBK_CUDA(cudaMemsetAsync(sys->dTileCounter, 0, sizeof(int), st));
    d.tileCounter = sys->dTileCounter;
    RBK_CUDA(devAlloc(sys->dSeriesCtl, 1));
    const rbk::SeriesControl ctl0 = {1, 0u, 0u, 0u, 1};           // start on the middle rung (order 13); the kernels move it
    RBK_CUDA(copyAsync(sys->dSeriesCtl, &ctl0, sizeof(ctl0), cudaMemcpyHostToDevice, st));
    RBK_CUDA(cudaStreamSynchronize(st));
    d.seriesCtl = sys->dSeriesCtl;
    RBK_CUDA(cudaHostAlloc((void**) &sys->hRung, sizeof(int), cudaHostAllocMapped));
    *sys->hRung = ctl0.rung;
    RBK_CUDA(cudaHostGetDevicePointer((void**) &d.hostRungDevice, sys->hRung, 0));
    d.hostRung = sys->hRung;
    const char* full = std::getenv("RBK_FULL_LADDER");
    d.fullLadderOnly = full && full[0] == '1';
    d.tileMaps = d.numWarpTiles > 0 && encodeTileMaps(sys) ? &sys->tileMaps : nullptr;
    d.atomLoc = nullptr;
    d.bodyRun = nullptr;
    d.numSlots = nF + nA;
    const char* noBulk = std::getenv("RBK_NO_BULK_PART2");
    d.noBulkPart2 = noBulk && noBulk[0] == '1';
    const char* noRide = std::getenv("RBK_NO_FREE_RIDE");
    d.noFreeRide = noRide && noRide[0] == '1';
    const char* keepW = std::getenv("RBK_KEEP_VELM_W");        // =1: never write velm.w (partial-sector velocity stores; A/B runs)
    d.atomInvMass = keepW && keepW[0] == '1' ? nullptr : sys->dAtomInvMass;
    d.freeInvMass = sys->dFreeInvMass;
    d.savedPos = sys->dSavedPos;
    sys->allocated = true;
    return RBK_OK;
}

// out[p][s*width + j] = in[p][src[s]*width + j] for every plane p: moves whole bodies (width = atoms per body, or 1)
__global__ void gatherRowsKernel(double* __restrict__ out, const double* __restrict__ in, const int* __restrict__ src, int n, int width,
                                 size_t stride, int planes) {
    const long long i = (long long) blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= (long long) n*width) return;
    const int s = (int) (i/width), j = (int) (i - (long long) s*width);
    const size_t from = (size_t) src[s]*width + j;
    for (int p = 0; p < planes; p++) out[p*stride + i] = in[p*stride + from];
}

// Re-sort one device array (planes x stride doubles) in place through the scratch buffer.
int resortArray(rbk_system* sys, double* data, int n, int width, size_t stride, int planes, cudaStream_t st) {
    if (data == nullptr || n == 0) return RBK_OK;
    const size_t need = stride*planes;
    if (sys->scratchDoubles < need) {
        cudaFree(sys->dScratch);
        sys->dScratch = nullptr;
        sys->scratchDoubles = 0;
        RBK_CUDA(devAlloc(sys->dScratch, need));
        sys->scratchDoubles = need;
    }
    const long long items = (long long) n*width;
    gatherRowsKernel<<<(unsigned) ((items + 255)/256), 256, 0, st>>>(sys->dScratch, data, sys->dGather, n, width, stride, planes);
    RBK_CUDA(cudaGetLastError());
    for (int p = 0; p < planes; p++)                   // (only the entries that exist: the planes' padding stays as it was)
        RBK_CUDA(cudaMemcpyAsync(data + p*stride, sys->dScratch + p*stride, (size_t) items*sizeof(double), cudaMemcpyDeviceToDevice, st));
    return RBK_OK;
}

// Bring the device arrays from the current storage order to (bodyOrder, freeOrder).
int resortStorage(rbk_system* sys, const std::vector<int>& bodyOrder, const std::vector<int>& freeOrder, cudaStream_t st) {
    const DeviceSystem& d = sys->dev;
    const int nB = d.numBodies, nF = d.numFree, N = sys->uniformBodySize;
    if (!sys->dGather) RBK_CUDA(devAlloc(sys->dGather, (size_t) std::max(std::max(nB, nF), 1)));
    std::vector<int> gather;
    if (nB > 0 && bodyOrder != sys->bodyOrder) {
        gather.resize(nB);
        for (int s = 0; s < nB; s++) gather[s] = sys->bodyPos[bodyOrder[s]];          // where the body that goes to s is now
        RBK_CUDA(copyAsync(sys->dGather, gather.data(), (size_t) nB*sizeof(int), cudaMemcpyHostToDevice, st));
        if (int rc = resortArray(sys, sys->dState, nB, 1, d.bodyStride, rbk::NPLANES, st)) return rc;
        if (int rc = resortArray(sys, sys->dDxyz, nB, N, d.atomStride, 3, st)) return rc;
        if (int rc = resortArray(sys, sys->dAtomMass, nB, N, d.atomStride, 1, st)) return rc;
        if (int rc = resortArray(sys, sys->dAtomInvMass, nB, N, d.atomStride, 1, st)) return rc;
        if (int rc = resortArray(sys, sys->refined.rdot, nB, 1, d.bodyStride, 3, st)) return rc;
        if (int rc = resortArray(sys, sys->refined.qdot, nB, 1, d.bodyStride, 4, st)) return rc;
        RBK_CUDA(cudaStreamSynchronize(st));                                            // `gather` is reused below
        sys->bodyOrder = bodyOrder;
        for (int s = 0; s < nB; s++) sys->bodyPos[bodyOrder[s]] = s;
    }
    if (nF > 0 && freeOrder != sys->freeOrder) {
        gather.resize(nF);
        for (int k = 0; k < nF; k++) gather[k] = sys->freePos[freeOrder[k]];
        RBK_CUDA(copyAsync(sys->dGather, gather.data(), (size_t) nF*sizeof(int), cudaMemcpyHostToDevice, st));
        if (int rc = resortArray(sys, sys->dFreeInvMass, nF, 1, d.freeStride, 1, st)) return rc;
        if (int rc = resortArray(sys, sys->dSavedPos, nF, 1, d.freeStride, 3, st)) return rc;
        if (int rc = resortArray(sys, sys->refined.posDot, nF, 1, d.freeStride, 3, st)) return rc;
        RBK_CUDA(cudaStreamSynchronize(st));
        sys->freeOrder = freeOrder;
        for (int k = 0; k < nF; k++) sys->freePos[freeOrder[k]] = k;
    }
    sys->sorted = false;
    for (int s = 0; s < nB && !sys->sorted; s++) sys->sorted = sys->bodyOrder[s] != s;
    for (int k = 0; k < nF && !sys->sorted; k++) sys->sorted = sys->freeOrder[k] != k;
    return RBK_OK;
}

// storage atom index (the order of dxyz, atomLoc, ...) of atom j of plugin body b / of plugin free atom k
inline size_t storageBodyAtom(const rbk_system* sys, int b, int j) {
    return sys->sorted ? (size_t) sys->bodyPos[b]*sys->uniformBodySize + j : (size_t) sys->host.body[b].loc + j;
}

int setLocation(rbk_system* sys, const int* location, cudaStream_t st) {
    const HostModel& h = sys->host;
    const int nF = h.numFree, nB = h.numBodies, N = sys->uniformBodySize;
    const int n = nF + h.numBodyAtoms;               // slots actually addressed by the kernels
    const int* src = location ? location : h.atomIndex.data();
    for (int i = 0; i < n; i++)
        if (src[i] < 0) return fail(RBK_EINVAL, "rbk_set_atom_location: negative location");
    if (n > 0) RBK_CUDA(copyAsync(sys->dPluginLoc, src, (size_t) n*sizeof(int), cudaMemcpyHostToDevice, st));
    // Can the storage order follow the caller's?  Needs bodies of one size whose atoms the caller keeps together, in order.
    bool sortable = nB == 0 || N > 0;
    for (int b = 0; b < nB && sortable; b++) {
        const int* a = src + nF + (size_t) b*N;
        for (int j = 1; j < N && sortable; j++) sortable = a[j] == a[0] + j;
    }
    std::vector<int> bodyOrder(nB), freeOrder(nF);
    for (int b = 0; b < nB; b++) bodyOrder[b] = b;
    for (int k = 0; k < nF; k++) freeOrder[k] = k;
    if (sortable) {
        std::sort(bodyOrder.begin(), bodyOrder.end(), [&](int x, int y) { return src[nF + (size_t) x*N] < src[nF + (size_t) y*N]; });
        std::sort(freeOrder.begin(), freeOrder.end(), [&](int x, int y) { return src[x] < src[y]; });
    }
    if (bodyOrder != sys->bodyOrder || freeOrder != sys->freeOrder)
        if (int rc = resortStorage(sys, bodyOrder, freeOrder, st)) return rc;
    // the table the kernels use: storage atom -> caller slot
    std::vector<int>& table = sys->locationTable;
    table.resize(n);
    for (int k = 0; k < nF; k++) table[k] = src[sys->freeOrder[k]];
    if (sys->sorted)
        for (int s = 0; s < nB; s++) {
            const int* a = src + nF + (size_t) sys->bodyOrder[s]*N;
            for (int j = 0; j < N; j++) table[nF + (size_t) s*N + j] = a[j];
        }
    else for (int i = nF; i < n; i++) table[i] = src[i];
    bool identity = true;
    for (int i = 0; i < n && identity; i++) identity = table[i] == i;
    if (!identity) RBK_CUDA(copyAsync(sys->dAtomLoc, table.data(), (size_t) n*sizeof(int), cudaMemcpyHostToDevice, st));
    // bodies whose atoms the caller keeps in consecutive slots, in order: large-body Part 2 fetches each body's forces as one run
    std::vector<int> run(nB);
    bool runs = sys->dev.splitPart1 != 0;            // (only that kernel uses the table)
    int numSlots = 0;
    for (int i = 0; i < n; i++) numSlots = std::max(numSlots, table[i] + 1);
    for (int s = 0; s < nB && runs; s++) {
        const HostBody& hb = h.body[sys->sorted ? sys->bodyOrder[s] : s];
        const int* a = table.data() + storageBodyAtom(sys, sys->sorted ? sys->bodyOrder[s] : s, 0) + nF;
        run[s] = a[0];
        for (int j = 1; j < hb.N && runs; j++) runs = a[j] == a[0] + j;
    }
    if (runs && nB > 0) RBK_CUDA(copyAsync(sys->dBodyRun, run.data(), (size_t) nB*sizeof(int), cudaMemcpyHostToDevice, st));
    RBK_CUDA(cudaStreamSynchronize(st));             // `location` is the caller's
    sys->dev.atomLoc = identity ? nullptr : sys->dAtomLoc;
    sys->dev.bodyRun = runs && nB > 0 ? sys->dBodyRun : nullptr;
    sys->dev.numSlots = numSlots;
    return RBK_OK;
}

int viewOf(const void* p, int layout, long long stride, AtomView& v) {
    v.p = (double*) p;
    v.fmt = rbk::FMT_F64;
    v.aux = nullptr;
    if (layout == RBK_LAYOUT_VEC3) { v.sa = 3; v.sc = 1; }
    else if (layout == RBK_LAYOUT_SOA) {
        if (stride <= 0) return fail(RBK_EINVAL, "RBK_LAYOUT_SOA needs a positive plane stride");
        v.sa = 1; v.sc = stride;
    }
    else return fail(RBK_EINVAL, "unknown atom layout");
    return RBK_OK;
}

} // namespace

extern "C" {

int rbk_version(void) { return RBK_VERSION; }

int rbk_debug_series_order(rbk_system* sys, int* out, void* stream) {
    if (!sys || !out) return fail(RBK_EINVAL, "rbk_debug_series_order: NULL argument");
    if (!sys->allocated) return fail(RBK_ESTATE, "rbk_debug_series_order: call rbk_upload first");
    const bool bodyTiles = sys->dev.splitPart1 && sys->dev.numTiles > 0;      // large bodies: the rotation kernel's ladder 12 / 13 / 16
    if (sys->dev.numWarpTiles == 0 && !bodyTiles) {            // four-warp atom tiles: the fixed order
        *out = rbk::kSeriesOrder;
        return RBK_OK;
    }
    rbk::SeriesControl ctl;
    RBK_CUDA(cudaMemcpyAsync(&ctl, sys->dSeriesCtl, sizeof(ctl), cudaMemcpyDeviceToHost, (cudaStream_t) stream));
    RBK_CUDA(cudaStreamSynchronize((cudaStream_t) stream));
    const int rung = ctl.rung < 0 ? 0 : (ctl.rung > 2 ? 2 : ctl.rung);
    *out = bodyTiles && rung == 0 ? rbk::kSeriesOrder : rbk::kSeriesLadder[rung];
    return RBK_OK;
}

int rbk_debug_copy_counters(const rbk_system*, long long* out) {
    if (!out) return fail(RBK_EINVAL, "rbk_debug_copy_counters: NULL argument");
    for (int i = 0; i < 4; i++) out[i] = g_copies[i].load();
    return RBK_OK;
}
int rbk_debug_launches_per_call(const rbk_system* sys, int* out) {
    if (!sys || !out) return fail(RBK_EINVAL, "rbk_debug_launches_per_call: NULL argument");
    if (!sys->allocated) return fail(RBK_ESTATE, "rbk_debug_launches_per_call: call rbk_upload first");
    rbk::launchesPerCall(sys->dev, out);
    return RBK_OK;
}
const char* rbk_last_error(void) { return g_error.c_str(); }

int rbk_create(int numAtoms, const int* bodyIndices, const double* masses, const unsigned char* isVirtual,
               int numConstraints, const int* constraintAtoms, int rotationMode, rbk_system** out) {
    if (!out) return fail(RBK_EINVAL, "rbk_create: out is NULL");
    *out = nullptr;
    if (numConstraints < 0 || (numConstraints > 0 && !constraintAtoms)) return fail(RBK_EINVAL, "rbk_create: bad constraint list");
    rbk_system* sys = new (std::nothrow) rbk_system();
    if (!sys) return fail(RBK_ENOMEM, "rbk_create: out of memory");
    std::string err;
    try {
        err = sys->host.initialize(numAtoms, bodyIndices, masses, isVirtual, numConstraints, constraintAtoms, rotationMode);
    }
    catch (const std::exception& e) {
        delete sys;
        return fail(RBK_ENOMEM, std::string("rbk_create: ") + e.what());
    }
    if (!err.empty()) {
        delete sys;
        return fail(err.find("Constraints") == 0 ? RBK_ECONSTRAINT : RBK_EINVAL, err);
    }
    *out = sys;
    return RBK_OK;
}

void rbk_destroy(rbk_system* sys) { delete sys; }

int rbk_get_counts(const rbk_system* sys, int* out) {
    if (!sys || !out) return fail(RBK_EINVAL, "rbk_get_counts: NULL argument");
    const HostModel& h = sys->host;
    out[0] = h.numBodies; out[1] = h.numFree; out[2] = h.numActualAtoms; out[3] = h.numBodyAtoms; out[4] = h.numDOF;
    return RBK_OK;
}

int rbk_get_body_index(const rbk_system* sys, int* out) {
    if (!sys || !out) return fail(RBK_EINVAL, "rbk_get_body_index: NULL argument");
    std::memcpy(out, sys->host.bodyIndex.data(), sys->host.bodyIndex.size()*sizeof(int));
    return RBK_OK;
}

int rbk_get_atom_index(const rbk_system* sys, int* out) {
    if (!sys || !out) return fail(RBK_EINVAL, "rbk_get_atom_index: NULL argument");
    std::memcpy(out, sys->host.atomIndex.data(), sys->host.atomIndex.size()*sizeof(int));
    return RBK_OK;
}

namespace {
// Device body state -> host model (everything a later rbk_upload writes back: r p q pi F tau 1/m 1/I I and the body-frame
// coordinates).  Needed before a velocities-only rebuild on a system that has been stepped or built on the device: the
// host copy still holds the configuration of the last setPositions, and rebuilding pi with that q and uploading it would
// rewind every body (RigidBodyIntegrator::stateChanged(Velocities) after step(), openmmapi/src/RigidBodyIntegrator.cpp:63-74).
int pullHostModel(rbk_system* sys) {
    HostModel& h = sys->host;
    const DeviceSystem& d = sys->dev;
    const size_t ld = d.bodyStride, as = d.atomStride;
    RBK_CUDA(cudaDeviceSynchronize());                 // no stream argument here: whatever was queued must be done
    std::vector<double>& buf = sys->staging;
    buf.resize(std::max(buf.size(), std::max(ld*rbk::NPLANES, as*3)));
    RBK_CUDA(copySync(buf.data(), sys->dState, ld*rbk::NPLANES*sizeof(double), cudaMemcpyDeviceToHost));
    for (int pb = 0; pb < h.numBodies; pb++) {
        HostBody& B = h.body[pb];
        const size_t b = (size_t) sys->bodyPos[pb];
        for (int c = 0; c < 3; c++) {
            B.rcm[c] = buf[(rbk::PL_R + c)*ld + b];
            B.pcm[c] = buf[(rbk::PL_P + c)*ld + b];
            B.force[c] = buf[(rbk::PL_F + c)*ld + b];
            B.tau[c] = buf[(rbk::PL_TAU + c)*ld + b];
            B.I[c] = buf[(rbk::PL_I + c)*ld + b];
            B.invI[c] = buf[(rbk::PL_INVI + c)*ld + b];
        }
        for (int c = 0; c < 4; c++) {
            B.q[c] = buf[(rbk::PL_Q + c)*ld + b];
            B.pi[c] = buf[(rbk::PL_PI + c)*ld + b];
        }
        B.invMass = buf[rbk::PL_INVM*ld + b];
        B.mass = 1.0/B.invMass;
        const double* q = B.q;                          // C(q) tau, the form the reference stores (RigidBody.h:40)
        B.torque[0] = -q[1]*B.tau[0] - q[2]*B.tau[1] - q[3]*B.tau[2];
        B.torque[1] =  q[0]*B.tau[0] + q[3]*B.tau[1] - q[2]*B.tau[2];
        B.torque[2] = -q[3]*B.tau[0] + q[0]*B.tau[1] + q[1]*B.tau[2];
        B.torque[3] =  q[2]*B.tau[0] - q[1]*B.tau[1] + q[0]*B.tau[2];
    }
    if (sys->hostStale) {                               // geometry was built on the device: the coordinates too
        RBK_CUDA(copySync(buf.data(), sys->dDxyz, as*3*sizeof(double), cudaMemcpyDeviceToHost));
        for (int pb = 0; pb < h.numBodies; pb++)
            for (int j = 0; j < h.body[pb].N; j++) {
                const size_t a = storageBodyAtom(sys, pb, j), ha = (size_t) h.body[pb].loc + j;
                for (int c = 0; c < 3; c++) h.d[3*ha + c] = buf[c*as + a];
            }
    }
    sys->hostStale = false;
    sys->deviceAhead = false;
    return RBK_OK;
}
} // namespace

int rbk_update(rbk_system* sys, const double* R, const double* V, const double* F, int geometry, int velocities) {
    if (!sys) return fail(RBK_EINVAL, "rbk_update: NULL system");
    if (geometry && (!R || !F)) return fail(RBK_EINVAL, "rbk_update: geometry needs positions and forces");
    if (velocities && !V) return fail(RBK_EINVAL, "rbk_update: velocities needed");
    if (velocities && !geometry && sys->host.numBodies > 0 && sys->host.body[0].mass == 0.0 && !sys->hostStale)
        return fail(RBK_ESTATE, "rbk_update: velocities before any geometry build");
    if (!geometry && sys->uploaded && (sys->deviceAhead || sys->hostStale)) {
        if (sys->forceTorqueStale)
            return fail(RBK_ESTATE, "rbk_update: velocities cannot be rebuilt between rbk_part2_part1 and the closing rbk_part2");
        if (int rc = pullHostModel(sys)) return rc;
    }
    sys->host.update(R, V, F, geometry != 0, velocities != 0);
    if (geometry) sys->hostStale = false;
    return RBK_OK;
}

int rbk_get_host_bodies(const rbk_system* sys, int* N, int* dof, int* loc, double* mass, double* I, double* invI,
                        double* rcm, double* pcm, double* q, double* pi, double* force, double* torque, double* twoK) {
    if (!sys) return fail(RBK_EINVAL, "rbk_get_host_bodies: NULL system");
    if (sys->hostStale) return fail(RBK_ESTATE, "rbk_get_host_bodies: bodies were built on the device (rbk_update_device); use rbk_download_bodies");
    const HostModel& h = sys->host;
    for (int b = 0; b < h.numBodies; b++) {
        const HostBody& B = h.body[b];
        if (N) N[b] = B.N;
        if (dof) dof[b] = B.dof;
        if (loc) loc[b] = B.loc;
        if (mass) mass[b] = B.mass;
        for (int c = 0; c < 3; c++) {
            if (I) I[3*b+c] = B.I[c];
            if (invI) invI[3*b+c] = B.invI[c];
            if (rcm) rcm[3*b+c] = B.rcm[c];
            if (pcm) pcm[3*b+c] = B.pcm[c];
            if (force) force[3*b+c] = B.force[c];
        }
        for (int c = 0; c < 4; c++) {
            if (q) q[4*b+c] = B.q[c];
            if (pi) pi[4*b+c] = B.pi[c];
            if (torque) torque[4*b+c] = B.torque[c];
        }
        if (twoK) { twoK[2*b] = B.twoKt; twoK[2*b+1] = B.twoKr; }
    }
    return RBK_OK;
}

int rbk_get_body_fixed(const rbk_system* sys, double* d) {
    if (!sys || !d) return fail(RBK_EINVAL, "rbk_get_body_fixed: NULL argument");
    std::memcpy(d, sys->host.d.data(), sys->host.d.size()*sizeof(double));
    return RBK_OK;
}

int rbk_upload(rbk_system* sys, void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_upload: NULL system");
    if (sys->hostStale)
        return fail(RBK_ESTATE, "rbk_upload: the bodies were built on the device (rbk_update_device); the host copy is not current");
    cudaStream_t st = (cudaStream_t) stream;
    if (!sys->allocated) {
        int rc = allocateDevice(sys, st);
        if (rc != RBK_OK) return rc;
        rc = setLocation(sys, nullptr, st);
        if (rc != RBK_OK) return rc;
    }
    RBK_CUDA(cudaMemsetAsync(sys->dTileCounter, 0, sizeof(int), st));     // self-resetting per launch; re-armed here in case one failed
    const HostModel& h = sys->host;
    const DeviceSystem& d = sys->dev;
    const size_t ld = d.bodyStride;
    std::vector<double>& buf = sys->staging;
    buf.assign(std::max(ld*rbk::NPLANES, d.atomStride*3), 0.0);
    for (int pb = 0; pb < h.numBodies; pb++) {
        const HostBody& B = h.body[pb];
        const size_t b = (size_t) sys->bodyPos[pb];           // storage position of plugin body pb
        for (int c = 0; c < 3; c++) {
            buf[(rbk::PL_R + c)*ld + b] = B.rcm[c];
            buf[(rbk::PL_P + c)*ld + b] = B.pcm[c];
            buf[(rbk::PL_F + c)*ld + b] = B.force[c];
            buf[(rbk::PL_TAU + c)*ld + b] = B.tau[c];
            buf[(rbk::PL_I + c)*ld + b] = B.I[c];
            buf[(rbk::PL_INVI + c)*ld + b] = B.invI[c];
        }
        for (int c = 0; c < 4; c++) {
            buf[(rbk::PL_Q + c)*ld + b] = B.q[c];
            buf[(rbk::PL_PI + c)*ld + b] = B.pi[c];
        }
        buf[rbk::PL_INVM*ld + b] = B.invMass;
    }
    RBK_CUDA(copyAsync(sys->dState, buf.data(), ld*rbk::NPLANES*sizeof(double), cudaMemcpyHostToDevice, st));
    RBK_CUDA(cudaStreamSynchronize(st));
    const size_t as = d.atomStride;
    for (int pb = 0; pb < h.numBodies; pb++)
        for (int j = 0; j < h.body[pb].N; j++) {
            const size_t a = storageBodyAtom(sys, pb, j), ha = (size_t) h.body[pb].loc + j;
            for (int c = 0; c < 3; c++) buf[c*as + a] = h.d[3*ha + c];
        }
    RBK_CUDA(copyAsync(sys->dDxyz, buf.data(), as*3*sizeof(double), cudaMemcpyHostToDevice, st));
    RBK_CUDA(cudaStreamSynchronize(st));
    sys->uploaded = true;
    sys->mirrorsLoaded = false;
    sys->deviceAhead = false;
    sys->forceTorqueStale = false;
    return RBK_OK;
}

namespace {
int updateDevice(rbk_system* sys, AtomView p, AtomView v, AtomView f, int geometry, int velocities, cudaStream_t st) {
    if (!sys->allocated) {
        int rc = allocateDevice(sys, st);
        if (rc != RBK_OK) return rc;
        rc = setLocation(sys, nullptr, st);
        if (rc != RBK_OK) return rc;
    }
    if (velocities && !geometry && !sys->uploaded) return fail(RBK_ESTATE, "rbk_update_device: velocities before any geometry build");
    RBK_CUDA(rbk::launchBuild(sys->dev, sys->dAtomMass, p, v, f, sys->dDxyz, geometry != 0, velocities != 0, sys->dDofSum, st));
    if (geometry) {
        int dofSum = 0;
        RBK_CUDA(copyAsync(&dofSum, sys->dDofSum, sizeof(int), cudaMemcpyDeviceToHost, st));
        RBK_CUDA(cudaStreamSynchronize(st));
        sys->host.numDOF = sys->host.numFree - sys->host.numConstraints + dofSum;      // RigidBodySystem.cpp:130-134
    }
    sys->uploaded = true;
    sys->hostStale = true;
    sys->mirrorsLoaded = false;
    return RBK_OK;
}
} // namespace

int rbk_update_device(rbk_system* sys, const double* pos, const double* vel, const double* force, int layout,
                      long long stride, int geometry, int velocities, void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_update_device: NULL system");
    if (geometry && (!pos || !force)) return fail(RBK_EINVAL, "rbk_update_device: geometry needs positions and forces");
    if (velocities && !vel) return fail(RBK_EINVAL, "rbk_update_device: velocities needed");
    AtomView p, v, f;
    if (viewOf(pos, layout, stride, p) || viewOf(vel, layout, stride, v) || viewOf(force, layout, stride, f)) return RBK_EINVAL;
    return updateDevice(sys, p, v, f, geometry, velocities, (cudaStream_t) stream);
}

int rbk_set_atom_location(rbk_system* sys, const int* location, void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_set_atom_location: NULL system");
    if (!sys->allocated) {                           // a caller that builds on the device never uploads: allocate here
        if (int rc = allocateDevice(sys, (cudaStream_t) stream)) return rc;
    }
    return setLocation(sys, location, (cudaStream_t) stream);
}

namespace {
// The handle's second stream (large-body systems with free atoms only; created on first use, NULL when it cannot be)
const rbk::SideStream* sideStream(rbk_system* sys) {
    if (!(sys->dev.splitPart1 && sys->dev.numFree > 0 && sys->dev.numTiles > 0)) return nullptr;
    static const bool disabled = [] { const char* e = std::getenv("RBK_NO_SIDE_STREAM"); return e && e[0] == '1'; }();
    if (disabled) return nullptr;                            // A/B measurements: free atoms in the caller's stream, before the body kernels
    if (!sys->side.stream) {
        rbk::SideStream s{nullptr, nullptr, nullptr};
        if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) {
            if (s.stream) cudaStreamDestroy(s.stream);
            if (s.fork) cudaEventDestroy(s.fork);
            if (s.join) cudaEventDestroy(s.join);
            cudaGetLastError();
            return nullptr;                                  // the kernels then run in one stream
        }
        sys->side = s;
    }
    return &sys->side;
}

// Part 1 / Part 2 with the refined-energy passes around them when the diagnostics are on (rbk_refined.cu)
cudaError_t stepPart1(rbk_system* sys, double dt, AtomView p, AtomView v, AtomView f, const AtomView* delta, cudaStream_t st) {
    sys->deviceAhead = true;
    if (sys->refinedMode != RBK_REFINED_OFF) {
        cudaError_t e = rbk::launchRefinedBodies(sys->dev, sys->refined, dt, 1, st);
        if (e == cudaSuccess && sys->refinedMode == RBK_REFINED_ALL && !delta)
            e = rbk::launchRefinedFree(sys->dev, sys->refined, dt, 1, v, f, st);
        if (e != cudaSuccess) return e;
    }
    return delta ? rbk::launchPart1Delta(sys->dev, dt, p, v, f, *delta, st) : rbk::launchPart1(sys->dev, dt, p, v, f, st, sideStream(sys));
}

cudaError_t stepPart2(rbk_system* sys, double dt, AtomView p, AtomView v, AtomView f, cudaStream_t st) {
    sys->deviceAhead = true;
    sys->forceTorqueStale = false;
    cudaError_t e = rbk::launchPart2(sys->dev, dt, p, v, f, st, sideStream(sys));
    if (e != cudaSuccess || sys->refinedMode == RBK_REFINED_OFF) return e;
    e = rbk::launchRefinedBodies(sys->dev, sys->refined, dt, 2, st);
    if (e == cudaSuccess && sys->refinedMode == RBK_REFINED_ALL) e = rbk::launchRefinedFree(sys->dev, sys->refined, dt, 2, v, f, st);
    return e;
}

int stepPart2Part1(rbk_system* sys, double dt, AtomView p, AtomView v, AtomView f, cudaStream_t st) {
    if (sys->refinedMode != RBK_REFINED_OFF) {          // the diagnostics sit between the two halves: no one-pass kernel
        RBK_CUDA(stepPart2(sys, dt, p, v, f, st));
        RBK_CUDA(stepPart1(sys, dt, p, v, f, nullptr, st));
        return RBK_OK;
    }
    sys->deviceAhead = true;
    RBK_CUDA(rbk::launchPart2Part1(sys->dev, dt, p, v, f, st, sideStream(sys)));
    sys->forceTorqueStale = !rbk::part2Part1LeavesForceTorque(sys->dev);
    return RBK_OK;
}

int reduceOut(rbk_system* sys, double* out, int count, cudaStream_t st) {
    RBK_CUDA(copyAsync(sys->hKinOut, sys->dKinOut, 2*sizeof(double), cudaMemcpyDeviceToHost, st));
    RBK_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < count; i++) out[i] = sys->hKinOut[i];
    return RBK_OK;
}

int needRefined(rbk_system* sys, const char* who) {
    if (!sys) return fail(RBK_EINVAL, std::string(who) + ": NULL system");
    if (!sys->uploaded) return fail(RBK_ESTATE, std::string(who) + ": body system not uploaded");
    if (sys->refinedMode == RBK_REFINED_OFF) return fail(RBK_ESTATE, std::string(who) + ": refined energies are not enabled");
    return RBK_OK;
}
} // namespace

int rbk_set_refined_energies(rbk_system* sys, int mode, void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_set_refined_energies: NULL system");
    if (mode != RBK_REFINED_OFF && mode != RBK_REFINED_ALL && mode != RBK_REFINED_BODIES)
        return fail(RBK_EINVAL, "rbk_set_refined_energies: unknown mode");
    if (mode != RBK_REFINED_OFF) {
        if (!sys->allocated) return fail(RBK_ESTATE, "rbk_set_refined_energies: call rbk_upload first");
        if (!sys->refined.rdot) {
            RBK_CUDA(devAlloc(sys->refined.rdot, sys->dev.bodyStride*3));
            RBK_CUDA(devAlloc(sys->refined.qdot, sys->dev.bodyStride*4));
            RBK_CUDA(devAlloc(sys->refined.posDot, sys->dev.freeStride*3));
        }
        cudaStream_t st = (cudaStream_t) stream;
        RBK_CUDA(cudaMemsetAsync(sys->refined.rdot, 0, sys->dev.bodyStride*3*sizeof(double), st));
        RBK_CUDA(cudaMemsetAsync(sys->refined.qdot, 0, sys->dev.bodyStride*4*sizeof(double), st));
        RBK_CUDA(cudaMemsetAsync(sys->refined.posDot, 0, sys->dev.freeStride*3*sizeof(double), st));
    }
    sys->refinedMode = mode;
    return RBK_OK;
}

int rbk_refined_kinetic(rbk_system* sys, double dt, const double* vel, int layout, long long stride, double* out, void* stream) {
    if (int rc = needRefined(sys, "rbk_refined_kinetic")) return rc;
    if (!out) return fail(RBK_EINVAL, "rbk_refined_kinetic: NULL argument");
    AtomView v;
    if (viewOf(vel, layout, stride, v)) return RBK_EINVAL;
    cudaStream_t st = (cudaStream_t) stream;
    RBK_CUDA(rbk::launchRefinedKinetic(sys->dev, sys->refined, dt, v, sys->dKinPartial, sys->dKinCounter, sys->dKinOut, st));
    return reduceOut(sys, out, 2, st);
}

int rbk_potential_refinement(rbk_system* sys, double dt, const double* force, int layout, long long stride, double* out,
                             void* stream) {
    if (int rc = needRefined(sys, "rbk_potential_refinement")) return rc;
    if (!out) return fail(RBK_EINVAL, "rbk_potential_refinement: NULL argument");
    AtomView f;
    if (viewOf(force, layout, stride, f)) return RBK_EINVAL;
    cudaStream_t st = (cudaStream_t) stream;
    RBK_CUDA(rbk::launchPotentialRefinement(sys->dev, sys->refined, dt, f, sys->dKinPartial, sys->dKinCounter, sys->dKinOut, st));
    return reduceOut(sys, out, 1, st);
}

int rbk_part1(rbk_system* sys, double dt, double* pos, double* vel, const double* force, int layout, long long stride,
              void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_part1: NULL system");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_part1: body system not uploaded");
    AtomView p, v, f;
    if (viewOf(pos, layout, stride, p) || viewOf(vel, layout, stride, v) || viewOf(force, layout, stride, f)) return RBK_EINVAL;
    RBK_CUDA(stepPart1(sys, dt, p, v, f, nullptr, (cudaStream_t) stream));
    return RBK_OK;
}

int rbk_part2(rbk_system* sys, double dt, const double* pos, double* vel, const double* force, int layout,
              long long stride, void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_part2: NULL system");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_part2: body system not uploaded");
    AtomView p, v, f;
    if (viewOf(pos, layout, stride, p) || viewOf(vel, layout, stride, v) || viewOf(force, layout, stride, f)) return RBK_EINVAL;
    RBK_CUDA(stepPart2(sys, dt, p, v, f, (cudaStream_t) stream));
    return RBK_OK;
}

int rbk_part2_part1(rbk_system* sys, double dt, double* pos, double* vel, const double* force, int layout,
                    long long stride, void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_part2_part1: NULL system");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_part2_part1: body system not uploaded");
    AtomView p, v, f;
    if (viewOf(pos, layout, stride, p) || viewOf(vel, layout, stride, v) || viewOf(force, layout, stride, f)) return RBK_EINVAL;
    return stepPart2Part1(sys, dt, p, v, f, (cudaStream_t) stream);
}

int rbk_kinetic(rbk_system* sys, const double* vel, int layout, long long stride, double* out, void* stream) {
    if (!sys || !out) return fail(RBK_EINVAL, "rbk_kinetic: NULL argument");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_kinetic: body system not uploaded");
    cudaStream_t st = (cudaStream_t) stream;
    AtomView v;
    if (viewOf(vel, layout, stride, v)) return RBK_EINVAL;
    RBK_CUDA(rbk::launchKinetic(sys->dev, v, sys->dKinPartial, sys->dKinCounter, sys->dKinOut, st));
    RBK_CUDA(copyAsync(sys->hKinOut, sys->dKinOut, 2*sizeof(double), cudaMemcpyDeviceToHost, st));
    RBK_CUDA(cudaStreamSynchronize(st));
    out[0] = sys->hKinOut[0];
    out[1] = sys->hKinOut[1];
    return RBK_OK;
}

namespace {
int openmmViews(void* posq, void* posqCorrection, void* velm, const long long* force, int paddedNumAtoms, int precision,
                AtomView& p, AtomView& v, AtomView& f) {
    if (paddedNumAtoms <= 0) return fail(RBK_EINVAL, "OpenMM layout: paddedNumAtoms must be positive");
    p = AtomView{(double*) posq, 0, 0, 0, nullptr};
    v = AtomView{(double*) velm, 0, 0, 0, nullptr};
    f = AtomView{(double*) force, 1, (long long) paddedNumAtoms, rbk::FMT_FORCE_FIXED, nullptr};
    if (precision == RBK_OPENMM_SINGLE) { p.fmt = rbk::FMT_REAL4_F32; v.fmt = rbk::FMT_REAL4_F32; }
    else if (precision == RBK_OPENMM_MIXED) {
        if (!posqCorrection) return fail(RBK_EINVAL, "OpenMM mixed precision needs posqCorrection");
        p.fmt = rbk::FMT_POSQ_MIXED; p.aux = posqCorrection; v.fmt = rbk::FMT_REAL4_F64;
    }
    else if (precision == RBK_OPENMM_DOUBLE) { p.fmt = rbk::FMT_REAL4_F64; v.fmt = rbk::FMT_REAL4_F64; }
    else return fail(RBK_EINVAL, "unknown OpenMM precision");
    return RBK_OK;
}
} // namespace

int rbk_part1_openmm(rbk_system* sys, double dt, void* posq, void* posqCorrection, void* velm, const long long* force,
                     int paddedNumAtoms, int precision, void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_part1_openmm: NULL system");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_part1_openmm: body system not uploaded");
    AtomView p, v, f;
    if (openmmViews(posq, posqCorrection, velm, force, paddedNumAtoms, precision, p, v, f)) return RBK_EINVAL;
    RBK_CUDA(stepPart1(sys, dt, p, v, f, nullptr, (cudaStream_t) stream));
    return RBK_OK;
}

int rbk_part2_openmm(rbk_system* sys, double dt, void* posq, void* posqCorrection, void* velm, const long long* force,
                     int paddedNumAtoms, int precision, void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_part2_openmm: NULL system");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_part2_openmm: body system not uploaded");
    AtomView p, v, f;
    if (openmmViews(posq, posqCorrection, velm, force, paddedNumAtoms, precision, p, v, f)) return RBK_EINVAL;
    RBK_CUDA(stepPart2(sys, dt, p, v, f, (cudaStream_t) stream));
    return RBK_OK;
}

int rbk_part2_part1_openmm(rbk_system* sys, double dt, void* posq, void* posqCorrection, void* velm, const long long* force,
                           int paddedNumAtoms, int precision, void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_part2_part1_openmm: NULL system");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_part2_part1_openmm: body system not uploaded");
    AtomView p, v, f;
    if (openmmViews(posq, posqCorrection, velm, force, paddedNumAtoms, precision, p, v, f)) return RBK_EINVAL;
    return stepPart2Part1(sys, dt, p, v, f, (cudaStream_t) stream);
}

namespace {
// forces of the atoms the integrator owns: planes (stride) -> packed [3][n] in plugin order, and back through a new map
__global__ void gatherForcesKernel(const long long* __restrict__ force, long long stride, const int* __restrict__ loc, int n,
                                   long long* __restrict__ packed) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long a = loc ? loc[i] : i;
#pragma unroll
    for (int c = 0; c < 3; c++) packed[(size_t) c*n + i] = force[a + c*stride];
}
__global__ void scatterForcesKernel(long long* __restrict__ force, long long stride, const int* __restrict__ loc, int n,
                                    const long long* __restrict__ packed) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long a = loc ? loc[i] : i;
#pragma unroll
    for (int c = 0; c < 3; c++) force[a + c*stride] = packed[(size_t) c*n + i];
}
} // namespace

int rbk_reorder_openmm(rbk_system* sys, const int* location, long long* force, int paddedNumAtoms, void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_reorder_openmm: NULL system");
    if (!sys->allocated) return fail(RBK_ESTATE, "rbk_reorder_openmm: call rbk_upload first");
    cudaStream_t st = (cudaStream_t) stream;
    const int n = sys->host.numFree + sys->host.numBodyAtoms;
    if (force && n > 0) {
        if (paddedNumAtoms <= 0) return fail(RBK_EINVAL, "rbk_reorder_openmm: paddedNumAtoms must be positive");
        if (!sys->dForcePacked) RBK_CUDA(devAlloc(sys->dForcePacked, (size_t) 3*n));
        gatherForcesKernel<<<(n + 255)/256, 256, 0, st>>>(force, paddedNumAtoms, sys->dPluginLoc, n, sys->dForcePacked);
        RBK_CUDA(cudaGetLastError());
    }
    if (int rc = setLocation(sys, location, st)) return rc;
    if (force && n > 0) {
        scatterForcesKernel<<<(n + 255)/256, 256, 0, st>>>(force, paddedNumAtoms, sys->dPluginLoc, n, sys->dForcePacked);
        RBK_CUDA(cudaGetLastError());
    }
    return RBK_OK;
}

int rbk_free_delta_openmm(rbk_system* sys, double dt, const void* velm, const long long* force, int paddedNumAtoms,
                          int precision, void* posDelta, void* stream) {
    if (!sys || !posDelta) return fail(RBK_EINVAL, "rbk_free_delta_openmm: NULL argument");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_free_delta_openmm: body system not uploaded");
    AtomView p, v, f;
    int dummy = 0;
    if (openmmViews(&dummy, &dummy, (void*) velm, force, paddedNumAtoms, precision, p, v, f)) return RBK_EINVAL;
    const AtomView d{(double*) posDelta, 0, 0, v.fmt, nullptr};
    RBK_CUDA(rbk::launchFreeDelta(sys->dev, dt, v, f, d, (cudaStream_t) stream));
    return RBK_OK;
}

int rbk_part1_delta_openmm(rbk_system* sys, double dt, void* posq, void* posqCorrection, void* velm,
                           const long long* force, int paddedNumAtoms, int precision, const void* posDelta, void* stream) {
    if (!sys || !posDelta) return fail(RBK_EINVAL, "rbk_part1_delta_openmm: NULL argument");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_part1_delta_openmm: body system not uploaded");
    AtomView p, v, f;
    if (openmmViews(posq, posqCorrection, velm, force, paddedNumAtoms, precision, p, v, f)) return RBK_EINVAL;
    const AtomView d{(double*) posDelta, 0, 0, v.fmt, nullptr};
    RBK_CUDA(stepPart1(sys, dt, p, v, f, &d, (cudaStream_t) stream));
    return RBK_OK;
}

int rbk_update_device_openmm(rbk_system* sys, void* posq, void* posqCorrection, void* velm, const long long* force,
                             int paddedNumAtoms, int precision, int geometry, int velocities, void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_update_device_openmm: NULL system");
    AtomView p, v, f;
    if (openmmViews(posq, posqCorrection, velm, force, paddedNumAtoms, precision, p, v, f)) return RBK_EINVAL;
    return updateDevice(sys, p, v, f, geometry, velocities, (cudaStream_t) stream);
}

int rbk_kinetic_openmm(rbk_system* sys, const void* velm, int precision, double* out, void* stream) {
    if (!sys || !out) return fail(RBK_EINVAL, "rbk_kinetic_openmm: NULL argument");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_kinetic_openmm: body system not uploaded");
    cudaStream_t st = (cudaStream_t) stream;
    AtomView v{(double*) velm, 0, 0, precision == RBK_OPENMM_SINGLE ? rbk::FMT_REAL4_F32 : rbk::FMT_REAL4_F64, nullptr};
    RBK_CUDA(rbk::launchKinetic(sys->dev, v, sys->dKinPartial, sys->dKinCounter, sys->dKinOut, st));
    RBK_CUDA(copyAsync(sys->hKinOut, sys->dKinOut, 2*sizeof(double), cudaMemcpyDeviceToHost, st));
    RBK_CUDA(cudaStreamSynchronize(st));
    out[0] = sys->hKinOut[0];
    out[1] = sys->hKinOut[1];
    return RBK_OK;
}

int rbk_free_dot_openmm(rbk_system* sys, const void* posDelta, int precision, double factor, int restart, void* stream) {
    if (int rc = needRefined(sys, "rbk_free_dot_openmm")) return rc;
    if (!posDelta) return fail(RBK_EINVAL, "rbk_free_dot_openmm: NULL argument");
    const AtomView d{(double*) posDelta, 0, 0, precision == RBK_OPENMM_SINGLE ? rbk::FMT_REAL4_F32 : rbk::FMT_REAL4_F64, nullptr};
    RBK_CUDA(rbk::launchFreeDot(sys->dev, sys->refined, d, factor, restart != 0, (cudaStream_t) stream));
    return RBK_OK;
}

int rbk_refined_kinetic_openmm(rbk_system* sys, double dt, const void* velm, int precision, double* out, void* stream) {
    if (int rc = needRefined(sys, "rbk_refined_kinetic_openmm")) return rc;
    if (!out) return fail(RBK_EINVAL, "rbk_refined_kinetic_openmm: NULL argument");
    cudaStream_t st = (cudaStream_t) stream;
    const AtomView v{(double*) velm, 0, 0, precision == RBK_OPENMM_SINGLE ? rbk::FMT_REAL4_F32 : rbk::FMT_REAL4_F64, nullptr};
    RBK_CUDA(rbk::launchRefinedKinetic(sys->dev, sys->refined, dt, v, sys->dKinPartial, sys->dKinCounter, sys->dKinOut, st));
    return reduceOut(sys, out, 2, st);
}

int rbk_potential_refinement_openmm(rbk_system* sys, double dt, const long long* force, int paddedNumAtoms, double* out,
                                    void* stream) {
    if (int rc = needRefined(sys, "rbk_potential_refinement_openmm")) return rc;
    if (!out || paddedNumAtoms <= 0) return fail(RBK_EINVAL, "rbk_potential_refinement_openmm: bad argument");
    cudaStream_t st = (cudaStream_t) stream;
    const AtomView f{(double*) force, 1, (long long) paddedNumAtoms, rbk::FMT_FORCE_FIXED, nullptr};
    RBK_CUDA(rbk::launchPotentialRefinement(sys->dev, sys->refined, dt, f, sys->dKinPartial, sys->dKinCounter, sys->dKinOut, st));
    return reduceOut(sys, out, 1, st);
}

// host-buffer variants: after rbk_execute_host the handle's device mirrors hold the current V and F
int rbk_refined_kinetic_host(rbk_system* sys, double dt, const double* V, double* out, void* stream) {
    if (int rc = needRefined(sys, "rbk_refined_kinetic_host")) return rc;
    if (!V || !out) return fail(RBK_EINVAL, "rbk_refined_kinetic_host: NULL argument");
    if (!sys->mVel) return fail(RBK_ESTATE, "rbk_refined_kinetic_host: no step has been taken with rbk_execute_host");
    cudaStream_t st = (cudaStream_t) stream;
    const size_t bytes = (size_t) sys->host.numAtoms*3*sizeof(double);
    if (sys->host.numFree > 0 && !sys->hostVelStale) RBK_CUDA(copyAsync(sys->mVel, V, bytes, cudaMemcpyHostToDevice, st));
    return rbk_refined_kinetic(sys, dt, sys->mVel, RBK_LAYOUT_VEC3, 0, out, stream);
}

int rbk_potential_refinement_host(rbk_system* sys, double dt, const double* F, double* out, void* stream) {
    if (int rc = needRefined(sys, "rbk_potential_refinement_host")) return rc;
    if (!F || !out) return fail(RBK_EINVAL, "rbk_potential_refinement_host: NULL argument");
    if (!sys->mForce) return fail(RBK_ESTATE, "rbk_potential_refinement_host: no step has been taken with rbk_execute_host");
    cudaStream_t st = (cudaStream_t) stream;
    const size_t bytes = (size_t) sys->host.numAtoms*3*sizeof(double);
    if (sys->host.numFree > 0) RBK_CUDA(copyAsync(sys->mForce, F, bytes, cudaMemcpyHostToDevice, st));
    return rbk_potential_refinement(sys, dt, sys->mForce, RBK_LAYOUT_VEC3, 0, out, stream);
}

int rbk_kinetic_host(rbk_system* sys, const double* V, double* out, void* stream) {
    if (!sys || !V || !out) return fail(RBK_EINVAL, "rbk_kinetic_host: NULL argument");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_kinetic_host: body system not uploaded");
    cudaStream_t st = (cudaStream_t) stream;
    const size_t bytes = (size_t) sys->host.numAtoms*3*sizeof(double);
    if (!sys->mVel) {
        RBK_CUDA(cudaMalloc((void**) &sys->mPos, bytes));
        RBK_CUDA(cudaMalloc((void**) &sys->mVel, bytes));
        RBK_CUDA(cudaMalloc((void**) &sys->mForce, bytes));
    }
    // (after an rbk_execute_host call that left the velocities on the device, V == NULL, the mirror is the current copy)
    if (sys->host.numFree > 0 && !sys->hostVelStale) RBK_CUDA(copyAsync(sys->mVel, V, bytes, cudaMemcpyHostToDevice, st));
    return rbk_kinetic(sys, sys->mVel, RBK_LAYOUT_VEC3, 0, out, stream);
}

int rbk_download_bodies(rbk_system* sys, double* rcm, double* pcm, double* q, double* pi, double* force,
                        double* torque, void* stream) {
    if (!sys) return fail(RBK_EINVAL, "rbk_download_bodies: NULL system");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_download_bodies: body system not uploaded");
    if ((force || torque) && sys->forceTorqueStale)
        return fail(RBK_ESTATE, "rbk_download_bodies: force and torque are not kept between rbk_part2_part1 and the closing rbk_part2");
    cudaStream_t st = (cudaStream_t) stream;
    const size_t ld = sys->dev.bodyStride;
    std::vector<double>& buf = sys->staging;
    buf.resize(std::max(buf.size(), ld*rbk::NPLANES));
    RBK_CUDA(copyAsync(buf.data(), sys->dState, ld*rbk::NPLANES*sizeof(double), cudaMemcpyDeviceToHost, st));
    RBK_CUDA(cudaStreamSynchronize(st));
    for (int b = 0; b < sys->host.numBodies; b++) {              // b: plugin order (the caller's); sb: where it is stored
        const size_t sb = (size_t) sys->bodyPos[b];
        double qq[4], tt[3];
        for (int c = 0; c < 3; c++) {
            if (rcm) rcm[3*b+c] = buf[(rbk::PL_R + c)*ld + sb];
            if (pcm) pcm[3*b+c] = buf[(rbk::PL_P + c)*ld + sb];
            if (force) force[3*b+c] = buf[(rbk::PL_F + c)*ld + sb];
            tt[c] = buf[(rbk::PL_TAU + c)*ld + sb];
        }
        for (int c = 0; c < 4; c++) {
            qq[c] = buf[(rbk::PL_Q + c)*ld + sb];
            if (q) q[4*b+c] = qq[c];
            if (pi) pi[4*b+c] = buf[(rbk::PL_PI + c)*ld + sb];
        }
        if (torque) {                                   // C(q) tau, the form the reference stores
            torque[4*b+0] = -qq[1]*tt[0] - qq[2]*tt[1] - qq[3]*tt[2];
            torque[4*b+1] =  qq[0]*tt[0] + qq[3]*tt[1] - qq[2]*tt[2];
            torque[4*b+2] = -qq[3]*tt[0] + qq[0]*tt[1] + qq[1]*tt[2];
            torque[4*b+3] =  qq[2]*tt[0] - qq[1]*tt[1] + qq[0]*tt[2];
        }
    }
    return RBK_OK;
}

int rbk_execute_host(rbk_system* sys, double dt, int steps, double* R, double* V, double* F, rbk_force_fn forces,
                     void* user, void* stream) {
    return rbk_execute_host_hooks(sys, dt, steps, R, V, F, forces, nullptr, nullptr, user, stream);
}

namespace {
int executeHost(rbk_system* sys, double dt, int steps, double* R, double* V, double* F, rbk_force_fn forces,
                rbk_positions_fn constrainPositions, rbk_velocities_fn constrainVelocities, void* user, cudaStream_t st) {
    const size_t bytes = (size_t) sys->host.numAtoms*3*sizeof(double);
    if (!sys->mPos) {
        RBK_CUDA(cudaMalloc((void**) &sys->mPos, bytes));
        RBK_CUDA(cudaMalloc((void**) &sys->mVel, bytes));
        RBK_CUDA(cudaMalloc((void**) &sys->mForce, bytes));
    }
    if (!sys->mForce2) {
        RBK_CUDA(cudaMalloc((void**) &sys->mForce2, bytes));
        RBK_CUDA(cudaStreamCreateWithFlags(&sys->h2dStream, cudaStreamNonBlocking));
        RBK_CUDA(cudaStreamCreateWithFlags(&sys->d2hStream, cudaStreamNonBlocking));
        for (cudaEvent_t* e : {&sys->evStart, &sys->evForces, &sys->evPart1, &sys->evPositions})
            RBK_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    const bool hostFlow = forces || constrainPositions;       // something on the host needs R between Part 1 and Part 2
    if (!sys->mirrorsLoaded) {
        if (!V) return fail(RBK_EINVAL, "rbk_execute_host: the first call after an upload needs V (the device mirror is empty)");
        RBK_CUDA(copyAsync(sys->mPos, R, bytes, cudaMemcpyHostToDevice, st));
        RBK_CUDA(copyAsync(sys->mVel, V, bytes, cudaMemcpyHostToDevice, st));
        RBK_CUDA(copyAsync(sys->mForce, F, bytes, cudaMemcpyHostToDevice, st));
        sys->mirrorsLoaded = true;
        sys->hostVelStale = false;
    }
    else if (hostFlow)
        // F holds the forces at the current positions and the caller may have re-evaluated them since the last call
        // (updateParametersInContext, setParameter): the reference reads data.forces every step
        // (ReferenceRigidBodyKernels.cpp:96), so Part 1's half kick must not use a stale mirror
        RBK_CUDA(copyAsync(sys->mForce, F, bytes, cudaMemcpyHostToDevice, st));
    const AtomView p{sys->mPos, 3, 1, rbk::FMT_F64, nullptr}, v{sys->mVel, 3, 1, rbk::FMT_F64, nullptr};
    // step(n) = part1, forces, [part2+part1 in one pass, forces] x (n-1), part2 whenever nothing sits between Part 2 of
    // one step and Part 1 of the next (no velocity hook, no diagnostics): rbk_part2_part1
    const bool fuse = steps > 1 && !constrainVelocities && sys->refinedMode == RBK_REFINED_OFF;
    for (int i = 0; i < steps; i++) {
        const bool last = i == steps - 1;
        // new forces land in the second mirror while Part 1 still reads the old ones; without a host force evaluation the
        // call's F is uploaded once (step 0) and serves every later step of the call
        const bool newForces = hostFlow || i == 0;
        const AtomView fOld{sys->mForce, 3, 1, rbk::FMT_F64, nullptr}, fNew{newForces ? sys->mForce2 : sys->mForce, 3, 1, rbk::FMT_F64, nullptr};
        if (hostFlow) {
            // forces depend on the new positions: Part 1 -> positions to the host -> hooks -> forces to the device
            if (constrainPositions) {
                // the solver needs the old positions of the atoms it may move - the free atoms; copying those alone keeps
                // this O(numFree) instead of a host memcpy of every position per step (rigid-body atoms are outputs of the
                // body state and never constrained, RigidBodySystem.cpp:107-113)
                sys->oldPositions.resize((size_t) sys->host.numAtoms*3);
                const int* freeAtom = sys->host.atomIndex.data();
                for (int k = 0; k < sys->host.numFree; k++) {
                    const size_t a = 3*(size_t) freeAtom[k];
                    sys->oldPositions[a] = R[a]; sys->oldPositions[a + 1] = R[a + 1]; sys->oldPositions[a + 2] = R[a + 2];
                }
            }
            if (i == 0 || !fuse) RBK_CUDA(stepPart1(sys, dt, p, v, fOld, nullptr, st));
            RBK_CUDA(copyAsync(R, sys->mPos, bytes, cudaMemcpyDeviceToHost, st));
            RBK_CUDA(cudaStreamSynchronize(st));
            if (constrainPositions && constrainPositions(sys->oldPositions.data(), R, sys->host.numAtoms, user))
                RBK_CUDA(copyAsync(sys->mPos, R, bytes, cudaMemcpyHostToDevice, st));
            if (forces) forces(R, F, sys->host.numAtoms, user);
            RBK_CUDA(copyAsync(sys->mForce2, F, bytes, cudaMemcpyHostToDevice, st));
        }
        else if (i == 0) {
            // F is this call's input, known up front and the same for all of its steps: the upload (H2D engine) runs under
            // Part 1 - and, in a one-step call, under the download of the positions (D2H engine); Part 2 waits for it only
            RBK_CUDA(cudaEventRecord(sys->evStart, st));
            RBK_CUDA(cudaStreamWaitEvent(sys->h2dStream, sys->evStart, 0));
            RBK_CUDA(copyAsync(sys->mForce2, F, bytes, cudaMemcpyHostToDevice, sys->h2dStream));
            RBK_CUDA(cudaEventRecord(sys->evForces, sys->h2dStream));
            RBK_CUDA(stepPart1(sys, dt, p, v, fOld, nullptr, st));
            if (last) {
                RBK_CUDA(cudaEventRecord(sys->evPart1, st));
                RBK_CUDA(cudaStreamWaitEvent(sys->d2hStream, sys->evPart1, 0));
                RBK_CUDA(copyAsync(R, sys->mPos, bytes, cudaMemcpyDeviceToHost, sys->d2hStream));
                RBK_CUDA(cudaEventRecord(sys->evPositions, sys->d2hStream));
            }
            RBK_CUDA(cudaStreamWaitEvent(st, sys->evForces, 0));
        }
        else if (!fuse) RBK_CUDA(stepPart1(sys, dt, p, v, fOld, nullptr, st));
        if (fuse && !last) {
            if (int rc = stepPart2Part1(sys, dt, p, v, fNew, st)) return rc;
            if (!hostFlow && i == steps - 2) {                // positions are final after the last Part 1
                RBK_CUDA(cudaEventRecord(sys->evPart1, st));
                RBK_CUDA(cudaStreamWaitEvent(sys->d2hStream, sys->evPart1, 0));
                RBK_CUDA(copyAsync(R, sys->mPos, bytes, cudaMemcpyDeviceToHost, sys->d2hStream));
                RBK_CUDA(cudaEventRecord(sys->evPositions, sys->d2hStream));
            }
        }
        else {
            if (!hostFlow && !fuse && last && steps > 1) {
                RBK_CUDA(cudaEventRecord(sys->evPart1, st));
                RBK_CUDA(cudaStreamWaitEvent(sys->d2hStream, sys->evPart1, 0));
                RBK_CUDA(copyAsync(R, sys->mPos, bytes, cudaMemcpyDeviceToHost, sys->d2hStream));
                RBK_CUDA(cudaEventRecord(sys->evPositions, sys->d2hStream));
            }
            RBK_CUDA(stepPart2(sys, dt, p, v, fNew, st));
        }
        if (constrainVelocities) {
            RBK_CUDA(copyAsync(V, sys->mVel, bytes, cudaMemcpyDeviceToHost, st));
            RBK_CUDA(cudaStreamSynchronize(st));
            if (constrainVelocities(R, V, sys->host.numAtoms, user))
                RBK_CUDA(copyAsync(sys->mVel, V, bytes, cudaMemcpyHostToDevice, st));
        }
        if (newForces) std::swap(sys->mForce, sys->mForce2);             // mForce = the forces of the latest evaluation
    }
    // velocities leave the device once per call (nothing on the host reads them between the steps of a call unless a
    // velocity hook is installed), and not at all when the caller passes V = NULL
    if (V && !constrainVelocities) RBK_CUDA(copyAsync(V, sys->mVel, bytes, cudaMemcpyDeviceToHost, st));
    sys->hostVelStale = V == nullptr;
    if (!hostFlow) RBK_CUDA(cudaStreamWaitEvent(st, sys->evPositions, 0));     // R is complete when `st` is
    RBK_CUDA(cudaStreamSynchronize(st));
    return RBK_OK;
}
} // namespace

int rbk_execute_host_hooks(rbk_system* sys, double dt, int steps, double* R, double* V, double* F, rbk_force_fn forces,
                           rbk_positions_fn constrainPositions, rbk_velocities_fn constrainVelocities, void* user,
                           void* stream) {
    if (!sys || !R || !F) return fail(RBK_EINVAL, "rbk_execute_host: NULL argument");
    if (!V && constrainVelocities) return fail(RBK_EINVAL, "rbk_execute_host: a velocity hook needs V");
    if (!sys->uploaded) return fail(RBK_ESTATE, "rbk_execute_host: body system not uploaded");
    if (steps <= 0) return RBK_OK;
    cudaStream_t st = (cudaStream_t) stream;
    const int rc = executeHost(sys, dt, steps, R, V, F, forces, constrainPositions, constrainVelocities, user, st);
    if (rc != RBK_OK) {
        // leave nothing in flight that still reads or writes the caller's buffers; the mirrors are reloaded next time
        const std::string message = g_error;
        cudaStreamSynchronize(st);
        if (sys->h2dStream) cudaStreamSynchronize(sys->h2dStream);
        if (sys->d2hStream) cudaStreamSynchronize(sys->d2hStream);
        cudaGetLastError();
        sys->mirrorsLoaded = false;
        g_error = message;
    }
    return rc;
}

} // extern "C"
