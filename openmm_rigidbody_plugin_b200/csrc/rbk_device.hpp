// rbk_device.hpp - device-side data layout of librbk and the kernel launchers.
//
// HBM layout (all fp64 unless noted; see DESIGN.md "Data layout"):
//   body state : NPLANES structure-of-arrays planes of bodyStride doubles each
//                r[3] p[3] q[4] pi[4] F[3] tau[3] invm I[3] invI[3]          (27 planes)
//   body atoms : dxyz = 3 planes of atomStride doubles (body-frame coordinates, body-major order),
//                localBody = 1 byte per body atom (index of its body inside its tile)
//   maps       : loc[nB+1] prefix offsets, tileBody[nTiles+1], atomLoc[numActualAtoms] (or NULL = identity)
//   free atoms : freeInvMass[nF], savedPos = 3 planes of freeStride doubles
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace rbk {

enum Plane : int {
    PL_R = 0, PL_P = 3, PL_Q = 6, PL_PI = 10, PL_F = 14, PL_TAU = 17, PL_INVM = 20, PL_I = 21, PL_INVI = 24, NPLANES = 27
};

constexpr int kBlock = 128;            // threads per CTA = max bodies per tile
constexpr int kTileAtoms = 768;        // soft cap of body atoms per tile = staged d capacity (a single larger body gets its own tile)
constexpr int kFreePerBlock = 512;     // free atoms per CTA (4 per thread)

struct DeviceSystem {
    int numBodies, numFree, numBodyAtoms, numTiles, numFreeBlocks;
    int rotationMode, maxBodySize, numSMs;
    size_t bodyStride, atomStride, freeStride;
    double* state;
    const double* dxyz;
    const uint8_t* localBody;
    const int* loc;
    const int* tileBody;
    const int4* tileMeta;        // per tile: first body, #bodies, first body-atom, #atoms
    const int* atomLoc;
    const double* freeInvMass;
    double* savedPos;
};

// Caller-owned atom arrays: element (atom i, component c) lives at p[i*sa + c*sc].
struct AtomView {
    double* p;
    long long sa, sc;
};

cudaError_t launchPart1(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st);
cudaError_t launchPart2(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st);
// partial: scratch of 2*kKineticBlocks doubles; counter: zero-initialised unsigned; out: 2 doubles (device)
constexpr int kKineticBlocks = 592;    // 148 SMs x 4
cudaError_t launchKinetic(const DeviceSystem& S, AtomView vel, double* partial, unsigned* counter, double* out, cudaStream_t st);

} // namespace rbk
