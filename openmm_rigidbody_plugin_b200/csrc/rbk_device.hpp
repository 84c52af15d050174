// rbk_device.hpp - device-side data layout of librbk and the kernel launchers.
//
// HBM layout (all fp64 unless noted; see DESIGN.md "Data layout"):
//   body state : NPLANES structure-of-arrays planes of bodyStride doubles each
//                r[3] p[3] q[4] pi[4] invm invI[3] F[3] tau[3] I[3]          (27 planes)
//   body atoms : dxyz = 3 planes of atomStride doubles (body-frame coordinates, body-major order),
//                localBody = 1 byte per body atom (index of its body inside its tile)
//   maps       : loc[nB+1] prefix offsets, tile descriptors int4[nTiles], atomLoc[numActualAtoms] (or NULL = identity)
//   free atoms : freeInvMass[nF], savedPos = 3 planes of freeStride doubles
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace rbk {

// Plane order: what the step-fused kernel stages comes first (r p q pi 1/m 1/I = planes 0..17), then what Part 1 alone
// needs on top (F tau = 18..23), so that ONE 2-D TMA box (rows = planes, columns = a tile's bodies) fetches either set.
enum Plane : int {
    PL_R = 0, PL_P = 3, PL_Q = 6, PL_PI = 10, PL_INVM = 14, PL_INVI = 15, PL_F = 18, PL_TAU = 21, PL_I = 24, NPLANES = 27
};

constexpr int kBlock = 128;            // threads per CTA = max bodies per tile
// Two tilings of the body list.  "Atom tiles" (<=128 bodies, <=kTileAtoms atoms, a larger single body
// alone) drive every thread-per-atom phase: many CTAs, coordinates staged in shared memory.  "Body tiles"
// (<=128 bodies, <=kMaxTileAtoms atoms) drive the stand-alone rotation kernel used when bodies are large,
// so that its thread-per-body phase runs full warps.  For small bodies (water) the two coincide.
constexpr int kTileAtoms = 512;
#ifndef RBK_LARGE_PER_THREAD
#define RBK_LARGE_PER_THREAD 3         // (config 4, Part 2 alone: 256-atom tiles 127 us, 384 112 us, 512 120 us)
#endif
constexpr int kLargePerThread = RBK_LARGE_PER_THREAD;          // atoms per thread of the large-body atom kernels
constexpr int kLargeBodyTileAtoms = kBlock*kLargePerThread;    // atom-tile cap when bodies are large: a tile's forces and
                                                               // arms are staged in shared memory by part2LargeKernel
constexpr int kMaxTileAtoms = 8192;
constexpr int kSplitAtomsPerBody = 8;  // mean body size above which part 1 runs as rotation kernel + atom kernel
#ifndef RBK_WARP_TILE_ATOMS
#define RBK_WARP_TILE_ATOMS 128
#endif
constexpr int kWarpTileAtoms = RBK_WARP_TILE_ATOMS;    // atom capacity of the one-warp tiles (32 bodies of <= 4 atoms) of the step-fused kernel
constexpr int kFreePerBlock = 512;     // free atoms per CTA (4 per thread)

struct TileMaps;

// Device-resident control block of the exact rotation's series ladder (rbk_math.cuh, exactRotationLadder): the hot water
// kernels read `rung` at launch, count the bodies whose truncation check failed at this rung / would have failed one rung
// lower, and the last CTA of the launch moves the rung for the next launch.  No host involvement: works under CUDA graphs,
// and the sequence of rungs is a deterministic function of the trajectory.
struct SeriesControl {
    int rung;
    unsigned done, fails, lower;
    int published;               // the value last written to the host's copy of the rung (mapped pinned memory): written again only
                                 // when it changes - a store to host memory at the end of a kernel costs ~3 us of PCIe round trip
};

struct DeviceSystem {
    int numBodies, numFree, numBodyAtoms, numTiles, numBodyTiles, numFreeBlocks;
    int rotationMode, maxBodySize, numSMs, splitPart1;
    int fusable;                 // every atom tile fits the shared-memory staging of the step-fused kernel
    int lazyForceTorque;         // rbk_part2_part1 keeps F and tau in registers (nothing reads the planes before the next Part 2)
    int numWarpTiles;            // > 0: bodies have <= 4 atoms and the step-fused kernel runs one warp per 32-body tile
    int stageBodies;             // large-body systems: most bodies in any atom tile of <= kLargeBodyTileAtoms atoms (multiple of 4)
    size_t bodyStride, atomStride, freeStride;
    double* state;
    const double* dxyz;
    const uint8_t* localBody;
    const int* loc;
    const int4* tileMeta;        // per atom tile: first body, #bodies, first body-atom, #atoms
    const int4* bodyTileMeta;    // per body tile, same fields
    const int4* warpTileMeta;    // per one-warp tile (subdivision of the atom tiles), same fields
    const TileMaps* tileMaps;    // HOST pointer (kernel-parameter copies are made at launch); NULL = no TMA tensor path
    const int* atomLoc;
    const int* bodyRun;          // per body (storage order): caller slot of its first atom when EVERY body's atoms sit in consecutive
                                 // slots, in order (NULL otherwise) - large-body Part 2 then moves whole runs with TMA bulk copies
    int numSlots;                // 1 + the largest caller slot any atom of the system occupies (the arrays are at least this long)
    int noFreeRide;              // RBK_NO_FREE_RIDE=1: the free atoms of large-body systems always get their own launch (A/B measurements, tests)
    int noBulkPart2;             // RBK_NO_BULK_PART2=1: large-body Part 2 always takes the per-atom request kernel (A/B measurements, tests)
    SeriesControl* seriesCtl;
    const volatile int* hostRung;    // HOST pointer: the rung as the kernels last published it (mapped pinned memory) ...
    int* hostRungDevice;             // ... and the device alias they write it through
    int fullLadderOnly;              // RBK_FULL_LADDER=1: never use the lean rung-0 kernel (A/B measurements, tests)
    int* tileCounter;            // zero between launches: tiles claimed so far by the persistent step-fused kernel
    const double* atomInvMass;   // body atoms, storage order: 1/m as OpenMM stores it in velm.w (NULL: velm.w is never written)
    const double* freeInvMass;
    double* savedPos;
};

// Caller-owned atom arrays.  fmt selects how atom i's three components are stored:
//   FMT_F64       double, component c at p[i*sa + c*sc]           (RBK_LAYOUT_VEC3: sa=3, sc=1; RBK_LAYOUT_SOA: sa=1, sc=stride)
//   FMT_POSQ_MIXED  OpenMM-CUDA mixed precision positions: float4 posq[i] + float4 posqCorrection[i] (aux); .w untouched
//   FMT_REAL4_F64 double4 per atom (OpenMM posq in double precision, velm in mixed/double); .w untouched
//   FMT_REAL4_F32 float4 per atom (OpenMM single precision posq / velm); .w untouched
//   FMT_FORCE_FIXED OpenMM-CUDA forces: long long planes x[sc] y[sc] z[sc], fixed point with scale 2^32 (read only)
enum AtomFormat : int { FMT_F64 = 0, FMT_POSQ_MIXED = 1, FMT_REAL4_F64 = 2, FMT_REAL4_F32 = 3, FMT_FORCE_FIXED = 4 };
struct AtomView {
    double* p;
    long long sa, sc;
    int fmt;
    void* aux;
};

struct SideStream;
cudaError_t launchPart1(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st,
                        const SideStream* side = nullptr);
cudaError_t launchPart2(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st,
                        const SideStream* side = nullptr);
// free-atom constraint hooks: delta = (v + f invm dt/2) dt for every free atom; part 1 that advances free atoms by delta
cudaError_t launchFreeDelta(const DeviceSystem& S, double dt, AtomView vel, AtomView force, AtomView delta, cudaStream_t st);
cudaError_t launchPart1Delta(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, AtomView delta, cudaStream_t st);
// A second stream of the system's own plus the two events that fork it from / join it to the caller's stream: large-body
// steps integrate the free atoms there, next to the body kernels (disjoint atoms, no data dependence).
struct SideStream {
    cudaStream_t stream;
    cudaEvent_t fork, join;
};
// part 2 of one step immediately followed by part 1 of the next (identical results, one pass over the data)
cudaError_t launchPart2Part1(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st,
                             const SideStream* side = nullptr);
// partial: scratch of 2*kKineticBlocks doubles; counter: zero-initialised unsigned; out: 2 doubles (device)
constexpr int kKineticBlocks = 592;    // 148 SMs x 4
// GPU-side body build (rbk_build.cu): geometry and/or dynamics of every body from the caller's atom arrays.
cudaError_t launchBuild(const DeviceSystem& S, const double* atomMass, AtomView pos, AtomView vel, AtomView force,
                        double* dxyz, bool geometry, bool velocities, int* dofSum, cudaStream_t st);
// Refined ("shadow") energy diagnostics (rbk_refined.cu): rdot = 3 planes, qdot = 4 planes of bodyStride doubles,
// posDot = 3 planes of freeStride doubles.  phase 1 = before Part 1, phase 2 = after Part 2.
// Opaque storage for a CUtensorMap (TMA descriptor, 128 bytes, 64-byte aligned); encoded on the host in rbk_api.cu.
struct alignas(64) TensorMapBlob {
    unsigned long long opaque[16];
};
// Descriptors of the one-warp-tile pipeline: state planes as a 2-D tensor [NPLANES][bodyStride] with boxes of
// 32 bodies x 18 planes (step-fused kernel) or x 24 planes (Part 1 alone), body-frame coordinates as [3][atomStride]
// with a box of kWarpTileAtoms atoms x 3 planes.
struct TileMaps {
    TensorMapBlob state18, state24, dxyz;
};

struct RefinedState {
    double* rdot;
    double* qdot;
    double* posDot;
};
cudaError_t launchRefinedBodies(const DeviceSystem& S, const RefinedState& X, double dt, int phase, cudaStream_t st);
cudaError_t launchRefinedFree(const DeviceSystem& S, const RefinedState& X, double dt, int phase, AtomView vel, AtomView force,
                              cudaStream_t st);
cudaError_t launchFreeDot(const DeviceSystem& S, const RefinedState& X, AtomView delta, double factor, bool restart, cudaStream_t st);
cudaError_t launchRefinedKinetic(const DeviceSystem& S, const RefinedState& X, double dt, AtomView vel, double* partial,
                                 unsigned* counter, double* out, cudaStream_t st);
cudaError_t launchPotentialRefinement(const DeviceSystem& S, const RefinedState& X, double dt, AtomView force, double* partial,
                                      unsigned* counter, double* out, cudaStream_t st);
// false when launchPart2Part1 skips the F / tau stores of its bodies (S.lazyForceTorque and the one-pass kernel is used)
bool part2Part1LeavesForceTorque(const DeviceSystem& S);
int part1LaunchesPerStep(const DeviceSystem& S);   // 1 (fused) or 2 (rotation kernel + atom kernel)
void launchesPerCall(const DeviceSystem& S, int out[3]);   // kernel launches of launchPart1, launchPart2, launchPart2Part1
cudaError_t launchKinetic(const DeviceSystem& S, AtomView vel, double* partial, unsigned* counter, double* out, cudaStream_t st);

} // namespace rbk
