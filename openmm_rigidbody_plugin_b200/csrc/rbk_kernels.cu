// rbk_kernels.cu - hand-written CUDA kernels (sm_100a) of the RigidBodyIntegrator step.
//
// Work decomposition (DESIGN.md "Kernels"): bodies are cut into TILES of <=128 consecutive bodies
// (<= 512 atoms, <= 256 for large-body systems); a CTA owns a tile and alternates between two thread mappings
//   thread-per-BODY : coalesced SoA loads of the body state, half kick / rotation / second kick
//   thread-per-ATOM : coalesced loads of body-frame coordinates + atom forces, position / velocity
//                     reconstruction, segmented reduction of force and torque
// exchanging per-body quantities (q, r, v_cm, omega) through shared memory, so the rotation update
// and the atom scatter are ONE kernel and nothing per-body is re-read from HBM.  The reduction is bucketed by body
// size: bodies of <= 8 atoms are summed by their own thread (step-fused kernel), mixed small systems by a warp-shuffle
// segmented scan (part2Kernel), systems of large bodies by a lane group per body (part2LargeKernel).  Free atoms
// (velocity Verlet) have their own kernel.  No atomics on the data path (the persistent kernels' tile counter only
// decides which CTA takes a tile): results are bit-reproducible.
//
// Reference behaviour reproduced: RigidBodySystem::integratePart1/2, computeKineticEnergies
// (openmmapi/src/RigidBodySystem.cpp:170-220); this is NOT a port of platforms/cuda/src/kernels/*.cu
// (one thread per body, AoS, serial atom loops, host-side energy sums).
#include "rbk_atomio.cuh"
#include "rbk_device.hpp"
#include "rbk_step.cuh"

#include <cstddef>
#include <mutex>

namespace rbk {
namespace {

constexpr int kWarps = kBlock/32;
constexpr int kMaxDevices = 64;
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------------
// Part 1: half kick, drift, rotation, position reconstruction (+ free atoms: half kick, drift)
//
// Persistent CTAs (grid = SMs x resident CTAs) walk a tile list.  The 24 body-state planes of the
// NEXT tile are brought into shared memory with cp.async (LDGSTS) while the current tile's rotation
// update runs (double buffer, descriptors two tiles ahead); no value loaded from HBM is ever held in a
// register across the long fp64 phase.
//   FUSED  (small bodies, e.g. water): walks the ATOM tiles; the tile's body-frame coordinates are
//          requested at the start of the tile and consumed in its thread-per-atom phase, so rotation
//          update and atom scatter are one kernel and r, q never go back through HBM.
//   !FUSED (mean body size > kSplitAtomsPerBody): walks the BODY tiles (full warps in the rotation
//          phase) and leaves the positions to atomPositionKernel, which runs with full occupancy;
//          re-reading r, q costs 56 B per body, amortised over the body's many atoms.
// ------------------------------------------------------------------------------------------------
// Resident CTAs per SM.  Exact mode: the order-16 series keeps ~70 doubles live, 236 registers without
// spills -> 2 CTAs (3 CTAs at 168 registers spill and measured slower).  NO-SQUISH needs ~100 -> 3 CTAs
// (then shared memory, 66 KB per CTA, is the limit).
#ifndef RBK_P1_MINBLOCKS_EXACT
#define RBK_P1_MINBLOCKS_EXACT 2
#endif
#ifndef RBK_P1_MINBLOCKS_ROTATION
#define RBK_P1_MINBLOCKS_ROTATION 2     // the body-tile rotation kernel of large-body systems (RUNG >= 0)
#endif
#ifndef RBK_P1_MINBLOCKS_SPLIT
#define RBK_P1_MINBLOCKS_SPLIT 3
#endif
constexpr int kP1Planes = 24;                       // r3 p3 q4 pi4 F3 tau3 invm invI3

struct Part1Smem {
    double body[2][kP1Planes][kBlock];
    int4 meta[3];                                   // ring: descriptors of the current, next and next-but-one tile
    int tileIdx[3];                                 // ring: their tile numbers (claimed dynamically after the first wave)
    double d[3][kTileAtoms];                        // FUSED only (the !FUSED kernel allocates up to here)
    unsigned char localBody[kTileAtoms + 16];
};

__device__ __forceinline__ void cpAsync8(void* smem, const void* gmem) {
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpAsync4(void* smem, const void* gmem) {
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpAsync16(void* smem, const void* gmem) {
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpWait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Per-device launch configuration of one kernel instantiation (dynamic shared memory attribute, resident CTAs per SM).
// Handles on different host threads may launch concurrently: the cache is guarded by a mutex (taken once per launch,
// uncontended in the single-threaded case).
struct LaunchCache {
    std::mutex lock;
    size_t attribute[kMaxDevices] = {};             // largest dynamic shared memory size the kernel was configured for
    size_t occupancyFor[kMaxDevices] = {};          // the size perSM was computed for
    int perSM[kMaxDevices] = {};
    // returns resident CTAs per SM (>= 1) for `kernel` with `smem` bytes of dynamic shared memory, or an error
    template <class Kernel> cudaError_t get(Kernel kernel, int threads, size_t smem, int& blocks) {
        int device = 0;
        cudaError_t e = cudaGetDevice(&device);
        if (e != cudaSuccess) return e;
        const bool cached = device >= 0 && device < kMaxDevices;
        std::lock_guard<std::mutex> guard(lock);
        if (!cached || attribute[device] < smem) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
            if (e != cudaSuccess) return e;
            if (cached) attribute[device] = smem;
        }
        if (!cached || perSM[device] == 0 || occupancyFor[device] != smem) {
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, kernel, threads, smem);
            if (e != cudaSuccess) return e;
            if (blocks < 1) blocks = 1;
            if (cached) { occupancyFor[device] = smem; perSM[device] = blocks; }
            return cudaSuccess;
        }
        blocks = perSM[device];
        return cudaSuccess;
    }
};

// A persistent launch that fails leaves the tile counter wherever the CTAs that did run put it: re-arm it.
inline cudaError_t launchResult(const DeviceSystem& S, cudaStream_t st) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) cudaMemsetAsync(S.tileCounter, 0, sizeof(int), st);
    return e;
}

// smem plane k of a part1Kernel stage (r p q pi | F tau | 1/m | 1/I) <-> global state plane
__device__ __forceinline__ int globalPlane(int k) {
    return k < 14 ? k : (k < 17 ? (int) PL_F + (k - 14) : (k < 20 ? (int) PL_TAU + (k - 17) : (k == 20 ? (int) PL_INVM : (int) PL_INVI + (k - 21))));
}

// RUNG >= 0 (exact rotation, body-tile kernel of large-body systems): ONE rung of a series ladder of orders 12 / 13 / 16
// compiled in, picked by the launcher from the host's copy of the rung (see part2Part1Kernel, LADDER); the kernel counts the
// bodies whose check failed at its order / would fail one rung lower and its last CTA moves the rung.  The lowest rung is the
// fixed order of the other four-warp kernels (12, not the water kernels' 11): in a 128-thread tile a body on the retry path
// holds up its whole CTA, so failures cost more than a lower order saves.  RUNG < 0: fixed order, no control block.
template <bool EXACT, bool FUSED, bool NATIVE, int RUNG = -1>
__global__ void __launch_bounds__(kBlock, EXACT ? (RUNG >= 0 ? RBK_P1_MINBLOCKS_ROTATION : RBK_P1_MINBLOCKS_EXACT) : RBK_P1_MINBLOCKS_SPLIT)
part1Kernel(const DeviceSystem S, const double dt, const AtomView pos, const AtomView vel, const AtomView force) {
    static_assert(RUNG < 0 || (EXACT && !FUSED), "the ladder variant is the exact-rotation body-tile kernel");
    unsigned ladderFails = 0u, ladderLower = 0u;               // bodies of this thread (RUNG >= 0)
    extern __shared__ __align__(128) unsigned char smemRaw[];
    Part1Smem& sm = *reinterpret_cast<Part1Smem*>(smemRaw);
    const int tid = threadIdx.x;
    const int G = gridDim.x;
    const size_t ld = S.bodyStride, as = S.atomStride;
    const int4* __restrict__ tiles = FUSED ? S.tileMeta : S.bodyTileMeta;
    const int numTiles = FUSED ? S.numTiles : S.numBodyTiles;

    auto requestBody = [&](int4 m, int stage) {
        if (tid < m.y) {
            const double* g = S.state + (size_t) (m.x + tid);
#pragma unroll
            for (int k = 0; k < kP1Planes; k++) cpAsync8(&sm.body[stage][k][tid], g + globalPlane(k)*ld);
        }
    };
    auto requestAtoms = [&](int4 m) {
        if (m.w > kTileAtoms) return;
        for (int j = tid; j < m.w; j += kBlock) {
            const double* g = S.dxyz + (size_t) (m.z + j);
            cpAsync8(&sm.d[0][j], g);
            cpAsync8(&sm.d[1][j], g + as);
            cpAsync8(&sm.d[2][j], g + 2*as);
        }
        const int first = m.z & ~3;                            // 4-byte granules of the byte array
        for (int w = tid; 4*w < m.z + m.w - first; w += kBlock)
            cpAsync4(&sm.localBody[4*w], S.localBody + first + 4*w);
    };

    // the first G tiles are the CTAs' own, later ones are claimed from the global counter (see part2Part1Kernel)
    auto claim = [&]() {
        const int old = atomicAdd(S.tileCounter, 1);
        if (old == numTiles - 1) *S.tileCounter = 0;
        return G + old;
    };
    const int tile0 = blockIdx.x;
    if (tile0 < numTiles) {
        if (tid == 0) {
            const int tile1 = claim();
            sm.tileIdx[0] = tile0;
            sm.tileIdx[1] = tile1;
            sm.meta[0] = tiles[tile0];
            if (tile1 < numTiles) sm.meta[1] = tiles[tile1];
        }
        __syncthreads();
        requestBody(sm.meta[0], 0);
        cpCommit();
        for (int it = 0; sm.tileIdx[it % 3] < numTiles; it++) {
            const int stage = it & 1;
            const int4 m = sm.meta[it % 3];
            if (FUSED) requestAtoms(m);
            cpCommit();
            cpWait<1>();                                       // body state of this tile (+ descriptor of the next) landed
            __syncthreads();
            if (sm.tileIdx[(it + 1) % 3] < numTiles) {
                requestBody(sm.meta[(it + 1) % 3], stage ^ 1);
                if (tid == 0) {
                    const int after = claim();
                    sm.tileIdx[(it + 2) % 3] = after;
                    if (after < numTiles) cpAsync16(&sm.meta[(it + 2) % 3], tiles + after);
                }
            }
            cpCommit();

            double (*B)[kBlock] = sm.body[stage];
            if (tid < m.y) {                                   // ---- thread per body
                d3 r = {B[0][tid], B[1][tid], B[2][tid]};
                d3 p = {B[3][tid], B[4][tid], B[5][tid]};
                d4 q = {B[6][tid], B[7][tid], B[8][tid], B[9][tid]};
                d4 pi = {B[10][tid], B[11][tid], B[12][tid], B[13][tid]};
                const d3 F = {B[14][tid], B[15][tid], B[16][tid]};
                const d3 tau = {B[17][tid], B[18][tid], B[19][tid]};
                const double invm = B[20][tid];
                const d3 invI = {B[21][tid], B[22][tid], B[23][tid]};
                if (RUNG >= 0) {
                    unsigned flags = 0u;
                    p = p + F*(0.5*dt);
                    pi = pi + quatC(q, tau)*dt;
                    r = r + p*(invm*dt);
                    if (RUNG == 0) exactRotationRung<kSeriesOrder, 0>(dt, invI, q, pi, flags);
                    else if (RUNG == 1) exactRotationRung<13, kSeriesOrder>(dt, invI, q, pi, flags);
                    else exactRotationRung<16, 13>(dt, invI, q, pi, flags);
                    ladderFails += flags & 1u;
                    ladderLower += (flags >> 1) & 1u;
                }
                else bodyPart1<EXACT>(dt, S.rotationMode, F, tau, invm, invI, r, p, q, pi);
                double* s = S.state + (size_t) (m.x + tid);
                storePlane3(s + PL_R*ld, ld, r);
                storePlane3(s + PL_P*ld, ld, p);
                storePlane4(s + PL_Q*ld, ld, q);
                storePlane4(s + PL_PI*ld, ld, pi);
                if (FUSED) {                                   // hand r, q to the atom phase
                    B[0][tid] = r.x; B[1][tid] = r.y; B[2][tid] = r.z;
                    B[6][tid] = q.w; B[7][tid] = q.x; B[8][tid] = q.y; B[9][tid] = q.z;
                }
            }
            if (FUSED) {
                cpWait<1>();                                   // this tile's coordinates have landed
                __syncthreads();
                if (m.w <= kTileAtoms) {                       // ---- thread per atom
                    const int shift = m.z & 3;
                    for (int j = tid; j < m.w; j += kBlock) {
                        const int k = sm.localBody[j + shift];
                        const d3 d = {sm.d[0][j], sm.d[1][j], sm.d[2][j]};
                        const d4 q = {B[6][k], B[7][k], B[8][k], B[9][k]};
                        const d3 r = {B[0][k], B[1][k], B[2][k]};
                        storeAtom<NATIVE>(pos, atomSlot(S, S.numFree + m.z + j), atomPosition(r, q, d));
                    }
                }
                else {                                         // one body larger than the staging buffer
                    for (int a = m.z + tid; a < m.z + m.w; a += kBlock) {
                        const int k = S.localBody[a];
                        const d3 d = {S.dxyz[a], S.dxyz[a + as], S.dxyz[a + 2*as]};
                        const d4 q = {B[6][k], B[7][k], B[8][k], B[9][k]};
                        const d3 r = {B[0][k], B[1][k], B[2][k]};
                        storeAtom<NATIVE>(pos, atomSlot(S, S.numFree + a), atomPosition(r, q, d));
                    }
                }
            }
            __syncthreads();                                   // buffers are reused by the next tile
        }
        cpWait<0>();
    }
    if (RUNG >= 0 && tile0 < numTiles) {
        // publish the counts; the last CTA of the launch moves the rung for the next launch (same policy as part2Part1Kernel)
        SeriesControl* c = S.seriesCtl;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            ladderFails += __shfl_xor_sync(kFull, ladderFails, off);
            ladderLower += __shfl_xor_sync(kFull, ladderLower, off);
        }
        if ((tid & 31) == 0 && (ladderFails | ladderLower)) {
            if (ladderFails) atomicAdd(&c->fails, ladderFails);
            if (ladderLower) atomicAdd(&c->lower, ladderLower);
            __threadfence();                                   // the counts are visible before this CTA's `done`
        }
        __syncthreads();
        if (tid == 0 && atomicAdd(&c->done, 1u) == gridDim.x - 1) {
            __threadfence();
            const volatile unsigned* counts = &c->fails;       // (both loads in flight together: one round trip at the launch's tail)
            const unsigned fails = counts[0], lower = counts[1];
            const double n = (double) S.numBodies;
            int next = RUNG;
            if ((double) fails > 4.0e-4*n && RUNG < 2) next = RUNG + 1;
            else if (RUNG > 0 && (double) lower < 1.0e-4*n) next = RUNG - 1;
            c->rung = next;
            if (next != c->published) {                        // the host's hint for its choice of kernel (mapped pinned memory)
                *S.hostRungDevice = next;
                c->published = next;
            }
            c->fails = 0u;
            c->lower = 0u;
            c->done = 0u;
        }
    }
}

// Positions of the body atoms from the updated (r, q): one CTA per atom tile, thread per atom.
// Second half of part 1 for large-body systems (see part1Kernel, !FUSED).  Each thread requests the coordinates, body
// bytes and array slots of kLargePerThread atoms (a whole tile in one round) before it touches any of them - the slot
// look-up in atomLoc is not a dependent load in front of the store any more; the kernel is a pure stream and lives on
// the number of loads in flight.
template <bool NATIVE>
__global__ void __launch_bounds__(kBlock) atomPositionKernel(const DeviceSystem S, const AtomView pos) {
    __shared__ double sB[7][kBlock];
    const int tid = threadIdx.x;
    const int4 m = S.tileMeta[blockIdx.x];
    const size_t ld = S.bodyStride, as = S.atomStride;
    const int a1 = m.z + m.w;
    int key[kLargePerThread], slot[kLargePerThread];
    d3 d[kLargePerThread];
#pragma unroll
    for (int u = 0; u < kLargePerThread; u++) {
        const int a = m.z + tid + u*kBlock;
        if (a < a1) {
            key[u] = S.localBody[a];
            slot[u] = S.atomLoc ? S.atomLoc[S.numFree + a] : S.numFree + a;
            d[u] = {S.dxyz[a], S.dxyz[a + as], S.dxyz[a + 2*as]};
        }
    }
    if (tid < m.y) {
        const double* s = S.state + (size_t) (m.x + tid);
        const d3 r = loadPlane3(s + PL_R*ld, ld);
        const d4 q = loadPlane4(s + PL_Q*ld, ld);
        sB[0][tid] = r.x; sB[1][tid] = r.y; sB[2][tid] = r.z;
        sB[3][tid] = q.w; sB[4][tid] = q.x; sB[5][tid] = q.y; sB[6][tid] = q.z;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kLargePerThread; u++) {
        const int a = m.z + tid + u*kBlock;
        if (a < a1) {
            const int k = key[u];
            const d3 r = {sB[0][k], sB[1][k], sB[2][k]};
            const d4 q = {sB[3][k], sB[4][k], sB[5][k], sB[6][k]};
            storeAtom<NATIVE>(pos, slot[u], atomPosition(r, q, d[u]));
        }
    }
    for (int a = m.z + tid + kLargePerThread*kBlock; a < a1; a += kBlock) {      // a single body larger than a tile
        const int k = S.localBody[a];
        const d3 dd = {S.dxyz[a], S.dxyz[a + as], S.dxyz[a + 2*as]};
        const d3 r = {sB[0][k], sB[1][k], sB[2][k]};
        const d4 q = {sB[3][k], sB[4][k], sB[5][k], sB[6][k]};
        storeAtom<NATIVE>(pos, atomSlot(S, S.numFree + a), atomPosition(r, q, dd));
    }
}

// ------------------------------------------------------------------------------------------------
// Part 2: force/torque segmented reduction, second half kick, velocity reconstruction.
// One CTA per atom tile.
// ------------------------------------------------------------------------------------------------
#ifndef RBK_P2_MINBLOCKS
#define RBK_P2_MINBLOCKS 9
#endif
template <bool NATIVE>
__global__ void __launch_bounds__(kBlock, RBK_P2_MINBLOCKS) part2Kernel(const DeviceSystem S, const double dt, const AtomView pos,
                                                     const AtomView vel, const AtomView force) {
    __shared__ double sQ[4][kBlock];
    __shared__ double sAcc[6][kBlock];        // (F, tau) per body, later (v_cm, omega_space)
    __shared__ double sHead[kWarps][6];       // partial sums of a body that started in an earlier warp's range
    __shared__ int sHeadKey[kWarps];
    __shared__ int sLoc[kBlock + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if ((int) blockIdx.x < S.numTiles) {
        const int4 m = S.tileMeta[blockIdx.x];
        const int nb = m.y;
        const size_t ld = S.bodyStride;
        double* s = S.state + (size_t) (m.x + tid);
        if (tid < nb) {                                        // ---- A: thread per body, stage q
            const d4 q = loadPlane4(s + PL_Q*ld, ld);
            sLoc[tid] = S.loc[m.x + tid];
            sQ[0][tid] = q.w; sQ[1][tid] = q.x; sQ[2][tid] = q.y; sQ[3][tid] = q.z;
        }
        if (tid == 0) sLoc[nb] = m.z + m.w;
#pragma unroll
        for (int k = 0; k < 6; k++) sAcc[k][tid] = 0.0;
        if (tid < kWarps*6) sHead[tid/6][tid%6] = 0.0;
        __syncthreads();

        // ---- B: thread per atom.  Each warp walks a contiguous range of the tile's atoms in steps of
        // 32, so that partial sums of one body are always added in the same order (deterministic).
        const int a0 = m.z, a1 = m.z + m.w;
        const size_t as = S.atomStride;
        const int per = ((m.w + kBlock - 1)/kBlock)*32;
        const int wBeg = a0 + warp*per;
        const int wEnd = min(wBeg + per, a1);
        int firstKey = -1;
        if (wBeg < wEnd) {
            const int k0 = S.localBody[wBeg];
            if (sLoc[k0] < wBeg) firstKey = k0;                // this body began in an earlier warp's range
        }
        if (lane == 0) sHeadKey[warp] = firstKey;
        for (int base = wBeg; base < wEnd; base += 32) {
            const int a = base + lane;
            const bool valid = a < wEnd;
            int key = 0x7fffffff;
            double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            if (valid) {
                key = S.localBody[a];
                const d3 d = {S.dxyz[a], S.dxyz[a + as], S.dxyz[a + 2*as]};
                const d3 f = loadAtom<NATIVE>(force, atomSlot(S, S.numFree + a));
                const d4 q = {sQ[0][key], sQ[1][key], sQ[2][key], sQ[3][key]};
                const d3 t = cross(bodyToSpace(q, d), f);
                v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = t.x; v[4] = t.y; v[5] = t.z;
            }
            // inclusive segmented scan over lanes (keys are sorted, so equal keys are contiguous)
            for (int off = 1; off < 32 && off < S.maxBodySize; off <<= 1) {
                const int kp = __shfl_up_sync(kFull, key, off);
                const bool take = lane >= off && kp == key;
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    const double t = __shfl_up_sync(kFull, v[k], off);
                    if (take) v[k] += t;
                }
            }
            const int kn = __shfl_down_sync(kFull, key, 1);
            if (valid && (lane == 31 || kn != key)) {          // last lane of a segment owns its sum
                if (key == firstKey) {
#pragma unroll
                    for (int k = 0; k < 6; k++) sHead[warp][k] += v[k];
                }
                else {
#pragma unroll
                    for (int k = 0; k < 6; k++) sAcc[k][key] += v[k];
                }
            }
            __syncwarp();
        }
        __syncthreads();

        if (tid < nb) {                                        // ---- C: thread per body, second kick
            double sum[6];
#pragma unroll
            for (int k = 0; k < 6; k++) sum[k] = sAcc[k][tid];
#pragma unroll
            for (int w = 1; w < kWarps; w++)
                if (sHeadKey[w] == tid) {
#pragma unroll
                    for (int k = 0; k < 6; k++) sum[k] += sHead[w][k];
                }
            const d3 F = {sum[0], sum[1], sum[2]}, tau = {sum[3], sum[4], sum[5]};
            d3 p = loadPlane3(s + PL_P*ld, ld);
            d4 pi = loadPlane4(s + PL_PI*ld, ld);
            const double invm = s[PL_INVM*ld];
            const d3 invI = loadPlane3(s + PL_INVI*ld, ld);
            const d4 q = {sQ[0][tid], sQ[1][tid], sQ[2][tid], sQ[3][tid]};
            d3 vcm, om;
            bodyPart2(dt, F, tau, invm, invI, q, p, pi, vcm, om);
            storePlane3(s + PL_P*ld, ld, p);
            storePlane4(s + PL_PI*ld, ld, pi);
            storePlane3(s + PL_F*ld, ld, F);
            storePlane3(s + PL_TAU*ld, ld, tau);
            sAcc[0][tid] = vcm.x; sAcc[1][tid] = vcm.y; sAcc[2][tid] = vcm.z;
            sAcc[3][tid] = om.x; sAcc[4][tid] = om.y; sAcc[5][tid] = om.z;
        }
        __syncthreads();

        for (int a = a0 + tid; a < a1; a += kBlock) {          // ---- D: thread per atom, velocities
            const int lb = S.localBody[a];
            const d3 d = {S.dxyz[a], S.dxyz[a + as], S.dxyz[a + 2*as]};
            const d4 q = {sQ[0][lb], sQ[1][lb], sQ[2][lb], sQ[3][lb]};
            const d3 vcm = {sAcc[0][lb], sAcc[1][lb], sAcc[2][lb]};
            const d3 om = {sAcc[3][lb], sAcc[4][lb], sAcc[5][lb]};
            storeAtom<NATIVE>(vel, atomSlot(S, S.numFree + a), atomVelocity(vcm, om, bodyToSpace(q, d)));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Step-fused kernel: Part 2 of step k + Part 1 of step k+1 in one pass (small-body systems).
//
// Nothing happens between the two halves in RigidBodyIntegrator::step, so a tile's bodies can take the
// second half kick, the next first half kick, the drift and the rotation back to back while their state
// sits in shared memory / registers: per body-step the state makes ONE round trip through HBM (read
// r p q pi 1/m 1/I, write r p q pi F tau) instead of two, F and tau are never re-read and the body-frame
// coordinates are read once.  Same persistent cp.async pipeline as part1Kernel, now also prefetching the
// tile's atom forces and coordinates; same deterministic warp-shuffle segmented reduction as part2Kernel,
// reading its operands from shared memory.  Results are bit-identical to part2Kernel + part1Kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kFPlanes = 18;                        // r3 p3 q4 pi4 invm invI3
constexpr int kSmallBody = 8;                       // bodies up to this size are reduced by their own thread

// BODIES = threads per CTA = bodies per tile, ATOMS = atom capacity of a tile.  Two shapes are instantiated:
// <128, 512> (four warps share a tile) and <32, 128> (ONE warp per CTA, for bodies of <= 4 atoms such as water: no
// CTA-wide barrier ever waits for a slower warp, eight independent CTAs per SM).
template <int BODIES, int ATOMS, bool GATHER>
struct alignas(128) FusedStage {                    // 128-byte alignment: destination of TMA tensor copies
    double body[kFPlanes][BODIES];
    double f[3*ATOMS];                              // atom forces as xyzxyz..., later the arms delta = A^T(q) d
    double d[3][ATOMS];
    double w[GATHER ? ATOMS : 2];                   // GATHER: the atoms' inverse masses (velm.w), see fullVelocity below
    int loc[BODIES + 4];
    unsigned char localBody[ATOMS + 32];
};
// GATHER (reordered atoms and / or the OpenMM boundary formats): the atoms' array slots travel two tiles ahead of the data in
// a three-entry ring, so that the per-thread force requests never wait for an atomLoc look-up; the descriptor ring is one
// entry deeper for it.
template <int BODIES, int ATOMS, int STAGES, bool GATHER>
struct FusedSmem {
    FusedStage<BODIES, ATOMS, GATHER> stage[STAGES];
    unsigned long long bar[2];                      // one mbarrier per stage (bulk-copy completion)
    double acc[6][BODIES];
    double head[BODIES/32][6];
    int headKey[BODIES/32];
    int4 meta[GATHER ? 4 : 3];
    int tileIdx[GATHER ? 4 : 3];                    // tile numbers of the current, next and next-but-one (GATHER: + one more) tile (ring)
    int slot[GATHER ? 3 : 1][GATHER ? ATOMS : 4];   // GATHER: array slots of the atoms of the current, next and next-but-one tile (ring)
};

// An atom force as staged in shared memory (raw 8-byte word) -> double: fp64 arrays as they are, OpenMM's fixed-point
// long long planes scaled by 2^-32 (platforms/cuda/src/kernels/rigidbodyintegrator.cu:341,369).
template <bool NATIVE> __device__ __forceinline__ double stagedForce(const AtomView& force, double raw) {
    if (NATIVE || force.fmt == FMT_F64) return raw;
    return (1.0/4294967296.0)*(double) __double_as_longlong(raw);
}

// TMA bulk copies (cp.async.bulk, SASS UBLKCP) with mbarrier completion: one elected thread moves a whole plane
// segment with one instruction instead of every thread issuing an 8-byte cp.async per element.
__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smemAddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulkCopy(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smemAddr(smem)), "l"(gmem), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}
// 2-D TMA tensor copy: the box described by `map` at element coordinates (x, y) -> shared memory (128-byte aligned)
__device__ __forceinline__ void tmaLoad2D(void* smem, const TensorMapBlob* map, int x, int y, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smemAddr(smem)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// stage plane k <-> global state plane: r p q pi are planes 0..13, then invm (20) and invI (24..26)
__device__ __forceinline__ int fusedGlobalPlane(int k) { return k < 14 ? k : (k == 14 ? (int) PL_INVM : (int) PL_INVI + (k - 15)); }

// ------------------------------------------------------------------------------------------------
// Part 2 for large-body systems (mean body size > kSplitAtomsPerBody; atom tiles of <= kLargeBodyTileAtoms atoms).
//
// The bucket for big bodies of the size-bucketed reduction.  Persistent CTAs (one wave) walk the tiles round-robin behind a
// three-deep pipeline: while tile i is processed, tile i+1's coordinates, forces, body bytes and body state arrive in the other
// shared-memory stage, tile i+2's per-body offsets (and first slots / atom slots) in a small ring, and tile i+3's descriptor -
// so no request ever waits for a load it depends on.  Per tile:
//   B   thread per atom: arm delta = A^T(q) d (kept in registers for phase D), torque delta x f written over the coordinates;
//   B2  thread per (body, component): its column of the staged forces / torques is added in atom order (four interleaved
//       partial sums, RigidBody::forceAndTorque, openmmapi/src/RigidBody.cpp:174-183, up to that association) - no shuffles, no
//       atomics, one instruction stream for all six components, deterministic and independent of how the forces came in;
//       (a balanced variant - threads per chunk of consecutive atoms, partial sums per body - executed 30 % more instructions: 181 us vs 136)
//   C   thread per body: second kick;   D   thread per atom: velocities.
// A body larger than a tile is alone in its tile and is reduced by the whole CTA (strided partial sums, fixed tree).
//
// The tile's coordinate planes and body bytes are the handle's own data and always arrive as four TMA bulk copies on the
// stage's mbarrier.  RUNS (S.bodyRun: the atoms of every body sit in consecutive slots of the caller's arrays, in order - what
// OpenMM keeps through its atom re-orderings, which move whole molecules, and what any topology-ordered atom list looks
// like): nothing is requested per atom - the forces of each body are ONE bulk copy (Vec3 rows) or three (planes: SoA doubles,
// OpenMM's fixed-point long long) issued by the body's thread, and an atom's array slot is its tile index plus a per-body
// constant.  Bulk copies move 16-byte granules between 16-byte aligned addresses: a run that starts at an odd 8-byte word is
// copied from one word earlier and lands at an even word of the stage (per-body offset table fo); the array's last word is
// fetched by an 8-byte cp.async when the rounded copy would pass the end of the caller's array.  !RUNS (atoms permuted one by
// one, misaligned planes): per-thread 8-byte cp.async through a ring of array slots that travels two tiles ahead.
// ncu on config 4, round 1 formulation (per-atom requests everywhere, 16-lane butterfly per body): 85 M warp-instructions, 33 % of
// them request code, 32 % the butterfly, 150 us; RUNS: 66 M, 128 us.
// ------------------------------------------------------------------------------------------------
#ifndef RBK_P2L_THREADS
#define RBK_P2L_THREADS 128
#endif
constexpr int kP2LThreads = RBK_P2L_THREADS;        // (256 threads per tile measured slower: 0.346 vs 0.313 ms/step on config 4)
constexpr int kP2LStatePlanes = 15;                 // q4 p3 pi4 invm invI3
constexpr int kP2LFree = 64;                        // most free atoms a tile takes along (see part2LargeKernel, freePhase)
template <int NB, bool RUNS> struct Part2LargeLayout {     // byte offsets inside one stage / the CTA's shared memory; pitches in doubles
    static constexpr int A = kLargeBodyTileAtoms;
    static constexpr int dp = A + 2;                // coordinate plane: the copy starts at an even atom
    static constexpr int fp = RUNS ? A + 2*NB + 2 : A;      // force plane: every body's run starts at an even word, two spare words each
    static constexpr int f = 3*dp*8;                // d[3][dp] sits at offset 0
    static constexpr int st = f + 3*fp*8;
    static constexpr int key = st + kP2LStatePlanes*NB*8;
    static constexpr int fo = key + A + 32;         // int fo[NB]: staged word of atom j, component c = fo[body] + j*fs + c*fc
    static constexpr int rd = fo + NB*4;            // int rd[NB]: array slot of atom j = j + rd[body]
    static constexpr int stageBytes = (rd + NB*4 + 127) & ~127;
    static constexpr int acc = 2*stageBytes;        // double acc[6][NB]
    static constexpr int ringFree = NB + 4 + (RUNS ? NB + 4 : A);      // 3 x { int loc[NB + 4]; int run[NB + 4] | int slot[A]; int freeSlot[kP2LFree]; }
    static constexpr int ringInts = ringFree + kP2LFree;
    static constexpr int ring = acc + 6*NB*8;
    static constexpr int meta = (ring + 3*ringInts*4 + 15) & ~15;      // int4[4]
    static constexpr int bar = meta + 4*16;         // two mbarriers
    static constexpr int total = bar + 16;
};
// the request code works on 32-bit shared-memory addresses (one conversion per kernel instead of one per instruction)
__device__ __forceinline__ void cpAsync8s(unsigned smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cpAsync4s(unsigned smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void bulkCopyS(unsigned smem, const void* gmem, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem), "l"(gmem), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbarArriveS(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void cpAsync16s(unsigned smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem), "l"(gmem) : "memory");
}
__device__ __forceinline__ void mbarWaitS(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbarExpectTxS(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}

// Register cap: five / four / three resident CTAs of 128 threads (tiles of 256 / 384 / 512 atoms).
#ifndef RBK_P2L_MAXNREG
#define RBK_P2L_MAXNREG (RBK_LARGE_PER_THREAD <= 2 ? 96 : RBK_LARGE_PER_THREAD == 3 ? 128 : 168)
#endif
// NB: body capacity of a tile's stage (S.stageBodies rounded up to one of the instantiated sizes): compile-time, so that the
// shared-memory layout is a set of immediates.
// freePhase != 0: the FREE atoms ride along (2 = their Part 2, 3 = Part 2 + Part 1 of the next step, as freeAtomsKernel): tile t
// takes the t-th slice of the free-atom list, at most kP2LFree atoms - their slots come through the ring two tiles ahead, the
// last threads of the CTA load them when the sums start (phase B2, where those warps have little or nothing to do), and
// finish them next to the thread-per-body kick (phase C, one busy warp).  No separate launch, no second stream, and the
// sectors a free atom shares with the body atoms around it are in flight at the same time.
template <bool NATIVE, int NB, bool RUNS>
__global__ void __maxnreg__(RBK_P2L_MAXNREG) part2LargeKernel(const DeviceSystem S, const double dt, const AtomView pos, const AtomView vel,
                                                              const AtomView force, const int freePhase) {
    extern __shared__ __align__(128) unsigned char smemRaw[];
    typedef Part2LargeLayout<NB, RUNS> L;
    constexpr int A = kLargeBodyTileAtoms;
    constexpr int kBlock = kP2LThreads, kWarps = kP2LThreads/32;          // shadow the file-level constants inside this kernel
    static_assert(kLargeBodyTileAtoms == kLargePerThread*kP2LThreads, "a thread keeps the arms of its kLargePerThread atoms in registers");
    static_assert(NB % 4 == 0 && NB <= kP2LThreads - 8, "body threads and the four plane-copy threads are different threads");
    double* const sAcc = reinterpret_cast<double*>(smemRaw + L::acc);     // [6][NB] (F, tau), later (v_cm, omega_space)
    int4* const sMeta = reinterpret_cast<int4*>(smemRaw + L::meta);       // ring of 4 tile descriptors
    int* const sRing = reinterpret_cast<int*>(smemRaw + L::ring);
    unsigned long long* const bar = reinterpret_cast<unsigned long long*>(smemRaw + L::bar);
    const unsigned sBase = smemAddr(smemRaw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, numTiles = S.numTiles;
    const size_t ld = S.bodyStride, as = S.atomStride;
    const int* const atomLoc = S.atomLoc;
    const bool rows = !RUNS || force.sa == 3;                             // staged as Vec3 rows; otherwise three planes (RUNS, force.sc apart)
    const int fs = rows ? 3 : 1, fc = rows ? 1 : L::fp;
    const bool fixedPoint = RUNS && !NATIVE && force.fmt == FMT_FORCE_FIXED;
    // free atoms of tile t: [freeFirst(t), freeFirst(t + 1)) - an even split of the list in 2^-20 fixed point
    const unsigned long long freeRatio = freePhase ? ((unsigned long long) S.numFree << 20)/(unsigned) numTiles : 0ull;
    auto freeFirst = [&](int t) { return t >= numTiles ? S.numFree : (int) (((unsigned long long) t*freeRatio) >> 20); };

    // per-body atom offsets (one more than bodies: the end of the last one) and first slots | atom slots of a tile -> ring entry r
    auto requestBodies = [&](int4 m, int r, int2 freeRange) {
        const unsigned dst = sBase + L::ring + (r*L::ringInts + tid)*4;
        if (tid <= m.y) cpAsync4s(dst, S.loc + m.x + tid);
        if (freePhase && atomLoc != nullptr && tid < freeRange.y - freeRange.x) cpAsync4s(dst + L::ringFree*4, atomLoc + freeRange.x + tid);
        if (RUNS) {
            if (tid < m.y) cpAsync4s(dst + (NB + 4)*4, S.bodyRun + m.x + tid);
        }
        else if (atomLoc != nullptr && m.w <= A) {
            const int* g = atomLoc + S.numFree + m.z;
            for (int j = tid; j < m.w; j += kBlock) cpAsync4s(dst + (NB + 4 + j - tid)*4, g + j);
        }
    };
    // coordinates, forces, body bytes and body state of the tile whose offsets are in ring entry r; every thread arrives on the
    // stage's barrier, the ones that start bulk copies with their byte counts
    auto requestData = [&](int4 m, int r, int stage) {
        const unsigned T = sBase + stage*L::stageBytes, sbar = sBase + L::bar + 8*stage;
        const int nb = m.y, a0 = m.z, na = m.w;
        if (tid < nb) {
            const double* g = S.state + (size_t) (m.x + tid);
            const unsigned dst = T + L::st + 8*tid;
#pragma unroll
            for (int k = 0; k < 4; k++) cpAsync8s(dst + k*NB*8, g + (PL_Q + k)*ld);
#pragma unroll
            for (int k = 0; k < 3; k++) cpAsync8s(dst + (4 + k)*NB*8, g + (PL_P + k)*ld);
#pragma unroll
            for (int k = 0; k < 4; k++) cpAsync8s(dst + (7 + k)*NB*8, g + (PL_PI + k)*ld);
            cpAsync8s(dst + 11*NB*8, g + PL_INVM*ld);
#pragma unroll
            for (int k = 0; k < 3; k++) cpAsync8s(dst + (12 + k)*NB*8, g + (PL_INVI + k)*ld);
        }
        if (na > A) { mbarArriveS(sbar); return; }               // one body larger than a tile: its atoms are read in place
        const int* const ring = sRing + r*L::ringInts;
        if (!RUNS)
            for (int j = tid; j < na; j += kBlock) {             // per-atom force requests through the slot ring
                const long long slot = atomLoc != nullptr ? ring[NB + 4 + j] : S.numFree + a0 + j;
                if (NATIVE) {
                    const double* fp = force.p + slot*force.sa;
                    const unsigned dst = T + L::f + 24*j;
                    cpAsync8s(dst, fp);
                    cpAsync8s(dst + 8, fp + force.sc);
                    cpAsync8s(dst + 16, fp + 2*force.sc);
                }
                else {
                    const d3 f = loadAtom<false>(force, slot);
                    double* sF = reinterpret_cast<double*>(smemRaw + stage*L::stageBytes + L::f);
                    sF[3*j] = f.x; sF[3*j + 1] = f.y; sF[3*j + 2] = f.z;
                }
            }
        if (RUNS && tid < nb) {
            const int j0 = ring[tid] - a0, n = ring[tid + 1] - ring[tid];
            const long long slot0 = ring[NB + 4 + tid];
            int* const tables = reinterpret_cast<int*>(smemRaw + stage*L::stageBytes + L::fo);
            tables[NB + tid] = (int) slot0 - j0;
            if (rows) {
                const long long w0 = 3*slot0;
                const int ph = (int) (w0 & 1), s = (3*j0 + 2*tid + 1) & ~1;
                int words = (ph + 3*n + 1) & ~1, tail = 0;
                if (w0 - ph + words > 3LL*S.numSlots) { words -= 2; tail = 1; }      // the rounded copy would pass the end of the array
                const double* src = force.p + (w0 - ph);
                tables[tid] = s + ph - 3*j0;
                mbarExpectTxS(sbar, 8u*words);
                bulkCopyS(T + L::f + 8*s, src, 8u*words, sbar);
                if (tail) cpAsync8s(T + L::f + 8*(s + words), src + words);
            }
            else {
                const int ph = (int) (slot0 & 1), s = (j0 + 2*tid + 1) & ~1, words = (ph + n + 1) & ~1;
                const double* src = force.p + (slot0 - ph);
                tables[tid] = s + ph - j0;
                mbarExpectTxS(sbar, 24u*words);
#pragma unroll
                for (int c = 0; c < 3; c++) bulkCopyS(T + L::f + 8*(c*L::fp + s), src + c*force.sc, 8u*words, sbar);
            }
        }
        else if (tid >= kBlock - 3) {
            const int c = kBlock - 1 - tid, first = a0 & ~1;
            const unsigned bytes = 8u*(((a0 + na + 1) & ~1) - first);
            mbarExpectTxS(sbar, bytes);
            bulkCopyS(T + 8*c*L::dp, S.dxyz + c*as + first, bytes, sbar);
        }
        else if (tid == kBlock - 4) {
            const int first = a0 & ~15;
            const unsigned bytes = ((a0 + na + 15) & ~15) - first;
            mbarExpectTxS(sbar, bytes);
            bulkCopyS(T + L::key, S.localBody + first, bytes, sbar);
        }
        else mbarArriveS(sbar);
    };

    const int tile0 = blockIdx.x;
    if (tile0 >= numTiles) return;
    if (tid == 0) {
        mbarInit(&bar[0], kBlock);
        mbarInit(&bar[1], kBlock);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 3 && tile0 + tid*G < numTiles) sMeta[tid] = S.tileMeta[tile0 + tid*G];
    __syncthreads();
    // free-atom ranges of this CTA's current, next and next-but-one tile (rolled along: two evaluations per tile)
    auto freeRangeOf = [&](int t) { return make_int2(freeFirst(t), freeFirst(t + 1)); };
    int2 freeCur = freeRangeOf(tile0), freeNext = freeRangeOf(tile0 + G), freeAfter = freeRangeOf(tile0 + 2*G);
    requestBodies(sMeta[0], 0, freeCur);
    if (tile0 + G < numTiles) requestBodies(sMeta[1], 1, freeNext);
    cpCommit();
    cpWait<0>();
    __syncthreads();
    requestData(sMeta[0], 0, 0);
    cpCommit();
    int it = 0, r0 = 0;                                          // r0 = it % 3: ring entry of the current tile
    for (int tile = tile0; tile < numTiles; tile += G, it++) {
        const int cur = it & 1;
        const int r1 = r0 == 2 ? 0 : r0 + 1, r2 = r1 == 2 ? 0 : r1 + 1;
        cpWait<0>();                                             // state (and forces) of tile `it`, the bodies of it+1, the descriptor of it+2 have landed
        mbarWaitS(sBase + L::bar + 8*cur, (it >> 1) & 1);        // ... and its coordinates, body bytes (and forces)
        __syncthreads();
        const int4 m = sMeta[it & 3];
        if (tile + G < numTiles) requestData(sMeta[(it + 1) & 3], r1, cur ^ 1);
        if (tile + 2*G < numTiles) requestBodies(sMeta[(it + 2) & 3], r2, freeAfter);
        if (tid == 0 && tile + 3*G < numTiles) cpAsync16s(sBase + L::meta + 16*((it + 3) & 3), S.tileMeta + tile + 3*G);
        cpCommit();
        if (freePhase && tile + G < numTiles) {                  // the free atoms of the NEXT tile: towards L2 now, loaded a tile later
            const int k = freeNext.x + (kBlock - 1 - tid);
            if (k < freeNext.y) {
                const long long slot = atomLoc != nullptr ? sRing[r1*L::ringInts + L::ringFree + kBlock - 1 - tid] : k;
                prefetchAtom(force, slot);
                prefetchAtom(pos, slot);
                prefetchAtom(vel, slot);
                prefetchL2(S.freeInvMass + k);
                if (freePhase & 2) { prefetchL2(S.savedPos + k); prefetchL2(S.savedPos + k + S.freeStride); prefetchL2(S.savedPos + k + 2*S.freeStride); }
            }
        }

        unsigned char* T = smemRaw + cur*L::stageBytes;
        double* const sD = reinterpret_cast<double*>(T) + (m.z & 1);          // [3][dp] body-frame coordinates, then the torques delta x f
        double* const sF = reinterpret_cast<double*>(T + L::f);               // forces: rows xyzxyz... | as copied (rows or planes, per-body offsets)
        double* const sSt = reinterpret_cast<double*>(T + L::st);             // [15][NB]
        const int* const sFo = reinterpret_cast<const int*>(T + L::fo);
        const int* const sRd = sFo + NB;
        const int* const sLoc = sRing + r0*L::ringInts;
        const int* const sSlot = sLoc + NB + 4;                               // !RUNS: array slots of the tile's atoms
        const unsigned char* const sKey = T + L::key + (m.z & 15);
        const int nb = m.y, a0 = m.z, na = m.w;
        r0 = r1;

        // the free atom this thread takes along (the CTA's last threads, one each)
        const int freeK = freeCur.x + (kBlock - 1 - tid);
        const bool freeMine = freePhase && freeK < freeCur.y;
        freeCur = freeNext;
        freeNext = freeAfter;
        if (freePhase) freeAfter = freeRangeOf(tile + 3*G);
        long long freeSlot = 0;
        d3 freeF, freeX, freeV, freeSaved;
        double freeInvm = 0.0;
        auto freeLoad = [&]() {
            freeSlot = atomLoc != nullptr ? sLoc[L::ringFree + kBlock - 1 - tid] : freeK;
            freeF = loadAtom<NATIVE>(force, freeSlot);
            freeInvm = S.freeInvMass[freeK];
            freeX = loadAtom<NATIVE>(pos, freeSlot);
            freeV = loadAtom<NATIVE>(vel, freeSlot);
            if (freePhase & 2) freeSaved = loadPlane3(S.savedPos + freeK, S.freeStride);
        };
        auto freeFinish = [&]() {
            if (freePhase & 2) freePart2(dt, freeF, freeInvm, freeX, freeSaved, freeV);
            if (freePhase & 1) {
                freePart1(dt, freeF, freeInvm, freeX, freeV);
                storeAtom<NATIVE>(pos, freeSlot, freeX);
                storePlane3(S.savedPos + freeK, S.freeStride, asStored<NATIVE>(pos, freeX));
            }
            if (!NATIVE && vel.fmt == FMT_REAL4_F64 && S.atomInvMass != nullptr) {      // whole double4, w = 1/m (see part2Part1Kernel)
                double2* out = reinterpret_cast<double2*>(vel.p + 4*freeSlot);
                out[0] = make_double2(freeV.x, freeV.y);
                out[1] = make_double2(freeV.z, freeInvm);
            }
            else storeAtom<NATIVE>(vel, freeSlot, freeV);
        };

        if (na > A) {                                            // ---- one body larger than a tile: CTA-wide reduction, atoms read in place
            double* s = S.state + (size_t) m.x;
            const long long slot0 = RUNS ? sLoc[NB + 4] : 0;
            auto slotOf = [&](int j) { return RUNS ? slot0 + j : atomSlot(S, S.numFree + a0 + j); };
            const d4 q = {sSt[0], sSt[NB], sSt[2*NB], sSt[3*NB]};
            double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int j = tid; j < na; j += kBlock) {
                const int a = a0 + j;
                const d3 d = {S.dxyz[a], S.dxyz[a + as], S.dxyz[a + 2*as]};
                const d3 f = loadAtom<NATIVE>(force, slotOf(j));
                const d3 t = cross(bodyToSpace(q, d), f);
                v[0] += f.x; v[1] += f.y; v[2] += f.z; v[3] += t.x; v[4] += t.y; v[5] += t.z;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                for (int k = 0; k < 6; k++) v[k] += __shfl_xor_sync(kFull, v[k], off);
            if (lane == 0)
#pragma unroll
                for (int k = 0; k < 6; k++) sF[warp*6 + k] = v[k];
            __syncthreads();
            if (tid == 0) {
#pragma unroll
                for (int w = 1; w < kWarps; w++)
#pragma unroll
                    for (int k = 0; k < 6; k++) v[k] += sF[w*6 + k];
                const d3 F = {v[0], v[1], v[2]}, tau = {v[3], v[4], v[5]};
                d3 p = {sSt[4*NB], sSt[5*NB], sSt[6*NB]};
                d4 pi = {sSt[7*NB], sSt[8*NB], sSt[9*NB], sSt[10*NB]};
                d3 vcm, om;
                bodyPart2(dt, F, tau, sSt[11*NB], d3{sSt[12*NB], sSt[13*NB], sSt[14*NB]}, q, p, pi, vcm, om);
                storePlane3(s + PL_P*ld, ld, p);
                storePlane4(s + PL_PI*ld, ld, pi);
                storePlane3(s + PL_F*ld, ld, F);
                storePlane3(s + PL_TAU*ld, ld, tau);
                sAcc[0] = vcm.x; sAcc[1] = vcm.y; sAcc[2] = vcm.z; sAcc[3] = om.x; sAcc[4] = om.y; sAcc[5] = om.z;
            }
            __syncthreads();
            const d3 vcm = {sAcc[0], sAcc[1], sAcc[2]}, om = {sAcc[3], sAcc[4], sAcc[5]};
            for (int j = tid; j < na; j += kBlock) {
                const int a = a0 + j;
                const d3 d = {S.dxyz[a], S.dxyz[a + as], S.dxyz[a + 2*as]};
                storeAtom<NATIVE>(vel, slotOf(j), atomVelocity(vcm, om, bodyToSpace(q, d)));
            }
            if (freeMine) { freeLoad(); freeFinish(); }
            fenceProxyAsync();
            __syncthreads();
            continue;
        }

        // ---- B: thread per atom: arm delta = A^T(q) d (kept in registers), torque delta x f over the coordinates
        d3 delta[kLargePerThread];
        int key[kLargePerThread];
#pragma unroll
        for (int u = 0; u < kLargePerThread; u++) {
            const int j = tid + u*kBlock;
            if (j < na) {
                const int k = sKey[j];
                key[u] = k;
                const d4 q = {sSt[k], sSt[NB + k], sSt[2*NB + k], sSt[3*NB + k]};
                delta[u] = bodyToSpace(q, d3{sD[j], sD[L::dp + j], sD[2*L::dp + j]});
                double* fw = sF + (RUNS ? sFo[k] + j*fs : 3*j);
                d3 f = {fw[0], fw[fc], fw[2*fc]};
                if (fixedPoint) {                                // fixed point -> double once, in place (the sums read doubles)
                    f = {stagedForce<NATIVE>(force, f.x), stagedForce<NATIVE>(force, f.y), stagedForce<NATIVE>(force, f.z)};
                    fw[0] = f.x; fw[fc] = f.y; fw[2*fc] = f.z;
                }
                const d3 t = cross(delta[u], f);
                sD[j] = t.x; sD[L::dp + j] = t.y; sD[2*L::dp + j] = t.z;
            }
        }
        __syncthreads();

        // ---- B2: thread per (body, component): the column of forces / torques of the body in atom order, four interleaved
        // partial sums (atoms 0 4 8 .., 1 5 9 .., ..) so that the chain of dependent additions is a quarter of the body
        if (freeMine) freeLoad();                                // (in flight while the sums run)
        for (int idx = tid; idx < 6*nb; idx += kBlock) {
            const int b = idx/6, c = idx - 6*b;
            const int j0 = sLoc[b] - a0, n = sLoc[b + 1] - sLoc[b];
            const double* ptr = c < 3 ? sF + (RUNS ? sFo[b] : 0) + j0*fs + c*fc : sD + (c - 3)*L::dp + j0;
            const int stride = c < 3 ? fs : 1;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int i = 0;
            for (; i + 4 <= n; i += 4, ptr += 4*stride) {
                s0 += ptr[0]; s1 += ptr[stride]; s2 += ptr[2*stride]; s3 += ptr[3*stride];
            }
            if (i < n) s0 += ptr[0];
            if (i + 1 < n) s1 += ptr[stride];
            if (i + 2 < n) s2 += ptr[2*stride];
            sAcc[c*NB + b] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();

        if (tid < nb) {                                          // ---- C: thread per body, second kick
            const d3 F = {sAcc[tid], sAcc[NB + tid], sAcc[2*NB + tid]};
            const d3 tau = {sAcc[3*NB + tid], sAcc[4*NB + tid], sAcc[5*NB + tid]};
            const d4 q = {sSt[tid], sSt[NB + tid], sSt[2*NB + tid], sSt[3*NB + tid]};
            d3 p = {sSt[4*NB + tid], sSt[5*NB + tid], sSt[6*NB + tid]};
            d4 pi = {sSt[7*NB + tid], sSt[8*NB + tid], sSt[9*NB + tid], sSt[10*NB + tid]};
            const double invm = sSt[11*NB + tid];
            const d3 invI = {sSt[12*NB + tid], sSt[13*NB + tid], sSt[14*NB + tid]};
            d3 vcm, om;
            bodyPart2(dt, F, tau, invm, invI, q, p, pi, vcm, om);
            double* s = S.state + (size_t) (m.x + tid);
            storePlane3(s + PL_P*ld, ld, p);
            storePlane4(s + PL_PI*ld, ld, pi);
            storePlane3(s + PL_F*ld, ld, F);
            storePlane3(s + PL_TAU*ld, ld, tau);
            sAcc[tid] = vcm.x; sAcc[NB + tid] = vcm.y; sAcc[2*NB + tid] = vcm.z;
            sAcc[3*NB + tid] = om.x; sAcc[4*NB + tid] = om.y; sAcc[5*NB + tid] = om.z;
        }
        if (freeMine) freeFinish();
        __syncthreads();

#pragma unroll
        for (int u = 0; u < kLargePerThread; u++) {              // ---- D: thread per atom, velocities
            const int j = tid + u*kBlock;
            if (j < na) {
                const int k = key[u];
                const d3 vcm = {sAcc[k], sAcc[NB + k], sAcc[2*NB + k]};
                const d3 om = {sAcc[3*NB + k], sAcc[4*NB + k], sAcc[5*NB + k]};
                const long long slot = RUNS ? j + sRd[k] : (atomLoc != nullptr ? sSlot[j] : S.numFree + a0 + j);
                storeAtom<NATIVE>(vel, slot, atomVelocity(vcm, om, delta[u]));
            }
        }
        fenceProxyAsync();                                       // the next bulk copies into this stage come after these accesses
        __syncthreads();                                         // the stage, the ring entry and sAcc are reused
    }
    cpWait<0>();
}

template <bool NATIVE, int NB, bool RUNS>
cudaError_t launchPart2LargeShape(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, int freePhase, cudaStream_t st) {
    const size_t smem = (size_t) Part2LargeLayout<NB, RUNS>::total;
    static LaunchCache cache;
    int blocks = 0;                                            // persistent CTAs: one full wave, whatever fits
    cudaError_t e = cache.get(part2LargeKernel<NATIVE, NB, RUNS>, kP2LThreads, smem, blocks);
    if (e != cudaSuccess) return e;
    const int resident = S.numSMs*blocks;
    part2LargeKernel<NATIVE, NB, RUNS><<<S.numTiles < resident ? S.numTiles : resident, kP2LThreads, smem, st>>>(S, dt, pos, vel, force, freePhase);
    return cudaGetLastError();
}

template <bool NATIVE, bool RUNS>
cudaError_t launchPart2LargeRuns(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, int freePhase, cudaStream_t st) {
    if (S.stageBodies <= 16) return launchPart2LargeShape<NATIVE, 16, RUNS>(S, dt, pos, vel, force, freePhase, st);
    if (S.stageBodies <= 32) return launchPart2LargeShape<NATIVE, 32, RUNS>(S, dt, pos, vel, force, freePhase, st);
    if (S.stageBodies <= 64) return launchPart2LargeShape<NATIVE, 64, RUNS>(S, dt, pos, vel, force, freePhase, st);
    return launchPart2LargeShape<NATIVE, kP2LThreads - 8, RUNS>(S, dt, pos, vel, force, freePhase, st);
}

// Can the free atoms ride along in part2LargeKernel?  (An even split of the list over the tiles, at most kP2LFree - 1 each.)
inline bool freeAtomsRide(const DeviceSystem& S) {
    return S.splitPart1 && S.numTiles > 0 && S.numFree > 0 && !S.noFreeRide && (long long) S.numFree <= (long long) (kP2LFree - 2)*S.numTiles;
}

// freePhase: 0 = bodies only, 2 / 3 = the free atoms ride along (the caller has checked freeAtomsRide)
template <bool NATIVE>
cudaError_t launchPart2Large(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, int freePhase, cudaStream_t st) {
    // bodies that are runs of the caller's arrays, forces as Vec3 rows or as three 16-byte aligned planes: bulk copies per body
    const bool rows = force.sa == 3 && force.sc == 1 && force.fmt == FMT_F64;
    const bool planes = force.sa == 1 && (force.sc & 1) == 0 && force.sc >= S.numSlots && (force.fmt == FMT_F64 || force.fmt == FMT_FORCE_FIXED);
    if (S.bodyRun != nullptr && !S.noBulkPart2 && (rows || planes) && (reinterpret_cast<size_t>(force.p) & 15) == 0)
        return launchPart2LargeRuns<NATIVE, true>(S, dt, pos, vel, force, freePhase, st);
    return launchPart2LargeRuns<NATIVE, false>(S, dt, pos, vel, force, freePhase, st);
}

// STAGES = 2: the next tile is staged while the current one is processed (exact rotation: registers allow two CTAs of
// 128 threads per SM anyway).  STAGES = 1: the next tile is requested when the current one is finished and the
// smaller footprint lets four CTAs share an SM, which hide each other's load latency (NO-SQUISH: 112 registers).
// P1ONLY: Part 1 alone through the same pipeline (the step's opening launch and callers that evaluate forces between
// two launches): no forces are staged, the stored F and tau planes take their place in the stage, Part 2's phases drop out.
// GATHER: the instantiation for everything that is not "fp64 arrays in plugin order" - reordered atoms (atomLoc) and the
// OpenMM-CUDA boundary formats (float4 posq + correction, mixed4 velm, fixed-point force planes): the state planes, body-frame
// coordinates, offsets and body bytes still arrive by TMA (they are the handle's own, in body order), the forces by
// per-thread 8-byte cp.async through the slot ring (raw words, converted where they are consumed); positions and velocities
// leave through the format-aware stores.  storeFT = false: interior step of step(n) - nothing reads F and tau before the
// next Part 2 rewrites them, so they stay in registers (48 B per body-step less).
// LADDER (one-warp exact-rotation tiles): 1 = the series order comes from the device-resident ladder control
// (SeriesControl); 2 = the lowest rung only (order 11, 232 registers instead of 254 and one inlined series instead of three:
// 4 % faster), launched when the host's copy of the rung says 0 - a hint that may be a few launches old, which is safe
// because every rung is a complete algorithm (bodies that fail its check take the retry path) and this kernel keeps the
// control block up to date like the full one; 0 = fixed order, no control block.
#ifndef RBK_UNROLL_TRIATOMIC
#define RBK_UNROLL_TRIATOMIC 1          // tiles of three-atom bodies: force sums and atom outputs written out (ILP 3)
#endif
#ifndef RBK_WARP_TILE_CTAS
#define RBK_WARP_TILE_CTAS 8            // resident one-warp CTAs per SM the register budget is cut for (254 registers at 8)
#endif
template <bool EXACT, bool SMALL, bool NATIVE, int BODIES, int ATOMS, int STAGES, bool P1ONLY, bool GATHER, int LADDER>
__global__ void __launch_bounds__(BODIES, BODIES == 32 ? RBK_WARP_TILE_CTAS : (STAGES == 2 ? 256 : 512)/BODIES)
part2Part1Kernel(const DeviceSystem S, const double dt, const AtomView pos, const AtomView vel, const AtomView force,
                 const __grid_constant__ TileMaps maps, const bool useMaps, const bool storeFT) {
    extern __shared__ __align__(128) unsigned char smemRaw[];
    typedef FusedStage<BODIES, ATOMS, GATHER> Stage;
    FusedSmem<BODIES, ATOMS, STAGES, GATHER>& sm = *reinterpret_cast<FusedSmem<BODIES, ATOMS, STAGES, GATHER>*>(smemRaw);
    constexpr int kBlock = BODIES, kWarps = BODIES/32;          // shadow the file-level constants inside this kernel
    constexpr int RING = GATHER ? 4 : 3;
    static_assert(!LADDER || (EXACT && BODIES == 32), "the series ladder is wired for one-warp exact-rotation tiles");
    const int rung = LADDER == 1 ? S.seriesCtl->rung : 0;
    // three-atom tiles written out (ILP 3) where it measured faster: the lean fp64 kernel (0.1094 -> 0.1073 ms at 1 M waters);
    // in the full-ladder kernel (255 registers) and the OpenMM-format kernel it measured slower (4 fs 0.199 -> 0.215, mixed 0.143 -> 0.149)
    constexpr bool kTriatomic = RBK_UNROLL_TRIATOMIC && BODIES == 32 && NATIVE && LADDER == 2;
    unsigned ladderFails = 0u, ladderLower = 0u;               // bodies of this CTA's tiles (warp-uniform)
    const int4* const tileMeta = BODIES == 32 ? S.warpTileMeta : S.tileMeta;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x;
    const size_t ld = S.bodyStride, as = S.atomStride;
    const int numTiles = BODIES == 32 ? S.numWarpTiles : S.numTiles;

    // The handle's own data of a tile (state planes, body-frame coordinates, offsets, body bytes) arrives by TMA: 2-D tensor
    // boxes for one-warp tiles, else 1-D bulk copies when every segment is 16-byte aligned and a multiple of 16 bytes long,
    // else per-thread cp.async.  The caller's forces ride along as ONE bulk copy when they are a contiguous, aligned Vec3 range
    // (water tiles in plugin order always are); otherwise (SoA planes, odd offsets, GATHER) by per-thread 8-byte cp.async.
    const bool contiguousForces = !GATHER && S.atomLoc == nullptr && force.sa == 3 && force.sc == 1 && (reinterpret_cast<size_t>(force.p) & 15) == 0;
    auto forcesBulk = [&](int4 m) { return contiguousForces && ((S.numFree + m.z) & 1) == 0 && (m.w & 1) == 0; };
    auto bulkOK = [&](int4 m) { return (m.x & 3) == 0 && (m.y & 3) == 0 && (m.z & 1) == 0 && (m.w & 1) == 0; };
    auto tensorOK = [&](int4 m) { return BODIES == 32 && useMaps; };
    // GATHER: array slots of a tile's atoms into ring entry `ring` (two tiles ahead of the data)
    auto requestSlots = [&](int4 m, int ring) {
        int* dst = sm.slot[GATHER ? ring : 0];
        if (S.atomLoc != nullptr)
            for (int j = tid; j < m.w; j += kBlock) cpAsync4(dst + j, S.atomLoc + S.numFree + m.z + j);
        else
            for (int j = tid; j < m.w; j += kBlock) dst[j] = S.numFree + m.z + j;
    };
    // OpenMM's velm is mixed4 = (vx, vy, vz, 1/m).  Writing 24 of its 32 bytes makes every sector a partial write that the
    // memory system completes with a fill read (32 B per atom); the handle knows the masses, so the one-pass kernel writes
    // the whole double4 with w = 1/m - the same bits OpenMM put there (1.0/mass in double, constant for a Context's lifetime).
    // posq.w (the charge) can change under updateParametersInContext and is never written.
    const bool fullVelocity = GATHER && !P1ONLY && vel.fmt == FMT_REAL4_F64 && S.atomInvMass != nullptr;
    // GATHER: the tile's forces, one raw 8-byte word per component (double or fixed-point long long), through the slots
    auto requestForces = [&](int4 m, Stage& T, const int* slots) {
        for (int j = tid; j < m.w; j += kBlock) {
            const double* fp = force.p + (GATHER ? (long long) slots[j] : atomSlot(S, S.numFree + m.z + j))*force.sa;
            cpAsync8(&T.f[3*j], fp);
            cpAsync8(&T.f[3*j + 1], fp + force.sc);
            cpAsync8(&T.f[3*j + 2], fp + 2*force.sc);
            if (fullVelocity) cpAsync8(&T.w[j], S.atomInvMass + m.z + j);
        }
    };
    auto request = [&](int4 m, int st, const int* slots) {
        Stage& T = sm.stage[st];
        const int lbFirst = m.z & ~15;                         // 16-byte granules of the byte array
        const unsigned lbBytes = (unsigned) (((m.z + m.w - lbFirst) + 15) & ~15);
        if (tensorOK(m)) {
            // one-warp tiles: the tile's state planes are ONE 2-D box (32 bodies x 18 or 24 planes), its body-frame
            // coordinates another (ATOMS atoms x 3 planes; atoms past the tile are fetched and ignored, past the array
            // zero-filled) - five copies per tile instead of 24, and no alignment rule on the tile's offsets
            if (tid == 0) {
                fenceProxyAsync();
                constexpr unsigned boxBytes = (P1ONLY ? 24u : 18u)*BODIES*8u + 3u*ATOMS*8u + 4u*BODIES;
                const bool fBulk = !P1ONLY && forcesBulk(m);
                mbarExpectTx(&sm.bar[st], boxBytes + (fBulk ? 24u*m.w : 0u) + lbBytes);
                tmaLoad2D(&T.body[0][0], P1ONLY ? &maps.state24 : &maps.state18, m.x, 0, &sm.bar[st]);
                tmaLoad2D(&T.d[0][0], &maps.dxyz, m.z, 0, &sm.bar[st]);
                bulkCopy(&T.loc[0], S.loc + m.x, 4u*BODIES, &sm.bar[st]);
                if (fBulk) bulkCopy(&T.f[0], force.p + 3*(size_t) (S.numFree + m.z), 24u*m.w, &sm.bar[st]);
                bulkCopy(&T.localBody[0], S.localBody + lbFirst, lbBytes, &sm.bar[st]);
            }
            if (!P1ONLY && !forcesBulk(m)) requestForces(m, T, slots);
            return;
        }
        if (bulkOK(m)) {
            if (tid == 0) {
                fenceProxyAsync();                             // earlier generic-proxy writes to this stage are ordered first
                const bool fBulk = !P1ONLY && forcesBulk(m);
                mbarExpectTx(&sm.bar[st], (unsigned) ((kFPlanes + (P1ONLY ? 6 : 0))*8*m.y + 4*m.y + (fBulk ? 48 : 24)*m.w) + lbBytes);
                const double* g = S.state + (size_t) m.x;
#pragma unroll
                for (int k = 0; k < kFPlanes; k++) bulkCopy(&T.body[k][0], g + fusedGlobalPlane(k)*ld, 8u*m.y, &sm.bar[st]);
                bulkCopy(&T.loc[0], S.loc + m.x, 4u*m.y, &sm.bar[st]);
#pragma unroll
                for (int c = 0; c < 3; c++) bulkCopy(&T.d[c][0], S.dxyz + (size_t) m.z + c*as, 8u*m.w, &sm.bar[st]);
                if (P1ONLY) {
#pragma unroll
                    for (int k = 0; k < 6; k++) bulkCopy(&T.f[k*BODIES], g + ((int) PL_F + k)*ld, 8u*m.y, &sm.bar[st]);
                }
                else if (fBulk) bulkCopy(&T.f[0], force.p + 3*(size_t) (S.numFree + m.z), 24u*m.w, &sm.bar[st]);
                bulkCopy(&T.localBody[0], S.localBody + lbFirst, lbBytes, &sm.bar[st]);
            }
            if (!P1ONLY && !forcesBulk(m)) requestForces(m, T, slots);
            return;
        }
        if (tid < m.y) {
            const double* g = S.state + (size_t) (m.x + tid);
#pragma unroll
            for (int k = 0; k < kFPlanes; k++) cpAsync8(&T.body[k][tid], g + fusedGlobalPlane(k)*ld);
            cpAsync4(&T.loc[tid], S.loc + m.x + tid);
            if (P1ONLY) {
#pragma unroll
                for (int k = 0; k < 6; k++) cpAsync8(&T.f[k*BODIES + tid], g + ((int) PL_F + k)*ld);
            }
        }
        for (int j = tid; j < m.w; j += kBlock) {
            const double* g = S.dxyz + (size_t) (m.z + j);
            cpAsync8(&T.d[0][j], g);
            cpAsync8(&T.d[1][j], g + as);
            cpAsync8(&T.d[2][j], g + 2*as);
            if (!P1ONLY) {
                if (fullVelocity) cpAsync8(&T.w[j], S.atomInvMass + m.z + j);
                const double* fp = force.p + (GATHER ? (long long) slots[j] : atomSlot(S, S.numFree + m.z + j))*force.sa;
                cpAsync8(&T.f[3*j], fp);
                cpAsync8(&T.f[3*j + 1], fp + force.sc);
                cpAsync8(&T.f[3*j + 2], fp + 2*force.sc);
            }
        }
        for (unsigned w = tid; 4*w < lbBytes; w += kBlock)
            cpAsync4(&T.localBody[4*w], S.localBody + lbFirst + 4*w);
    };
    // wait for a stage filled by request(): mbarrier phase for bulk tiles, cp.async groups otherwise
    unsigned phase0 = 0u, phase1 = 0u;                         // mbarrier phase parity per stage (scalars: stay in registers)
    auto arrived = [&](int4 m, int st) {
        if (tensorOK(m) || bulkOK(m)) {
            mbarWait(&sm.bar[st], st ? phase1 : phase0);
            if (st) phase1 ^= 1u; else phase0 ^= 1u;
        }
        cpWait<0>();
    };

    // Tiles: the first G are the CTAs' own (blockIdx), every later one is claimed from a global counter, one claim ahead
    // of the prefetch - a CTA held up by a slow tile (sub-stepped bodies) simply claims fewer.  Every CTA stops at its first
    // out-of-range claim, so a launch makes exactly numTiles claims and the last one puts the counter back to zero.
    auto claim = [&]() {
        const int old = atomicAdd(S.tileCounter, 1);
        if (old == numTiles - 1) *S.tileCounter = 0;
        return G + old;
    };
    // Requests issued while tile `it` is processed: the data of tile it+1 into stage `st`; then the claim of one more
    // tile and its descriptor - GATHER: one tile further ahead (it+3), with the array slots of tile it+2 in between.
    auto advance = [&](int it, int st) {
        if (sm.tileIdx[(it + 1) % RING] < numTiles) {
            request(sm.meta[(it + 1) % RING], st, sm.slot[GATHER ? (it + 1) % 3 : 0]);
            if (!GATHER) {
                if (tid == 0) {
                    const int after = claim();
                    sm.tileIdx[(it + 2) % RING] = after;
                    if (after < numTiles) cpAsync16(&sm.meta[(it + 2) % RING], tileMeta + after);
                }
            }
            else if (sm.tileIdx[(it + 2) % RING] < numTiles) {
                requestSlots(sm.meta[(it + 2) % RING], (it + 2) % 3);
                if (tid == 0) {
                    const int after = claim();
                    sm.tileIdx[(it + 3) % RING] = after;
                    if (after < numTiles) cpAsync16(&sm.meta[(it + 3) % RING], tileMeta + after);
                }
            }
        }
        cpCommit();
    };
    const int tile0 = blockIdx.x;
    if (tile0 < numTiles) {
        if (tid == 0) {
            const int tile1 = claim();
            sm.tileIdx[0] = tile0;
            sm.tileIdx[1] = tile1;
            sm.meta[0] = tileMeta[tile0];
            if (tile1 < numTiles) sm.meta[1] = tileMeta[tile1];
            if (GATHER) {
                const int tile2 = tile1 < numTiles ? claim() : tile1;      // a CTA stops claiming at its first out-of-range claim
                sm.tileIdx[2] = tile2;
                if (tile2 < numTiles) sm.meta[2] = tileMeta[tile2];
            }
            mbarInit(&sm.bar[0], 1);
            mbarInit(&sm.bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (GATHER) {
            requestSlots(sm.meta[0], 0);
            if (sm.tileIdx[1] < numTiles) requestSlots(sm.meta[1], 1);
            cpCommit();
            cpWait<0>();
            __syncthreads();
        }
        request(sm.meta[0], 0, sm.slot[0]);
        cpCommit();
        for (int it = 0; sm.tileIdx[it % RING] < numTiles; it++) {
            const int4 m = sm.meta[it % RING];
            const int cur = STAGES == 2 ? (it & 1) : 0;
            Stage& T = sm.stage[cur];
            const int* const slots = sm.slot[GATHER ? it % 3 : 0];
            if (!SMALL && !P1ONLY) {
#pragma unroll
                for (int k = 0; k < 6; k++) sm.acc[k][tid] = 0.0;
                if (tid < kWarps*6) sm.head[tid/6][tid%6] = 0.0;
            }
            arrived(m, cur);                                   // this tile (+ the next tile's descriptor) landed
            __syncthreads();
            if (STAGES == 2) advance(it, cur ^ 1);

            // ---- B: forces and torques -> per-body sums.
            // SMALL (every body has <= kSmallBody atoms, e.g. water): each body's thread sums its own atoms straight
            // from the staged forces/coordinates in phase C - sequential order, no shuffles, no extra barrier.
            // Otherwise: thread per atom + warp-shuffle segmented scan, exactly as in part2Kernel.
            const int shift = m.z & 15;
            if (!SMALL && !P1ONLY) {
                const int per = ((m.w + kBlock - 1)/kBlock)*32;
                const int wBeg = warp*per, wEnd = min(wBeg + per, m.w);        // tile-local atom indices
                int firstKey = -1;
                if (wBeg < wEnd) {
                    const int k0 = T.localBody[wBeg + shift];
                    if (T.loc[k0] - m.z < wBeg) firstKey = k0;
                }
                if (lane == 0) sm.headKey[warp] = firstKey;
                for (int base = wBeg; base < wEnd; base += 32) {
                    const int j = base + lane;
                    const bool valid = j < wEnd;
                    int key = 0x7fffffff;
                    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                    if (valid) {
                        key = T.localBody[j + shift];
                        const d3 d = {T.d[0][j], T.d[1][j], T.d[2][j]};
                        const d3 f = {stagedForce<NATIVE>(force, T.f[3*j]), stagedForce<NATIVE>(force, T.f[3*j + 1]), stagedForce<NATIVE>(force, T.f[3*j + 2])};
                        const d4 q = {T.body[6][key], T.body[7][key], T.body[8][key], T.body[9][key]};
                        const d3 delta = bodyToSpace(q, d);
                        T.f[3*j] = delta.x; T.f[3*j + 1] = delta.y; T.f[3*j + 2] = delta.z;     // kept for the velocities
                        const d3 t = cross(delta, f);
                        v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = t.x; v[4] = t.y; v[5] = t.z;
                    }
                    for (int off = 1; off < 32 && off < S.maxBodySize; off <<= 1) {
                        const int kp = __shfl_up_sync(kFull, key, off);
                        const bool take = lane >= off && kp == key;
    #pragma unroll
                        for (int k = 0; k < 6; k++) {
                            const double t = __shfl_up_sync(kFull, v[k], off);
                            if (take) v[k] += t;
                        }
                    }
                    const int kn = __shfl_down_sync(kFull, key, 1);
                    if (valid && (lane == 31 || kn != key)) {
                        if (key == firstKey) {
    #pragma unroll
                            for (int k = 0; k < 6; k++) sm.head[warp][k] += v[k];
                        }
                        else {
    #pragma unroll
                            for (int k = 0; k < 6; k++) sm.acc[k][key] += v[k];
                        }
                    }
                    __syncwarp();
                }
                __syncthreads();

            }
            // ---- C: thread per body: second kick of this step, then first kick + drift + rotation of the next
            double (*B)[kBlock] = T.body;
            unsigned flags = 0u;
            if (tid < m.y) {
                d3 r = {B[0][tid], B[1][tid], B[2][tid]};
                d3 p = {B[3][tid], B[4][tid], B[5][tid]};
                d4 q = {B[6][tid], B[7][tid], B[8][tid], B[9][tid]};
                d3 F = {0.0, 0.0, 0.0}, tau = {0.0, 0.0, 0.0};
                if (P1ONLY) {
                    F = {T.f[tid], T.f[BODIES + tid], T.f[2*BODIES + tid]};
                    tau = {T.f[3*BODIES + tid], T.f[4*BODIES + tid], T.f[5*BODIES + tid]};
                }
                else if (SMALL && kTriatomic && m.w == 3*m.y) {
                    // a tile of three-atom bodies (water): the same sums, ((f0 + f1) + f2 like the loop below), with the three
                    // atoms' rotations written out side by side - three independent dependency chains instead of one
                    const int j = 3*tid;
                    d3 f[3], delta[3];
#pragma unroll
                    for (int u = 0; u < 3; u++) {
                        const d3 d = {T.d[0][j + u], T.d[1][j + u], T.d[2][j + u]};
                        f[u] = {stagedForce<NATIVE>(force, T.f[3*(j + u)]), stagedForce<NATIVE>(force, T.f[3*(j + u) + 1]),
                                stagedForce<NATIVE>(force, T.f[3*(j + u) + 2])};
                        delta[u] = bodyToSpace(q, d);
                    }
#pragma unroll
                    for (int u = 0; u < 3; u++) {
                        T.f[3*(j + u)] = delta[u].x; T.f[3*(j + u) + 1] = delta[u].y; T.f[3*(j + u) + 2] = delta[u].z;
                    }
                    F = (f[0] + f[1]) + f[2];
                    tau = (cross(delta[0], f[0]) + cross(delta[1], f[1])) + cross(delta[2], f[2]);
                }
                else if (SMALL) {
                    const int j0 = T.loc[tid] - m.z, j1 = (tid + 1 < m.y ? T.loc[tid + 1] - m.z : m.w);
                    for (int j = j0; j < j1; j++) {
                        const d3 d = {T.d[0][j], T.d[1][j], T.d[2][j]};
                        const d3 f = {stagedForce<NATIVE>(force, T.f[3*j]), stagedForce<NATIVE>(force, T.f[3*j + 1]), stagedForce<NATIVE>(force, T.f[3*j + 2])};
                        const d3 delta = bodyToSpace(q, d);
                        T.f[3*j] = delta.x; T.f[3*j + 1] = delta.y; T.f[3*j + 2] = delta.z;     // kept for the velocities
                        F = F + f;
                        tau = tau + cross(delta, f);
                    }
                }
                else {
                    double sum[6];
#pragma unroll
                    for (int k = 0; k < 6; k++) sum[k] = sm.acc[k][tid];
#pragma unroll
                    for (int w = 1; w < kWarps; w++)
                        if (sm.headKey[w] == tid) {
#pragma unroll
                            for (int k = 0; k < 6; k++) sum[k] += sm.head[w][k];
                        }
                    F = {sum[0], sum[1], sum[2]};
                    tau = {sum[3], sum[4], sum[5]};
                }
                d4 pi = {B[10][tid], B[11][tid], B[12][tid], B[13][tid]};
                const double invm = B[14][tid];
                const d3 invI = {B[15][tid], B[16][tid], B[17][tid]};
                if (!P1ONLY) {
                    d3 vcm, om;
                    bodyPart2(dt, F, tau, invm, invI, q, p, pi, vcm, om);
                    sm.acc[0][tid] = vcm.x; sm.acc[1][tid] = vcm.y; sm.acc[2][tid] = vcm.z;
                    sm.acc[3][tid] = om.x; sm.acc[4][tid] = om.y; sm.acc[5][tid] = om.z;
                }
                if (LADDER == 2) bodyPart1Ladder(0, dt, F, tau, invm, invI, r, p, q, pi, flags);
                else if (LADDER == 1) bodyPart1Ladder(rung, dt, F, tau, invm, invI, r, p, q, pi, flags);
                else bodyPart1<EXACT>(dt, S.rotationMode, F, tau, invm, invI, r, p, q, pi);
                double* s = S.state + (size_t) (m.x + tid);
                storePlane3(s + PL_R*ld, ld, r);
                storePlane3(s + PL_P*ld, ld, p);
                storePlane4(s + PL_Q*ld, ld, q);
                storePlane4(s + PL_PI*ld, ld, pi);
                if (!P1ONLY && storeFT) {
                    storePlane3(s + PL_F*ld, ld, F);
                    storePlane3(s + PL_TAU*ld, ld, tau);
                }
                B[0][tid] = r.x; B[1][tid] = r.y; B[2][tid] = r.z;
                B[6][tid] = q.w; B[7][tid] = q.x; B[8][tid] = q.y; B[9][tid] = q.z;
            }
            if (LADDER) {
                ladderFails += __popc(__ballot_sync(kFull, (flags & 1u) != 0u));
                ladderLower += __popc(__ballot_sync(kFull, (flags & 2u) != 0u));
            }
            __syncthreads();

            // ---- D: thread per atom: velocities at the end of this step, positions of the next
            auto atomOut = [&](int j) {
                const int k = T.localBody[j + shift] & (BODIES - 1);      // index inside the 128-body atom tile -> this tile
                const long long slot = GATHER ? (long long) slots[j] : atomSlot(S, S.numFree + m.z + j);
                if (!P1ONLY) {
                    const d3 delta = {T.f[3*j], T.f[3*j + 1], T.f[3*j + 2]};
                    const d3 vcm = {sm.acc[0][k], sm.acc[1][k], sm.acc[2][k]};
                    const d3 om = {sm.acc[3][k], sm.acc[4][k], sm.acc[5][k]};
                    const d3 vv = atomVelocity(vcm, om, delta);
                    if (fullVelocity) {
                        double2* out = reinterpret_cast<double2*>(vel.p + 4*slot);
                        out[0] = make_double2(vv.x, vv.y);
                        out[1] = make_double2(vv.z, T.w[GATHER ? j : 0]);
                    }
                    else storeAtom<NATIVE>(vel, slot, vv);
                }
                const d3 d = {T.d[0][j], T.d[1][j], T.d[2][j]};
                const d4 q = {B[6][k], B[7][k], B[8][k], B[9][k]};
                const d3 r = {B[0][k], B[1][k], B[2][k]};
                storeAtom<NATIVE>(pos, slot, atomPosition(r, q, d));
            };
            if (kTriatomic && m.w == 96) {                     // a full tile of waters: three atoms per thread, written out
#pragma unroll
                for (int u = 0; u < 3; u++) atomOut(tid + 32*u);
            }
            else for (int j = tid; j < m.w; j += kBlock) atomOut(j);
            __syncthreads();
            if (STAGES == 1) advance(it, 0);                   // the single stage is free again: request the next tile
        }
        cpWait<0>();
    }
    if (LADDER && tid == 0) {
        // publish this CTA's counts; the last CTA of the launch moves the rung for the next launch: up when more than 4e-4
        // of the bodies needed the retry path, down when fewer than 1e-4 would need it one rung lower
        SeriesControl* c = S.seriesCtl;
        if (ladderFails | ladderLower) {                       // (rare; the fence orders the counts before this CTA's `done` -
            if (ladderFails) atomicAdd(&c->fails, ladderFails);       // a CTA without counts must not wait for its own last
            if (ladderLower) atomicAdd(&c->lower, ladderLower);       // position / velocity stores to drain)
            __threadfence();
        }
        if (atomicAdd(&c->done, 1u) == gridDim.x - 1) {
            __threadfence();
            const volatile unsigned* counts = &c->fails;       // (both loads in flight together: one round trip at the launch's tail)
            const unsigned fails = counts[0], lower = counts[1];
            const double n = (double) S.numBodies;
            int next = rung;
            if ((double) fails > 4.0e-4*n && rung < 2) next = rung + 1;
            else if (rung > 0 && (double) lower < 1.0e-4*n) next = rung - 1;
            c->rung = next;
            if (next != c->published) {                        // the host's hint for its choice of kernel (mapped pinned memory)
                *S.hostRungDevice = next;
                c->published = next;
            }
            c->fails = 0u;
            c->lower = 0u;
            c->done = 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Free atoms: velocity Verlet, one thread per atom, its own launch so that it runs at full occupancy next to
// the persistent body kernels.  PHASE 1 = part 1 (half kick + drift), 2 = part 2 (half kick + constraint
// displacement), 3 = part 2 of this step followed by part 1 of the next.
// ------------------------------------------------------------------------------------------------
// PER atoms per thread (k, k + blockDim, ...): every load of all of them is issued before the first is used - the side-stream
// launches get one or two small CTAs per SM next to a persistent body kernel and live on the loads they keep in flight.
template <int PHASE, bool NATIVE, int PER>
__global__ void __launch_bounds__(256) freeAtomsKernel(const DeviceSystem S, const double dt, const AtomView pos, const AtomView vel,
                                                       const AtomView force) {
    const int k0 = blockIdx.x*blockDim.x*PER + threadIdx.x;
    long long gi[PER];
    d3 f[PER], x[PER], v[PER], saved[PER];
    double invm[PER];
#pragma unroll
    for (int u = 0; u < PER; u++) {
        const int k = k0 + u*blockDim.x;
        gi[u] = k < S.numFree ? atomSlot(S, k) : 0;
    }
#pragma unroll
    for (int u = 0; u < PER; u++) {
        const int k = k0 + u*blockDim.x;
        if (k < S.numFree) {
            f[u] = loadAtom<NATIVE>(force, gi[u]);
            invm[u] = S.freeInvMass[k];
            x[u] = loadAtom<NATIVE>(pos, gi[u]);
            v[u] = loadAtom<NATIVE>(vel, gi[u]);
            if (PHASE & 2) saved[u] = loadPlane3(S.savedPos + k, S.freeStride);
        }
    }
#pragma unroll
    for (int u = 0; u < PER; u++) {
        const int k = k0 + u*blockDim.x;
        if (k >= S.numFree) continue;
        if (PHASE & 2) freePart2(dt, f[u], invm[u], x[u], saved[u], v[u]);
        if (PHASE & 1) {
            freePart1(dt, f[u], invm[u], x[u], v[u]);
            storeAtom<NATIVE>(pos, gi[u], x[u]);
            storePlane3(S.savedPos + k, S.freeStride, asStored<NATIVE>(pos, x[u]));
        }
        if (!NATIVE && vel.fmt == FMT_REAL4_F64 && S.atomInvMass != nullptr) {      // whole double4, w = 1/m (see part2Part1Kernel)
            double2* out = reinterpret_cast<double2*>(vel.p + 4*gi[u]);
            out[0] = make_double2(v[u].x, v[u].y);
            out[1] = make_double2(v[u].z, invm[u]);
        }
        else storeAtom<NATIVE>(vel, gi[u], v[u]);
    }
}

#ifndef RBK_SIDE_FREE_THREADS
#define RBK_SIDE_FREE_THREADS 64
#endif
constexpr int kSideFreeThreads = RBK_SIDE_FREE_THREADS;                // CTA size of the free-atom launch that shares the SMs with the body kernels
#ifndef RBK_SIDE_FREE_PER
#define RBK_SIDE_FREE_PER 2
#endif
constexpr int kSideFreePer = RBK_SIDE_FREE_PER;                        // ... and its atoms per thread

template <int PHASE, bool NATIVE>
cudaError_t launchFree(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st) {
    if (S.numFree == 0) return cudaSuccess;
    freeAtomsKernel<PHASE, NATIVE, 1><<<(S.numFree + 255)/256, 256, 0, st>>>(S, dt, pos, vel, force);
    return cudaGetLastError();
}

// Free atoms of a large-body launch sequence on the side stream: fork before the body kernels are launched, the free-atom
// kernel after the first (persistent) body kernel has taken its SMs, join when the sequence is complete.
inline bool sideUsable(const DeviceSystem& S, const SideStream* side) {
    return side != nullptr && side->stream != nullptr && S.numFree > 0 && S.numTiles > 0 && S.splitPart1;
}
inline cudaError_t sideFork(const SideStream* side, cudaStream_t st) {
    cudaError_t e = cudaEventRecord(side->fork, st);
    return e != cudaSuccess ? e : cudaStreamWaitEvent(side->stream, side->fork, 0);
}
template <int PHASE, bool NATIVE>
cudaError_t sideFree(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, const SideStream* side) {
    constexpr int perBlock = kSideFreeThreads*kSideFreePer;
    freeAtomsKernel<PHASE, NATIVE, kSideFreePer><<<(S.numFree + perBlock - 1)/perBlock, kSideFreeThreads, 0, side->stream>>>(S, dt, pos, vel, force);
    const cudaError_t e = cudaGetLastError();
    return e != cudaSuccess ? e : cudaEventRecord(side->join, side->stream);
}
inline cudaError_t sideJoin(const SideStream* side, cudaStream_t st) { return cudaStreamWaitEvent(st, side->join, 0); }

// Free-atom constraint hooks (the reference's freeAtomsDelta pre-pass, rigidbodyintegrator.cu:276-285, and the free-atom
// loop of its integrateRigidBodyPart1, :303-312).  CONSUME = false: delta = (v + f invm dt/2) dt, nothing else is
// touched, so the caller's constraint solver can correct the displacement.  CONSUME = true: the first half kick of the
// velocity, x += delta, savedPos = x.
template <bool CONSUME>
__global__ void __launch_bounds__(256) freeDeltaKernel(const DeviceSystem S, const double dt, const AtomView pos, const AtomView vel,
                                                       const AtomView force, const AtomView delta) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= S.numFree) return;
    const long long gi = atomSlot(S, k);
    const d3 f = loadAtom<false>(force, gi);
    d3 v = loadAtom<false>(vel, gi);
    v = v + f*S.freeInvMass[k]*(0.5*dt);
    if (!CONSUME) {
        storeAtom<false>(delta, gi, d3{__dmul_rn(v.x, dt), __dmul_rn(v.y, dt), __dmul_rn(v.z, dt)});
        return;
    }
    // x moves by the (constrained) displacement; savedPos remembers where the UNCONSTRAINED step would have put it, so that
    // Part 2's (x - savedPos)/dt hands the constraint displacement to the velocity - the Reference platform's arithmetic
    // (RigidBodySystem.cpp:172-176,196-197 with ReferenceConstraints::apply in between).  The reference's CUDA kernel saves
    // the constrained position instead (rigidbodyintegrator.cu:303-312), which drops that term: constrained free atoms
    // then lose the centripetal part of their velocity change every step and cool down.  Without a solver the two
    // displacements are the same bits (round(v dt), no FMA contraction) and the term is exactly zero.
    const d3 x0 = loadAtom<false>(pos, gi);
    const d3 du = {__dmul_rn(v.x, dt), __dmul_rn(v.y, dt), __dmul_rn(v.z, dt)};
    const d3 x = x0 + loadAtom<false>(delta, gi);
    storeAtom<false>(pos, gi, x);
    storePlane3(S.savedPos + k, S.freeStride, asStored<false>(pos, x0 + du));
    storeAtom<false>(vel, gi, v);
}

// ------------------------------------------------------------------------------------------------
// Kinetic energies: fixed-shape two-level tree (warp shuffles -> shared -> last CTA), no atomics
// on the data path, so the two doubles are bit-reproducible from run to run.
// ------------------------------------------------------------------------------------------------
constexpr int kKinThreads = 256;

__device__ __forceinline__ void blockSum2(double& a, double& b, double (*scratch)[2]) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(kFull, a, off);
        b += __shfl_xor_sync(kFull, b, off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { scratch[warp][0] = a; scratch[warp][1] = b; }
    __syncthreads();
    a = 0.0; b = 0.0;
    for (int w = 0; w < kKinThreads/32; w++) { a += scratch[w][0]; b += scratch[w][1]; }
}

template <bool NATIVE>
__global__ void __launch_bounds__(kKinThreads) kineticKernel(const DeviceSystem S, const AtomView vel, double* partial,
                                                            unsigned* counter, double* out) {
    __shared__ double scratch[kKinThreads/32][2];
    __shared__ bool isLast;
    const int g = blockIdx.x*kKinThreads + threadIdx.x, T = gridDim.x*kKinThreads;
    double kt = 0.0, kr = 0.0;
    const size_t ld = S.bodyStride;
    for (int b = g; b < S.numBodies; b += T) {
        const double* s = S.state + b;
        double t2, r2;
        bodyKinetic(loadPlane3(s + PL_P*ld, ld), loadPlane4(s + PL_Q*ld, ld), loadPlane4(s + PL_PI*ld, ld),
                    s[PL_INVM*ld], loadPlane3(s + PL_INVI*ld, ld), t2, r2);
        kt += t2;
        kr += r2;
    }
    for (int k = g; k < S.numFree; k += T)
        kt += freeKinetic(loadAtom<NATIVE>(vel, atomSlot(S, k)), S.freeInvMass[k]);
    blockSum2(kt, kr, scratch);
    if (threadIdx.x == 0) {
        partial[2*blockIdx.x] = kt;
        partial[2*blockIdx.x + 1] = kr;
        __threadfence();
        isLast = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (isLast) {
        __threadfence();
        kt = 0.0; kr = 0.0;
        for (int i = threadIdx.x; i < (int) gridDim.x; i += kKinThreads) {
            kt += __ldcg(partial + 2*i);
            kr += __ldcg(partial + 2*i + 1);
        }
        blockSum2(kt, kr, scratch);
        if (threadIdx.x == 0) {
            out[0] = 0.5*kt;
            out[1] = 0.5*kr;
            *counter = 0u;
        }
    }
}

template <bool EXACT, bool FUSED, bool NATIVE>
cudaError_t launchPart1Variant(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, bool freeAtoms, cudaStream_t st,
                               const SideStream* side = nullptr) {
    const size_t smem = FUSED ? sizeof(Part1Smem) : offsetof(Part1Smem, d);
    // exact rotation on body tiles (large bodies): the kernel compiled for the rung the device last published (a hint that may be a
    // launch old; every rung is a complete algorithm - bodies that fail its check take the retry path)
    const int rung = (EXACT && !FUSED) ? (S.fullLadderOnly ? 2 : *S.hostRung) : -1;
    static LaunchCache cache, cacheRung[3];
    int blocks = 0;
    cudaError_t e = rung < 0 ? cache.get(part1Kernel<EXACT, FUSED, NATIVE>, kBlock, smem, blocks)
                  : rung == 0 ? cacheRung[0].get(part1Kernel<EXACT, false, true, EXACT && !FUSED ? 0 : -1>, kBlock, smem, blocks)
                  : rung == 1 ? cacheRung[1].get(part1Kernel<EXACT, false, true, EXACT && !FUSED ? 1 : -1>, kBlock, smem, blocks)
                              : cacheRung[2].get(part1Kernel<EXACT, false, true, EXACT && !FUSED ? 2 : -1>, kBlock, smem, blocks);
    if (e != cudaSuccess) return e;
    // persistent CTAs: one wave that fills every SM
    const bool overlap = !FUSED && freeAtoms && sideUsable(S, side);
    if (overlap) e = sideFork(side, st);
    else if (freeAtoms) e = launchFree<1, NATIVE>(S, dt, pos, vel, force, st);
    if (e != cudaSuccess) return e;
    const int tiles = FUSED ? S.numTiles : S.numBodyTiles;
    const int resident = S.numSMs*(EXACT ? (rung >= 0 ? RBK_P1_MINBLOCKS_ROTATION : RBK_P1_MINBLOCKS_EXACT) : RBK_P1_MINBLOCKS_SPLIT);
    if (tiles > 0) {
        const int grid = tiles < resident ? tiles : resident;
        if (rung < 0) part1Kernel<EXACT, FUSED, NATIVE><<<grid, kBlock, smem, st>>>(S, dt, pos, vel, force);
        else if (rung == 0) part1Kernel<EXACT, false, true, EXACT && !FUSED ? 0 : -1><<<grid, kBlock, smem, st>>>(S, dt, pos, vel, force);
        else if (rung == 1) part1Kernel<EXACT, false, true, EXACT && !FUSED ? 1 : -1><<<grid, kBlock, smem, st>>>(S, dt, pos, vel, force);
        else part1Kernel<EXACT, false, true, EXACT && !FUSED ? 2 : -1><<<grid, kBlock, smem, st>>>(S, dt, pos, vel, force);
        e = launchResult(S, st);
        if (e != cudaSuccess) return e;
    }
    if (overlap) {
        e = sideFree<1, NATIVE>(S, dt, pos, vel, force, side);
        if (e != cudaSuccess) return e;
    }
    if (!FUSED && S.numTiles > 0) atomPositionKernel<NATIVE><<<S.numTiles, kBlock, 0, st>>>(S, pos);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return overlap ? sideJoin(side, st) : cudaSuccess;
}

template <bool EXACT, bool SMALL, int BODIES, int ATOMS, int STAGES, bool P1ONLY, bool GATHER, int LADDER>
cudaError_t launchFusedLadder(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st,
                              bool freeAtoms, bool storeFT);

template <bool EXACT, bool SMALL, int BODIES, int ATOMS, int STAGES, bool P1ONLY, bool GATHER>
cudaError_t launchFusedShape(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st,
                             bool freeAtoms = true, bool storeFT = true) {
#ifndef RBK_EXPERIMENT_NOLADDER
    if (EXACT && BODIES == 32) {
        // the host's copy of the rung (written by the kernels through mapped memory) picks the lean rung-0 kernel or the full ladder
        if (*S.hostRung == 0 && !S.fullLadderOnly) return launchFusedLadder<EXACT, SMALL, BODIES, ATOMS, STAGES, P1ONLY, GATHER, EXACT && BODIES == 32 ? 2 : 0>(S, dt, pos, vel, force, st, freeAtoms, storeFT);
        return launchFusedLadder<EXACT, SMALL, BODIES, ATOMS, STAGES, P1ONLY, GATHER, EXACT && BODIES == 32 ? 1 : 0>(S, dt, pos, vel, force, st, freeAtoms, storeFT);
    }
#endif
    return launchFusedLadder<EXACT, SMALL, BODIES, ATOMS, STAGES, P1ONLY, GATHER, 0>(S, dt, pos, vel, force, st, freeAtoms, storeFT);
}

template <bool EXACT, bool SMALL, int BODIES, int ATOMS, int STAGES, bool P1ONLY, bool GATHER, int LADDER>
cudaError_t launchFusedLadder(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st,
                              bool freeAtoms, bool storeFT) {
    typedef FusedSmem<BODIES, ATOMS, STAGES, GATHER> Smem;
    constexpr bool NATIVE = !GATHER;
    auto kernel = part2Part1Kernel<EXACT, SMALL, NATIVE, BODIES, ATOMS, STAGES, P1ONLY, GATHER, LADDER>;
    static LaunchCache cache;
    int blocks = 0;                                            // persistent CTAs: one full wave, whatever fits
    cudaError_t e = cache.get(kernel, BODIES, sizeof(Smem), blocks);
    if (e != cudaSuccess) return e;
    if (freeAtoms) {
        e = launchFree<P1ONLY ? 1 : 3, NATIVE>(S, dt, pos, vel, force, st);
        if (e != cudaSuccess) return e;
    }
    const int tiles = BODIES == 32 ? S.numWarpTiles : S.numTiles;
    const int resident = S.numSMs*blocks;
    if (tiles > 0) {
        static const TileMaps noMaps = {};
        const bool useMaps = BODIES == 32 && S.tileMaps != nullptr;
        kernel<<<tiles < resident ? tiles : resident, BODIES, sizeof(Smem), st>>>(S, dt, pos, vel, force, useMaps ? *S.tileMaps : noMaps,
                                                                                  useMaps, storeFT);
    }
    return launchResult(S, st);
}

template <bool EXACT, bool SMALL, bool GATHER>
cudaError_t launchFused(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st, bool storeFT) {
    // One warp per CTA pays off where a tile's phases are long and uneven (exact rotation: 0.1445 -> 0.137 ms at 1 M
    // waters).  NO-SQUISH (112 registers) instead runs four single-stage CTAs of four warps per SM: mode 10
    // 0.177 -> 0.150 ms, mode 3 0.110 ms; one-warp single-stage CTAs measured 0.147 / 0.118 ms, so the four-warp shape stays.
    if (EXACT && SMALL && S.numWarpTiles > 0)
        return launchFusedShape<EXACT, true, 32, kWarpTileAtoms, 2, false, GATHER>(S, dt, pos, vel, force, st, true, storeFT);
    return launchFusedShape<EXACT, SMALL, kBlock, kTileAtoms, EXACT ? 2 : 1, false, GATHER>(S, dt, pos, vel, force, st, true, storeFT);
}

bool nativeIO(const AtomView& a, const AtomView& b, const AtomView& c) {
    return a.fmt == FMT_F64 && b.fmt == FMT_F64 && c.fmt == FMT_F64;
}

template <bool NATIVE>
cudaError_t launchPart1Formats(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, bool freeAtoms, cudaStream_t st,
                               const SideStream* side = nullptr) {
    const bool exact = S.rotationMode == 0, fused = !S.splitPart1;
    // exact rotation on water-like bodies: Part 1 alone through the TMA-staged one-warp-tile pipeline (fp64 arrays in
    // plugin order: the plain instantiation; reordered atoms / OpenMM formats: the GATHER one)
    if (exact && S.numWarpTiles > 0) {
        if (NATIVE && S.atomLoc == nullptr) return launchFusedShape<true, true, 32, kWarpTileAtoms, 2, true, false>(S, dt, pos, vel, force, st, freeAtoms);
        return launchFusedShape<true, true, 32, kWarpTileAtoms, 2, true, true>(S, dt, pos, vel, force, st, freeAtoms);
    }
    if (exact) return fused ? launchPart1Variant<true, true, NATIVE>(S, dt, pos, vel, force, freeAtoms, st) : launchPart1Variant<true, false, NATIVE>(S, dt, pos, vel, force, freeAtoms, st, side);
    return fused ? launchPart1Variant<false, true, NATIVE>(S, dt, pos, vel, force, freeAtoms, st) : launchPart1Variant<false, false, NATIVE>(S, dt, pos, vel, force, freeAtoms, st, side);
}

} // namespace

cudaError_t launchPart1(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st, const SideStream* side) {
    if (S.numTiles + S.numFreeBlocks == 0) return cudaSuccess;
    return nativeIO(pos, vel, force) ? launchPart1Formats<true>(S, dt, pos, vel, force, true, st, side) : launchPart1Formats<false>(S, dt, pos, vel, force, true, st, side);
}

cudaError_t launchFreeDelta(const DeviceSystem& S, double dt, AtomView vel, AtomView force, AtomView delta, cudaStream_t st) {
    if (S.numFree > 0) freeDeltaKernel<false><<<(S.numFree + 255)/256, 256, 0, st>>>(S, dt, vel, vel, force, delta);
    return cudaGetLastError();
}

cudaError_t launchPart1Delta(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, AtomView delta, cudaStream_t st) {
    if (S.numFree > 0) freeDeltaKernel<true><<<(S.numFree + 255)/256, 256, 0, st>>>(S, dt, pos, vel, force, delta);
    if (S.numTiles == 0) return cudaGetLastError();
    return nativeIO(pos, vel, force) ? launchPart1Formats<true>(S, dt, pos, vel, force, false, st) : launchPart1Formats<false>(S, dt, pos, vel, force, false, st);
}

namespace {
template <bool NATIVE>
cudaError_t launchPart2Formats(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, bool freeAtoms, cudaStream_t st,
                               const SideStream* side = nullptr, int ridePhase = 2) {
    // large bodies: the free atoms ride along in the body kernel when their list splits into small enough slices
    if (freeAtoms && freeAtomsRide(S)) return launchPart2Large<NATIVE>(S, dt, pos, vel, force, ridePhase, st);
    const bool overlap = freeAtoms && sideUsable(S, side);
    if (overlap) {
        cudaError_t e = sideFork(side, st);
        if (e != cudaSuccess) return e;
    }
    else if (freeAtoms) {
        cudaError_t e = launchFree<2, NATIVE>(S, dt, pos, vel, force, st);
        if (e != cudaSuccess) return e;
    }
    if (S.numTiles > 0 && S.splitPart1) {
        cudaError_t e = launchPart2Large<NATIVE>(S, dt, pos, vel, force, 0, st);
        if (e != cudaSuccess || !overlap) return e;
        e = sideFree<2, NATIVE>(S, dt, pos, vel, force, side);
        return e != cudaSuccess ? e : sideJoin(side, st);
    }
    if (S.numTiles > 0) part2Kernel<NATIVE><<<S.numTiles, kBlock, 0, st>>>(S, dt, pos, vel, force);
    return cudaGetLastError();
}
} // namespace

cudaError_t launchPart2(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st, const SideStream* side) {
    if (S.numTiles + S.numFreeBlocks == 0) return cudaSuccess;
    return nativeIO(pos, vel, force) ? launchPart2Formats<true>(S, dt, pos, vel, force, true, st, side)
                                     : launchPart2Formats<false>(S, dt, pos, vel, force, true, st, side);
}

cudaError_t launchPart2Part1(const DeviceSystem& S, double dt, AtomView pos, AtomView vel, AtomView force, cudaStream_t st,
                             const SideStream* side) {
    if (S.numTiles + S.numFreeBlocks == 0) return cudaSuccess;
    // large bodies take separate kernels (the one-pass kernel stages a whole tile's atoms in shared memory)
    if (!S.fusable) {
        // the free atoms still take both halves in ONE launch (same arithmetic, one pass over their data).  Large-body
        // systems: that launch goes to the side stream, AFTER the persistent part2LargeKernel has taken its SMs - the
        // free atoms' small CTAs run in the registers the body kernels leave unused and in their tails (disjoint atoms,
        // no data dependence); the caller's stream continues when both are done.
        const bool native = nativeIO(pos, vel, force);
        if (freeAtomsRide(S)) {                                // ... or ride along in the body kernel, both halves at once
            cudaError_t e = native ? launchPart2Formats<true>(S, dt, pos, vel, force, true, st, nullptr, 3)
                                   : launchPart2Formats<false>(S, dt, pos, vel, force, true, st, nullptr, 3);
            if (e != cudaSuccess) return e;
            return native ? launchPart1Formats<true>(S, dt, pos, vel, force, false, st) : launchPart1Formats<false>(S, dt, pos, vel, force, false, st);
        }
        const bool overlap = sideUsable(S, side);
        if (overlap) {
            cudaError_t e = sideFork(side, st);
            if (e != cudaSuccess) return e;
        }
        else {
            cudaError_t e = native ? launchFree<3, true>(S, dt, pos, vel, force, st) : launchFree<3, false>(S, dt, pos, vel, force, st);
            if (e != cudaSuccess) return e;
        }
        if (S.numTiles == 0) return cudaSuccess;
        cudaError_t e = native ? launchPart2Formats<true>(S, dt, pos, vel, force, false, st) : launchPart2Formats<false>(S, dt, pos, vel, force, false, st);
        if (e != cudaSuccess) return e;
        if (overlap) {
            e = native ? sideFree<3, true>(S, dt, pos, vel, force, side) : sideFree<3, false>(S, dt, pos, vel, force, side);
            if (e != cudaSuccess) return e;
        }
        e = native ? launchPart1Formats<true>(S, dt, pos, vel, force, false, st) : launchPart1Formats<false>(S, dt, pos, vel, force, false, st);
        if (e != cudaSuccess) return e;
        return overlap ? sideJoin(side, st) : cudaSuccess;
    }
    const bool small = S.maxBodySize <= kSmallBody;
    const bool storeFT = !S.lazyForceTorque;
    if (nativeIO(pos, vel, force) && S.atomLoc == nullptr) {
        if (S.rotationMode == 0) return small ? launchFused<true, true, false>(S, dt, pos, vel, force, st, storeFT) : launchFused<true, false, false>(S, dt, pos, vel, force, st, storeFT);
        return small ? launchFused<false, true, false>(S, dt, pos, vel, force, st, storeFT) : launchFused<false, false, false>(S, dt, pos, vel, force, st, storeFT);
    }
    if (S.rotationMode == 0) return small ? launchFused<true, true, true>(S, dt, pos, vel, force, st, storeFT) : launchFused<true, false, true>(S, dt, pos, vel, force, st, storeFT);
    return small ? launchFused<false, true, true>(S, dt, pos, vel, force, st, storeFT) : launchFused<false, false, true>(S, dt, pos, vel, force, st, storeFT);
}

bool part2Part1LeavesForceTorque(const DeviceSystem& S) { return !(S.fusable && S.lazyForceTorque) || S.numTiles == 0; }

cudaError_t launchKinetic(const DeviceSystem& S, AtomView vel, double* partial, unsigned* counter, double* out,
                          cudaStream_t st) {
    if (vel.fmt == FMT_F64) kineticKernel<true><<<kKineticBlocks, kKinThreads, 0, st>>>(S, vel, partial, counter, out);
    else kineticKernel<false><<<kKineticBlocks, kKinThreads, 0, st>>>(S, vel, partial, counter, out);
    return cudaGetLastError();
}

int part1LaunchesPerStep(const DeviceSystem& S) { return (S.splitPart1 && S.numTiles > 0) ? 2 : 1; }

void launchesPerCall(const DeviceSystem& S, int out[3]) {
    const int bodies = S.numTiles > 0, freeAtoms = S.numFree > 0, large = bodies && S.splitPart1;
    const int ride = freeAtomsRide(S);                         // large bodies: the free atoms ride along in part2LargeKernel
    out[0] = freeAtoms + (large ? 2 : bodies);                 // [free] + rotation kernel + atomPositionKernel | one kernel
    out[1] = (freeAtoms && !ride) + bodies;
    out[2] = (freeAtoms && !ride) + (large ? 3 : bodies);
}

} // namespace rbk
