// rbk_refined.cu - refined ("shadow") energy diagnostics.
//
// What the reference computes only on its CUDA platform when setComputeRefinedEnergies(true)
// (COMPMOD paths of platforms/cuda/src/kernels/rigidbodyintegrator.cu:238-243,276-296,318-321,380-384,433-469, driven
// by platforms/cuda/src/CudaRigidBodyKernels.cpp:118-194,405-438,481-494): third-order finite-difference
// estimates of dr/dt and dq/dt at the end of a step from one VIRTUAL backward step taken at its beginning and
// one virtual forward step taken at its end,
//     rdot = 1/2 (r(-1) - 6 r0 + 3 r1 + 2 r(2)),   qdot likewise, projected orthogonal to q,
//     KE_t = [sum_b rdot.p + sum_free posDot.v m/2]/(6 dt),   KE_r = sum_b qdot.pi/(6 dt),
//     dU   = -(dt^2/24) [sum_b (F.F/M + tau_b.(tau_b/I)) + sum_free f.f/m].
// Kept out of the step kernels on purpose: the diagnostics run as a thread-per-body pass before Part 1 and one
// after Part 2 (two extra rotations per body-step), only when they are switched on.  State here is momentum
// based (p, pi) like the Reference platform, not the CUDA reference's velocities.
// Free atoms: posDot = (-1) (r(-1) - r0) + 5 (r1 - r0) + 2 (r(2) - r1) with the reference's host-side factors -1, 5, 2
// (CudaRigidBodyKernels.cpp:412,429,435).  The -1 is a reference quirk (the stencil needs +1; see oracle/rb_oracle.c) and
// is reproduced; callers that drive the free-atom passes themselves (constraint flow) pass their own factors.
#include "rbk_atomio.cuh"
#include "rbk_device.hpp"
#include "rbk_step.cuh"

namespace rbk {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kThreads = 128;

// virtualRotation (rigidbodyintegrator.cu:238-243): kick by C(q) tau dt, free rotation over dt, return q only
template <bool EXACT>
__device__ d4 virtualRotation(double dt, int nSplit, d4 q, d4 pi, d3 tau, d3 invI) {
    pi = pi + quatC(q, tau)*dt;
    if (EXACT) exactRotation(dt, invI, q, pi);
    else noSquish(dt, nSplit, invI, q, pi);
    return q;
}

// PHASE 1: start of the step, state before the first half kick (rigidbodyintegrator.cu:318-321)
// PHASE 2: end of the step, state after the second half kick   (rigidbodyintegrator.cu:380-384)
template <bool EXACT, int PHASE>
__global__ void __launch_bounds__(kThreads) refinedBodyKernel(const DeviceSystem S, const RefinedState X, const double dt) {
    const int b = blockIdx.x*kThreads + threadIdx.x;
    if (b >= S.numBodies) return;
    const size_t ld = S.bodyStride;
    const double* s = S.state + b;
    const d3 r = loadPlane3(s + PL_R*ld, ld), p = loadPlane3(s + PL_P*ld, ld), F = loadPlane3(s + PL_F*ld, ld);
    const d3 tau = loadPlane3(s + PL_TAU*ld, ld), invI = loadPlane3(s + PL_INVI*ld, ld);
    const d4 q = loadPlane4(s + PL_Q*ld, ld), pi = loadPlane4(s + PL_PI*ld, ld);
    const double invm = s[PL_INVM*ld], halfDt = 0.5*dt;
    const d3 dv = F*(invm*halfDt), v = p*invm;
    double* rd = X.rdot + b;
    double* qd = X.qdot + b;
    if (PHASE == 1) {
        storePlane3(rd, ld, r*2.5 + (v - dv)*halfDt);
        const d4 qv = virtualRotation<EXACT>(-dt, S.rotationMode, q, pi, tau, invI);
        storePlane4(qd, ld, qv*0.5 - q*3.0);
    }
    else {
        storePlane3(rd, ld, r*2.5 + (v + dv)*dt - loadPlane3(rd, ld));
        const d4 qv = virtualRotation<EXACT>(dt, S.rotationMode, q, pi, tau, invI);
        d4 qdot = loadPlane4(qd, ld) + (q*1.5 + qv);
        qdot = qdot - q*dot(qdot, q);
        storePlane4(qd, ld, qdot);
    }
}

// Free atoms without constraints: both virtual displacements are (v + f invm dt'/2) dt'.
template <int PHASE>
__global__ void __launch_bounds__(256) refinedFreeKernel(const DeviceSystem S, const RefinedState X, const double dt,
                                                         const AtomView vel, const AtomView force) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= S.numFree) return;
    const long long gi = atomSlot(S, k);
    const d3 f = loadAtom<false>(force, gi), v = loadAtom<false>(vel, gi);
    const double invm = S.freeInvMass[k];
    double* pd = X.posDot + k;
    if (PHASE == 1) {
        const d3 back = (v + f*invm*(0.5*-dt))*-dt;          // r(-1) - r0
        const d3 fwd = (v + f*invm*(0.5*dt))*dt;             // r1 - r0 (this step's displacement)
        storePlane3(pd, S.freeStride, back*-1.0 + fwd*5.0);
    }
    else storePlane3(pd, S.freeStride, loadPlane3(pd, S.freeStride) + ((v + f*invm*(0.5*dt))*dt)*2.0);   // r(2) - r1
}

// freeAtomsDot (rigidbodyintegrator.cu:291-297) for callers that run the constraint flow themselves
__global__ void __launch_bounds__(256) freeDotKernel(const DeviceSystem S, const RefinedState X, const AtomView delta,
                                                     const double factor, const int restart) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;
    if (k >= S.numFree) return;
    const d3 d = loadAtom<false>(delta, atomSlot(S, k))*factor;
    double* pd = X.posDot + k;
    storePlane3(pd, S.freeStride, restart ? d : loadPlane3(pd, S.freeStride) + d);
}

// Fixed-shape reduction (warp shuffles -> shared -> last CTA), bit-reproducible like kineticKernel.
// MODE 0: refined kinetic energies (rigidbodyintegrator.cu:437-448), out = {KE_t, KE_r}*scale
// MODE 1: potential-energy refinement (rigidbodyintegrator.cu:454-469), out[0] = U*scale
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
    for (int w = 0; w < 256/32; w++) { a += scratch[w][0]; b += scratch[w][1]; }
}

template <int MODE>
__global__ void __launch_bounds__(256) refinedReduceKernel(const DeviceSystem S, const RefinedState X, const AtomView atoms,
                                                           const double scale, double* partial, unsigned* counter, double* out) {
    __shared__ double scratch[256/32][2];
    __shared__ bool isLast;
    const int g = blockIdx.x*256 + threadIdx.x, T = gridDim.x*256;
    const size_t ld = S.bodyStride;
    double a = 0.0, c = 0.0;
    for (int b = g; b < S.numBodies; b += T) {
        const double* s = S.state + b;
        const double invm = s[PL_INVM*ld];
        if (MODE == 0) {
            const d3 v = loadPlane3(s + PL_P*ld, ld)*invm;
            a += dot(loadPlane3(X.rdot + b, ld), v)/invm;
            c += dot(loadPlane4(X.qdot + b, ld), loadPlane4(s + PL_PI*ld, ld));
        }
        else {
            const d4 q = loadPlane4(s + PL_Q*ld, ld);
            const d3 F = loadPlane3(s + PL_F*ld, ld), invI = loadPlane3(s + PL_INVI*ld, ld);
            const d3 tb = quatBt(q, quatC(q, loadPlane3(s + PL_TAU*ld, ld)));        // body-frame torque
            const d3 t2 = {tb.x*invI.x, tb.y*invI.y, tb.z*invI.z};
            a += dot(F, F)*invm + dot(t2, tb);
        }
    }
    for (int k = g; k < S.numFree; k += T) {
        const d3 x = loadAtom<false>(atoms, atomSlot(S, k));                        // MODE 0: velocity, MODE 1: force
        if (MODE == 0) a += dot(loadPlane3(X.posDot + k, S.freeStride), x)*(0.5/S.freeInvMass[k]);
        else a += dot(x, x)*S.freeInvMass[k];
    }
    blockSum2(a, c, scratch);
    if (threadIdx.x == 0) {
        partial[2*blockIdx.x] = a;
        partial[2*blockIdx.x + 1] = c;
        __threadfence();
        isLast = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (isLast) {
        __threadfence();
        a = 0.0; c = 0.0;
        for (int i = threadIdx.x; i < (int) gridDim.x; i += 256) {
            a += __ldcg(partial + 2*i);
            c += __ldcg(partial + 2*i + 1);
        }
        blockSum2(a, c, scratch);
        if (threadIdx.x == 0) {
            out[0] = a*scale;
            out[1] = c*scale;
            *counter = 0u;
        }
    }
}

} // namespace

cudaError_t launchRefinedBodies(const DeviceSystem& S, const RefinedState& X, double dt, int phase, cudaStream_t st) {
    if (S.numBodies == 0) return cudaSuccess;
    const int grid = (S.numBodies + kThreads - 1)/kThreads;
    const bool exact = S.rotationMode == 0;
    if (phase == 1) {
        if (exact) refinedBodyKernel<true, 1><<<grid, kThreads, 0, st>>>(S, X, dt);
        else refinedBodyKernel<false, 1><<<grid, kThreads, 0, st>>>(S, X, dt);
    }
    else {
        if (exact) refinedBodyKernel<true, 2><<<grid, kThreads, 0, st>>>(S, X, dt);
        else refinedBodyKernel<false, 2><<<grid, kThreads, 0, st>>>(S, X, dt);
    }
    return cudaGetLastError();
}

cudaError_t launchRefinedFree(const DeviceSystem& S, const RefinedState& X, double dt, int phase, AtomView vel, AtomView force,
                              cudaStream_t st) {
    if (S.numFree == 0) return cudaSuccess;
    const int grid = (S.numFree + 255)/256;
    if (phase == 1) refinedFreeKernel<1><<<grid, 256, 0, st>>>(S, X, dt, vel, force);
    else refinedFreeKernel<2><<<grid, 256, 0, st>>>(S, X, dt, vel, force);
    return cudaGetLastError();
}

cudaError_t launchFreeDot(const DeviceSystem& S, const RefinedState& X, AtomView delta, double factor, bool restart, cudaStream_t st) {
    if (S.numFree == 0) return cudaSuccess;
    freeDotKernel<<<(S.numFree + 255)/256, 256, 0, st>>>(S, X, delta, factor, restart ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launchRefinedKinetic(const DeviceSystem& S, const RefinedState& X, double dt, AtomView vel, double* partial,
                                 unsigned* counter, double* out, cudaStream_t st) {
    refinedReduceKernel<0><<<kKineticBlocks, 256, 0, st>>>(S, X, vel, 1.0/(6.0*dt), partial, counter, out);
    return cudaGetLastError();
}

cudaError_t launchPotentialRefinement(const DeviceSystem& S, const RefinedState& X, double dt, AtomView force, double* partial,
                                      unsigned* counter, double* out, cudaStream_t st) {
    refinedReduceKernel<1><<<kKineticBlocks, 256, 0, st>>>(S, X, force, -dt*dt/24.0, partial, counter, out);
    return cudaGetLastError();
}

} // namespace rbk
