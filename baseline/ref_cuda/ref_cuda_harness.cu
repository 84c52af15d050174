// baseline/ref_cuda: the reference's OWN CUDA kernels (platforms/cuda/src/kernels/{vectorOps,elliptic,rigidbodyintegrator}.cu),
// compiled IN PLACE from /root/reference for sm_100a and driven on the benchmark workload - bench / test infrastructure
// only ("the design to beat").  Nothing in the product links or loads this.
//
// OpenMM compiles these files at run time (CudaContext::createModule) after injecting type names and a few defines
// (platforms/cuda/src/CudaRigidBodyKernels.cpp:238-247); this shim injects the same names:
//   real / real4      posq element type      float (single, mixed) or double
//   mixed / mixed3/4  integrator precision   double (mixed, double) or float (single)
//   ROTATION, NSPLIT, COMPMOD                rotation routine, NO-SQUISH sub-steps, refined-energy code on/off
// RB_PRECISION: 0 = single, 1 = mixed, 2 = double.  Launch geometry as CudaContext::executeKernel: blocks of 128 threads,
// grid = min(ceil(work/128), numThreadBlocks) with numThreadBlocks = 4 x 15 x SMs / ... (OpenMM 7.x: 6 blocks per SM);
// the kernels are grid-stride loops, so the driver below also tries larger grids and reports the best time.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef RB_PRECISION
#define RB_PRECISION 1
#endif
#if RB_PRECISION == 0
typedef float real; typedef float2 real2; typedef float3 real3; typedef float4 real4;
typedef float mixed; typedef float2 mixed2; typedef float3 mixed3; typedef float4 mixed4;
#define make_mixed2 make_float2
#define make_mixed3 make_float3
#define make_mixed4 make_float4
#elif RB_PRECISION == 1
#define USE_MIXED_PRECISION 1
typedef float real; typedef float2 real2; typedef float3 real3; typedef float4 real4;
typedef double mixed; typedef double2 mixed2; typedef double3 mixed3; typedef double4 mixed4;
#define make_mixed2 make_double2
#define make_mixed3 make_double3
#define make_mixed4 make_double4
#else
// OpenMM's CudaContext defines USE_DOUBLE_PRECISION *instead of* USE_MIXED_PRECISION in double-precision mode, so the
// reference's elliptic.cu (which tests USE_MIXED_PRECISION) runs its Carlson / Jacobi routines with the single-precision
// tolerances (errtol 0.03, FLT_EPSILON) on doubles there - reproduced as it is.
#define USE_DOUBLE_PRECISION 1
typedef double real; typedef double2 real2; typedef double3 real3; typedef double4 real4;
typedef double mixed; typedef double2 mixed2; typedef double3 mixed3; typedef double4 mixed4;
#define make_mixed2 make_double2
#define make_mixed3 make_double3
#define make_mixed4 make_double4
#endif
#ifndef RB_MODE
#define RB_MODE 0
#endif
#if RB_MODE == 0
#define ROTATION exactRotation
#define NSPLIT 1
#else
#define ROTATION noSquishRotation
#define NSPLIT RB_MODE
#endif
#ifndef COMPMOD
#define COMPMOD 0
#endif

// (global scope: the files overload make_float3 & co. next to CUDA's own and declare their kernels extern "C")
#include "platforms/cuda/src/kernels/vectorOps.cu"
#include "platforms/cuda/src/kernels/elliptic.cu"
#include "platforms/cuda/src/kernels/rigidbodyintegrator.cu"

// ---- driver --------------------------------------------------------------------------------------------------------
// Host side of CudaIntegrateRigidBodyStepKernel reduced to what the kernels need (platforms/cuda/src/
// CudaRigidBodyKernels.cpp:196-230 allocateArrays, :293-372 uploadBodySystem, :377-444 execute, :118-194 energies):
// the caller hands over body data as plain double arrays and the OpenMM-format device arrays; no OpenMM involved.
struct RefCuda {
    int numFree, numBodies, numBodyAtoms, padded, numSMs;
    BodyData* bodyData;
    int* atomLocation;
    mixed3 *bodyFixedPos, *savedPos, *posDot;
    mixed4* posDelta;
    mixed *atomE, *bodyE1, *bodyE2;
};

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "ref_cuda: %s: %s\n", #x, cudaGetErrorString(e_)); return -1; } } while (0)

extern "C" {

int refcuda_precision() { return RB_PRECISION; }
int refcuda_mode() { return RB_MODE; }
int refcuda_compmod() { return COMPMOD; }
int refcuda_sizeof_body() { return (int) sizeof(BodyData); }

void* refcuda_create(int numFree, int numBodies, int numBodyAtoms, int paddedNumAtoms) {
    RefCuda* h = new RefCuda();
    std::memset(h, 0, sizeof(RefCuda));
    h->numFree = numFree; h->numBodies = numBodies; h->numBodyAtoms = numBodyAtoms; h->padded = paddedNumAtoms;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&h->numSMs, cudaDevAttrMultiProcessorCount, dev);
    const size_t nb = numBodies > 0 ? numBodies : 1, nf = numFree > 0 ? numFree : 1, na = numBodyAtoms > 0 ? numBodyAtoms : 1;
    if (cudaMalloc(&h->bodyData, nb*sizeof(BodyData)) != cudaSuccess || cudaMalloc(&h->atomLocation, (nf + na)*sizeof(int)) != cudaSuccess ||
        cudaMalloc(&h->bodyFixedPos, na*sizeof(mixed3)) != cudaSuccess || cudaMalloc(&h->savedPos, nf*sizeof(mixed3)) != cudaSuccess ||
        cudaMalloc(&h->posDot, nf*sizeof(mixed3)) != cudaSuccess || cudaMalloc(&h->posDelta, (size_t) paddedNumAtoms*sizeof(mixed4)) != cudaSuccess ||
        cudaMalloc(&h->atomE, nf*sizeof(mixed)) != cudaSuccess || cudaMalloc(&h->bodyE1, nb*sizeof(mixed)) != cudaSuccess ||
        cudaMalloc(&h->bodyE2, nb*sizeof(mixed)) != cudaSuccess) {
        std::fprintf(stderr, "ref_cuda: allocation failed\n");
        return NULL;
    }
    cudaMemset(h->posDelta, 0, (size_t) paddedNumAtoms*sizeof(mixed4));
    cudaMemset(h->posDot, 0, nf*sizeof(mixed3));
    cudaMemset(h->savedPos, 0, nf*sizeof(mixed3));
    return h;
}

void refcuda_destroy(void* p) {
    RefCuda* h = (RefCuda*) p;
    if (!h) return;
    cudaFree(h->bodyData); cudaFree(h->atomLocation); cudaFree(h->bodyFixedPos); cudaFree(h->savedPos); cudaFree(h->posDot);
    cudaFree(h->posDelta); cudaFree(h->atomE); cudaFree(h->bodyE1); cudaFree(h->bodyE2);
    delete h;
}

// uploadBodySystem (CudaRigidBodyKernels.cpp:293-372): per body N, loc, invMass, invI, rcm, v = pcm/mass, force, q, pi,
// torque (the 4-vector C(q) tau); rdot = qdot = 0; body-frame coordinates per body atom; atomLocation per actual atom
int refcuda_upload(void* p, const int* N, const int* loc, const double* mass, const double* invI, const double* rcm, const double* pcm,
                   const double* force, const double* q, const double* pi, const double* torque, const double* bodyFixed,
                   const int* atomLocation) {
    RefCuda* h = (RefCuda*) p;
    std::vector<BodyData> data(h->numBodies > 0 ? h->numBodies : 1);
    for (int i = 0; i < h->numBodies; i++) {
        BodyData& b = data[i];
        b.N = N[i]; b.loc = loc[i];
        b.invm = (mixed) (1.0/mass[i]);
        b.invI = make_mixed3((mixed) invI[3*i], (mixed) invI[3*i+1], (mixed) invI[3*i+2]);
        b.r = make_mixed3((mixed) rcm[3*i], (mixed) rcm[3*i+1], (mixed) rcm[3*i+2]);
        b.v = make_mixed3((mixed) (pcm[3*i]/mass[i]), (mixed) (pcm[3*i+1]/mass[i]), (mixed) (pcm[3*i+2]/mass[i]));
        b.F = make_mixed3((mixed) force[3*i], (mixed) force[3*i+1], (mixed) force[3*i+2]);
        b.rdot = make_mixed3(0, 0, 0);
        b.q = make_mixed4((mixed) q[4*i], (mixed) q[4*i+1], (mixed) q[4*i+2], (mixed) q[4*i+3]);
        b.pi = make_mixed4((mixed) pi[4*i], (mixed) pi[4*i+1], (mixed) pi[4*i+2], (mixed) pi[4*i+3]);
        b.Ctau = make_mixed4((mixed) torque[4*i], (mixed) torque[4*i+1], (mixed) torque[4*i+2], (mixed) torque[4*i+3]);
        b.qdot = make_mixed4(0, 0, 0, 0);
    }
    CK(cudaMemcpy(h->bodyData, data.data(), (size_t) h->numBodies*sizeof(BodyData), cudaMemcpyHostToDevice));
    std::vector<mixed3> d(h->numBodyAtoms > 0 ? h->numBodyAtoms : 1);
    for (int i = 0; i < h->numBodyAtoms; i++) d[i] = make_mixed3((mixed) bodyFixed[3*i], (mixed) bodyFixed[3*i+1], (mixed) bodyFixed[3*i+2]);
    CK(cudaMemcpy(h->bodyFixedPos, d.data(), (size_t) h->numBodyAtoms*sizeof(mixed3), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->atomLocation, atomLocation, (size_t) (h->numFree + h->numBodyAtoms)*sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemset(h->posDot, 0, (size_t) (h->numFree > 0 ? h->numFree : 1)*sizeof(mixed3)));
    return 0;
}

static int gridFor(const RefCuda* h, int work, int blocksPerSM) {        // CudaContext::executeKernel
    const int want = (work + 127)/128, cap = h->numSMs*blocksPerSM;
    return want < cap ? (want > 0 ? want : 1) : cap;
}

// which: 0 = freeAtomsDelta, 1 = freeAtomsDot, 2 = integrateRigidBodyPart1, 3 = integrateRigidBodyPart2
int refcuda_launch(void* p, int which, double dt, void* posq, void* posqCorrection, void* velm, const long long* force, int blocksPerSM,
                   int posDotRestart, int posDotFactor) {
    RefCuda* h = (RefCuda*) p;
    const int work = which < 2 ? h->numFree : (h->numFree > h->numBodies ? h->numFree : h->numBodies);
    if (work == 0) return 0;
    const int grid = gridFor(h, work, blocksPerSM);
#define RB_ARGS h->padded, h->numFree, h->numBodies, (mixed) dt, (real4*) posq, (real4*) posqCorrection, (mixed4*) velm, force, h->posDelta, \
                h->bodyData, h->atomLocation, h->bodyFixedPos, h->savedPos, posDotRestart, posDotFactor, h->posDot
    if (which == 0) freeAtomsDelta<<<grid, 128>>>(RB_ARGS);
    else if (which == 1) freeAtomsDot<<<grid, 128>>>(RB_ARGS);
    else if (which == 2) integrateRigidBodyPart1<<<grid, 128>>>(RB_ARGS);
    else integrateRigidBodyPart2<<<grid, 128>>>(RB_ARGS);
    CK(cudaGetLastError());
    return 0;
}

// `steps` x [Part 1 with forceA/forceB alternating as "old" forces, Part 2 with the other one], CUDA-event timed.
int refcuda_time_steps(void* p, double dt, int steps, void* posq, void* posqCorrection, void* velm, const long long* forceA,
                       const long long* forceB, int blocksPerSM, int startWith, float* msTotal) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int cur = startWith;
    CK(cudaEventRecord(e0, 0));
    for (int i = 0; i < steps; i++) {
        if (refcuda_launch(p, 2, dt, posq, posqCorrection, velm, cur ? forceB : forceA, blocksPerSM, 0, 0)) return -1;
        cur ^= 1;
        if (refcuda_launch(p, 3, dt, posq, posqCorrection, velm, cur ? forceB : forceA, blocksPerSM, 0, 0)) return -1;
    }
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(msTotal, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return cur;
}

// kineticEnergy<>() (CudaRigidBodyKernels.cpp:118-163): kernel + host-side sums; refined: factor 1/(6 dt) applied by the caller
int refcuda_kinetic(void* p, void* velm, int refined, int blocksPerSM, double* out) {
    RefCuda* h = (RefCuda*) p;
    const int work = h->numFree > h->numBodies ? h->numFree : h->numBodies;
    const int grid = gridFor(h, work > 0 ? work : 1, blocksPerSM);
    if (refined) refinedKineticEnergies<<<grid, 128>>>(h->numFree, h->numBodies, (mixed4*) velm, h->bodyData, h->atomLocation, h->posDot, h->atomE, h->bodyE1, h->bodyE2);
    else kineticEnergies<<<grid, 128>>>(h->numFree, h->numBodies, (mixed4*) velm, h->bodyData, h->atomLocation, h->posDot, h->atomE, h->bodyE1, h->bodyE2);
    CK(cudaGetLastError());
    std::vector<mixed> e((size_t) (work > 0 ? work : 1));
    mixed kt = 0, kr = 0;
    if (h->numFree) { CK(cudaMemcpy(e.data(), h->atomE, h->numFree*sizeof(mixed), cudaMemcpyDeviceToHost)); for (int i = 0; i < h->numFree; i++) kt += e[i]; }
    if (h->numBodies) {
        CK(cudaMemcpy(e.data(), h->bodyE1, h->numBodies*sizeof(mixed), cudaMemcpyDeviceToHost)); for (int i = 0; i < h->numBodies; i++) kt += e[i];
        CK(cudaMemcpy(e.data(), h->bodyE2, h->numBodies*sizeof(mixed), cudaMemcpyDeviceToHost)); for (int i = 0; i < h->numBodies; i++) kr += e[i];
    }
    out[0] = kt; out[1] = kr;
    return 0;
}

// potentialEnergyRefinement<>() (CudaRigidBodyKernels.cpp:168-194): the raw sum; the caller applies -dt^2/24
int refcuda_potential_refinement(void* p, void* velm, const long long* force, int blocksPerSM, double* out) {
    RefCuda* h = (RefCuda*) p;
    const int work = h->numFree > h->numBodies ? h->numFree : h->numBodies;
    potentialEnergyRefinement<<<gridFor(h, work > 0 ? work : 1, blocksPerSM), 128>>>(h->padded, h->numFree, h->numBodies, (mixed4*) velm, force, h->bodyData,
                                                                                 h->atomLocation, h->atomE, h->bodyE1);
    CK(cudaGetLastError());
    std::vector<mixed> e((size_t) (work > 0 ? work : 1));
    mixed u = 0;
    if (h->numFree) { CK(cudaMemcpy(e.data(), h->atomE, h->numFree*sizeof(mixed), cudaMemcpyDeviceToHost)); for (int i = 0; i < h->numFree; i++) u += e[i]; }
    if (h->numBodies) { CK(cudaMemcpy(e.data(), h->bodyE1, h->numBodies*sizeof(mixed), cudaMemcpyDeviceToHost)); for (int i = 0; i < h->numBodies; i++) u += e[i]; }
    out[0] = u;
    return 0;
}

// body state back to the host: r[3] v[3] q[4] pi[4] F[3] Ctau[4] per body (doubles)
int refcuda_download(void* p, double* r, double* v, double* q, double* pi, double* F, double* Ctau) {
    RefCuda* h = (RefCuda*) p;
    std::vector<BodyData> data(h->numBodies > 0 ? h->numBodies : 1);
    CK(cudaMemcpy(data.data(), h->bodyData, (size_t) h->numBodies*sizeof(BodyData), cudaMemcpyDeviceToHost));
    for (int i = 0; i < h->numBodies; i++) {
        const BodyData& b = data[i];
        if (r) { r[3*i] = b.r.x; r[3*i+1] = b.r.y; r[3*i+2] = b.r.z; }
        if (v) { v[3*i] = b.v.x; v[3*i+1] = b.v.y; v[3*i+2] = b.v.z; }
        if (F) { F[3*i] = b.F.x; F[3*i+1] = b.F.y; F[3*i+2] = b.F.z; }
        if (q) { q[4*i] = b.q.x; q[4*i+1] = b.q.y; q[4*i+2] = b.q.z; q[4*i+3] = b.q.w; }
        if (pi) { pi[4*i] = b.pi.x; pi[4*i+1] = b.pi.y; pi[4*i+2] = b.pi.z; pi[4*i+3] = b.pi.w; }
        if (Ctau) { Ctau[4*i] = b.Ctau.x; Ctau[4*i+1] = b.Ctau.y; Ctau[4*i+2] = b.Ctau.z; Ctau[4*i+3] = b.Ctau.w; }
    }
    return 0;
}

} // extern "C"
