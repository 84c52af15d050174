// rbk_build.cu - GPU-side body build: what RigidBodySystem::update does on the host in the reference
// (openmmapi/src/RigidBodySystem.cpp:120-142 -> RigidBody::buildGeometry / buildDynamics, RigidBody.cpp:26-142,
// eigenDecomposition.cpp:36-176, MatVec.cpp:344-373), one thread per body, writing straight into the SoA state
// planes and the body-frame coordinate planes.  At 1M bodies the host rebuild + upload on every
// setPositions/setVelocities costs seconds; this costs well under a millisecond.
//
// Compiled with -fmad=false so that every operation rounds exactly like the host model (csrc/rbk_host.cpp, which
// is bit-identical to the reference); only acos/cos come from a different libm, so moments, quaternions and
// body-frame coordinates agree with the host build to ~1e-15 relative (tests/test_gpu_build.py).
//
// Semantics note: the dynamics build SETS p = sum m v.  The reference accumulates into the host copy's stale pcm
// (RigidBody.cpp:126-130), which only differs when velocities are set more than once per setPositions - a flow in
// which the reference re-uploads stale positions as well (RigidBodyIntegrator.cpp:63-74).
#include "rbk_device.hpp"
#include "rbk_math.cuh"

#include <cfloat>

namespace rbk {
namespace {

constexpr double kPiRef = 3.14159265358979323846264338328;

struct m3 { d3 r[3]; };

__device__ __forceinline__ double comp(d3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
__device__ __forceinline__ void setComp(d3& v, int i, double s) { if (i == 0) v.x = s; else if (i == 1) v.y = s; else v.z = s; }
__device__ __forceinline__ d3 divide(d3 a, double s) { const double k = 1.0/s; return a*k; }

__device__ d3 loadAny(const AtomView& A, long long i) {
    if (A.fmt == FMT_F64) { const double* p = A.p + i*A.sa; return {p[0], p[A.sc], p[2*A.sc]}; }
    if (A.fmt == FMT_POSQ_MIXED) {
        const float4 hi = reinterpret_cast<const float4*>(A.p)[i], lo = reinterpret_cast<const float4*>(A.aux)[i];
        return {(double) hi.x + (double) lo.x, (double) hi.y + (double) lo.y, (double) hi.z + (double) lo.z};
    }
    if (A.fmt == FMT_REAL4_F64) { const double* p = A.p + 4*i; return {p[0], p[1], p[2]}; }
    if (A.fmt == FMT_REAL4_F32) { const float4 v = reinterpret_cast<const float4*>(A.p)[i]; return {(double) v.x, (double) v.y, (double) v.z}; }
    const long long* f = reinterpret_cast<const long long*>(A.p) + i;
    return {(double) f[0]/4294967296.0, (double) f[A.sc]/4294967296.0, (double) f[2*A.sc]/4294967296.0};
}

__device__ __forceinline__ long long slotOf(const DeviceSystem& S, int pluginIndex) {
    return S.atomLoc ? (long long) S.atomLoc[pluginIndex] : (long long) pluginIndex;
}

// eigenvalues of a symmetric matrix, descending (eigenDecomposition.cpp:72-103)
__device__ d3 principalMoments(const m3& A) {
    const double offsq = A.r[0].y*A.r[0].y + A.r[0].z*A.r[0].z + A.r[1].z*A.r[1].z;
    d3 w = {A.r[0].x, A.r[1].y, A.r[2].z};
    if (offsq < DBL_EPSILON) {
        double t;
        if (w.x < w.y) { t = w.x; w.x = w.y; w.y = t; }
        if (w.x < w.z) { t = w.x; w.x = w.z; w.z = t; }
        if (w.y < w.z) { t = w.y; w.y = w.z; w.z = t; }
        return w;
    }
    const double trace = w.x + w.y + w.z;
    const double mean = trace/3.0;
    const d3 dev = {w.x - mean, w.y - mean, w.z - mean};
    const double p2 = dot(dev, dev) + 2.0*offsq;
    const double p = sqrt(p2/6.0);
    const d3 a0 = {A.r[0].x - mean, A.r[0].y, A.r[0].z};
    const d3 a1 = {A.r[1].x, A.r[1].y - mean, A.r[1].z};
    const d3 a2 = {A.r[2].x, A.r[2].y, A.r[2].z - mean};
    const double det = a0.x*(a1.y*a2.z - a2.y*a1.z) - a0.y*(a1.x*a2.z - a2.x*a1.z) + a0.z*(a1.x*a2.y - a2.x*a1.y);
    const double r = det*(3.0/(p*p2));
    const double phi = r <= -1.0 ? kPiRef/3.0 : (r >= 1.0 ? 0.0 : acos(r)/3.0);
    const double hi = mean + 2.0*p*cos(phi);
    const double lo = mean + 2.0*p*cos(phi + 2.0*kPiRef/3.0);
    return {hi, trace - (hi + lo), lo};
}

// eigenDecomposition.cpp:36-68
__device__ void finishAxis(d3& v, const m3& a, double n1base, double n2base, double thresh) {
    double norm = dot(v, v);
    const double n1 = n1base + a.r[0].x*a.r[0].x;
    const double n2 = n2base + a.r[1].y*a.r[1].y;
    const double error = n1*n2;
    if (n1 <= thresh) v = {1.0, 0.0, 0.0};
    else if (n2 <= thresh) v = {0.0, 1.0, 0.0};
    else if (norm < 4096.0*DBL_EPSILON*DBL_EPSILON*error) {
        double t = fabs(a.r[0].y);
        double f = -a.r[0].x/a.r[0].y;
        if (fabs(a.r[1].y) > t) { t = fabs(a.r[1].y); f = -a.r[0].y/a.r[1].y; }
        if (fabs(a.r[1].z) > t) f = -a.r[0].z/a.r[1].z;
        norm = 1.0/sqrt(1.0 + f*f);
        v = {norm, f*norm, 0.0};
    }
    else v = v*sqrt(1.0/norm);
}

// eigenvectors as rows (eigenDecomposition.cpp:107-176)
__device__ m3 principalAxes(const m3& A, d3 w) {
    const double tiny = 8.0*DBL_EPSILON*fabs(w.x);
    const double thresh = tiny*tiny;
    m3 a = A;
    a.r[1].x = a.r[0].y; a.r[2].x = a.r[0].z; a.r[2].y = a.r[1].z;
    double n1 = a.r[0].y*a.r[0].y + a.r[0].z*a.r[0].z;
    const double n2 = a.r[0].y*a.r[0].y + a.r[1].z*a.r[1].z;
    const double c0 = a.r[0].y*a.r[1].z - a.r[0].z*a.r[1].y;
    const double c1 = a.r[0].z*a.r[0].y - a.r[1].z*a.r[0].x;
    const double c2 = a.r[0].y*a.r[0].y;
    a.r[0].x -= w.x;
    a.r[1].y -= w.x;
    d3 e0 = {c0 + a.r[0].z*w.x, c1 + a.r[1].z*w.x, a.r[0].x*a.r[1].y - c2};
    finishAxis(e0, a, n1, n2, thresh);
    d3 e1 = {c0, c1, c2};
    const double gap = w.x - w.y;
    if (fabs(gap) > tiny) {
        a.r[0].x += gap;
        a.r[1].y += gap;
        e1 = {c0 + a.r[0].z*w.y, c1 + a.r[1].z*w.y, a.r[0].x*a.r[1].y - c2};
        finishAxis(e1, a, n1, n2, thresh);
    }
    else {
        a.r[0].x += w.x;
        a.r[1].y += w.x;
        bool ok = false;
        for (int i = 0; i < 3 && !ok; i++) {
            if (i == 0) a.r[0].x -= w.y; else if (i == 1) a.r[1].y -= w.y; else a.r[2].z -= w.y;
            const d3 col = {comp(a.r[0], i), comp(a.r[1], i), comp(a.r[2], i)};
            n1 = dot(col, col);
            ok = n1 > thresh;
            if (ok) {
                e1 = cross(e0, col);
                const double norm = dot(e1, e1);
                ok = norm > 65536.0*DBL_EPSILON*DBL_EPSILON*n1;
                if (ok) e1 = e1*sqrt(1.0/norm);
            }
        }
        if (!ok) {
            int i = 0;
            while (i < 2 && comp(e0, i) == 0.0) i++;
            const double norm = 1.0/sqrt(comp(e0, i)*comp(e0, i) + comp(e0, i)*comp(e0, i));
            const double ei = comp(e0, i);
            setComp(e1, i, ei*norm);
            setComp(e1, i, -ei*norm);
            if (i + 1 < 3) setComp(e1, i + 1, 0.0);
        }
    }
    m3 out;
    out.r[0] = e0; out.r[1] = e1; out.r[2] = cross(e0, e1);
    return out;
}

// Shepperd's rotation matrix -> quaternion (MatVec.cpp:344-373)
__device__ d4 quaternionOf(const m3& A) {
    const double t0 = A.r[0].x, t1 = A.r[1].y, t2 = A.r[2].z;
    const double cand[4] = {1.0 + t0 + t1 + t2, 1.0 + t0 - t1 - t2, 1.0 - t0 + t1 - t2, 1.0 - t0 - t1 + t2};
    int k = 0;
    for (int i = 1; i < 4; i++) if (cand[i] > cand[k]) k = i;
    const double big = cand[k];
    const double f = 0.5/sqrt(big);
    double q[4] = {0.0, 0.0, 0.0, 0.0};
    if (k == 0) { q[1] = (A.r[1].z - A.r[2].y)*f; q[2] = (A.r[2].x - A.r[0].z)*f; q[3] = (A.r[0].y - A.r[1].x)*f; }
    else if (k == 1) { q[0] = (A.r[1].z - A.r[2].y)*f; q[2] = (A.r[0].y + A.r[1].x)*f; q[3] = (A.r[0].z + A.r[2].x)*f; }
    else if (k == 2) { q[0] = (A.r[2].x - A.r[0].z)*f; q[1] = (A.r[0].y + A.r[1].x)*f; q[3] = (A.r[1].z + A.r[2].y)*f; }
    else { q[0] = (A.r[0].y - A.r[1].x)*f; q[1] = (A.r[0].z + A.r[2].x)*f; q[2] = (A.r[1].z + A.r[2].y)*f; }
    q[k] = big*f;
    return {q[0], q[1], q[2], q[3]};
}

__device__ m3 complementProjector(d3 u) {             // (u.u) 1 - u u^T  (MatVec.cpp:555-564)
    m3 P;
    const double uu = dot(u, u);
    P.r[0] = {(-u.x)*u.x, (-u.y)*u.x, (-u.z)*u.x};
    P.r[1] = {(-u.x)*u.y, (-u.y)*u.y, (-u.z)*u.y};
    P.r[2] = {(-u.x)*u.z, (-u.y)*u.z, (-u.z)*u.z};
    P.r[0].x += uu; P.r[1].y += uu; P.r[2].z += uu;
    return P;
}

__global__ void __launch_bounds__(128) buildGeometryKernel(const DeviceSystem S, const double* __restrict__ atomMass,
                                                           const AtomView pos, const AtomView force, double* dxyz, int* dofSum) {
    const int b = blockIdx.x*blockDim.x + threadIdx.x;
    if (b >= S.numBodies) return;
    const int a0 = S.loc[b], N = S.loc[b + 1] - a0;
    const size_t ld = S.bodyStride, as = S.atomStride;
    double mass = 0.0;
    d3 com = {0.0, 0.0, 0.0};
    for (int j = 0; j < N; j++) {
        const double m = atomMass[a0 + j];
        mass += m;
        com = com + loadAny(pos, slotOf(S, S.numFree + a0 + j))*m;
    }
    com = divide(com, mass);
    // collinearity test (RigidBody.cpp:26-45): the reference keeps the LAST atom farther than atom 0
    const d3 x0 = loadAny(pos, slotOf(S, S.numFree + a0)) - com;
    const double first = dot(x0, x0);
    double longest = first;
    int jmax = 0;
    for (int j = 1; j < N; j++) {
        const d3 x = loadAny(pos, slotOf(S, S.numFree + a0 + j)) - com;
        const double d2 = dot(x, x);
        if (d2 > first) { jmax = j; longest = d2; }
    }
    const d3 axis = divide(loadAny(pos, slotOf(S, S.numFree + a0 + jmax)) - com, sqrt(longest));
    bool line = true;
    for (int j = 0; line && j < N; j++) {
        const d3 x = loadAny(pos, slotOf(S, S.numFree + a0 + j)) - com;
        const double d2 = dot(x, x), proj = dot(axis, x);
        line = line && (d2 < 1.0E-5*longest || fabs(proj*proj/d2 - 1.0) < 1.0E-5);
    }
    d3 I, invI;
    m3 A;
    int dof;
    if (line) {
        double moment = 0.0;
        for (int j = 0; j < N; j++) {
            const d3 x = loadAny(pos, slotOf(S, S.numFree + a0 + j)) - com;
            moment += atomMass[a0 + j]*dot(x, x);
        }
        I = {moment, moment, 0.0};
        invI = {1.0/moment, 1.0/moment, 0.0};
        int imin = axis.x < axis.y ? 0 : 1;
        if (axis.z < comp(axis, imin)) imin = 2;
        d3 e = {0.0, 0.0, 0.0};
        setComp(e, imin, 1.0);
        const m3 P = complementProjector(axis);
        d3 v = {dot(P.r[0], e), dot(P.r[1], e), dot(P.r[2], e)};
        v = divide(v, sqrt(dot(v, v)));
        const d3 w = cross(axis, v);
        A.r[0] = {v.x, w.x, axis.x}; A.r[1] = {v.y, w.y, axis.y}; A.r[2] = {v.z, w.z, axis.z};
        dof = 5;
    }
    else {
        m3 inertia;
        inertia.r[0] = inertia.r[1] = inertia.r[2] = {0.0, 0.0, 0.0};
        for (int j = 0; j < N; j++) {
            const m3 P = complementProjector(loadAny(pos, slotOf(S, S.numFree + a0 + j)) - com);
            const double m = atomMass[a0 + j];
            for (int r = 0; r < 3; r++) inertia.r[r] = inertia.r[r] + P.r[r]*m;
        }
        I = principalMoments(inertia);
        invI = {1.0/I.x, 1.0/I.y, 1.0/I.z};
        A = principalAxes(inertia, I);
        dof = 6;
    }
    const d4 q = quaternionOf(A);
    d3 F = {0.0, 0.0, 0.0}, tau = {0.0, 0.0, 0.0};
    for (int j = 0; j < N; j++) {
        const long long slot = slotOf(S, S.numFree + a0 + j);
        const d3 x = loadAny(pos, slot) - com;
        dxyz[a0 + j] = dot(A.r[0], x);
        dxyz[a0 + j + as] = dot(A.r[1], x);
        dxyz[a0 + j + 2*as] = dot(A.r[2], x);
        const d3 f = loadAny(force, slot);
        F = F + f;
        tau = tau + cross(x, f);
    }
    double* s = S.state + b;
    s[(PL_R + 0)*ld] = com.x; s[(PL_R + 1)*ld] = com.y; s[(PL_R + 2)*ld] = com.z;
    s[(PL_Q + 0)*ld] = q.w; s[(PL_Q + 1)*ld] = q.x; s[(PL_Q + 2)*ld] = q.y; s[(PL_Q + 3)*ld] = q.z;
    s[(PL_F + 0)*ld] = F.x; s[(PL_F + 1)*ld] = F.y; s[(PL_F + 2)*ld] = F.z;
    s[(PL_TAU + 0)*ld] = tau.x; s[(PL_TAU + 1)*ld] = tau.y; s[(PL_TAU + 2)*ld] = tau.z;
    s[PL_INVM*ld] = 1.0/mass;
    s[(PL_I + 0)*ld] = I.x; s[(PL_I + 1)*ld] = I.y; s[(PL_I + 2)*ld] = I.z;
    s[(PL_INVI + 0)*ld] = invI.x; s[(PL_INVI + 1)*ld] = invI.y; s[(PL_INVI + 2)*ld] = invI.z;
    atomicAdd(dofSum, dof);                               // integer sum: order-independent
}

__global__ void __launch_bounds__(128) buildDynamicsKernel(const DeviceSystem S, const double* __restrict__ atomMass,
                                                           const AtomView vel) {
    const int b = blockIdx.x*blockDim.x + threadIdx.x;
    if (b >= S.numBodies) return;
    const int a0 = S.loc[b], N = S.loc[b + 1] - a0;
    const size_t ld = S.bodyStride, as = S.atomStride;
    double* s = S.state + b;
    const d4 q = {s[(PL_Q + 0)*ld], s[(PL_Q + 1)*ld], s[(PL_Q + 2)*ld], s[(PL_Q + 3)*ld]};
    const d3 invI = {s[(PL_INVI + 0)*ld], s[(PL_INVI + 1)*ld], s[(PL_INVI + 2)*ld]};
    double mass = 0.0;
    d3 p = {0.0, 0.0, 0.0};
    for (int j = 0; j < N; j++) {
        const double m = atomMass[a0 + j];
        mass += m;
        p = p + loadAny(vel, slotOf(S, S.numFree + a0 + j))*m;
    }
    const d3 vcm = divide(p, mass);
    d3 L = {0.0, 0.0, 0.0};
    for (int j = 0; j < N; j++) {
        const d3 rel = loadAny(vel, slotOf(S, S.numFree + a0 + j)) - vcm;
        const d3 inBody = quatBt(q, quatC(q, rel))*atomMass[a0 + j];            // A(q) (v - vcm) m
        const d3 dj = {S.dxyz[a0 + j], S.dxyz[a0 + j + as], S.dxyz[a0 + j + 2*as]};
        L = L + cross(dj, inBody);
    }
    const d4 pi = quatB(q, L)*2.0;
    (void) invI;
    s[(PL_P + 0)*ld] = p.x; s[(PL_P + 1)*ld] = p.y; s[(PL_P + 2)*ld] = p.z;
    s[(PL_PI + 0)*ld] = pi.w; s[(PL_PI + 1)*ld] = pi.x; s[(PL_PI + 2)*ld] = pi.y; s[(PL_PI + 3)*ld] = pi.z;
}

} // namespace

cudaError_t launchBuild(const DeviceSystem& S, const double* atomMass, AtomView pos, AtomView vel, AtomView force,
                        double* dxyz, bool geometry, bool velocities, int* dofSum, cudaStream_t st) {
    if (S.numBodies == 0) return cudaSuccess;
    const int grid = (S.numBodies + 127)/128;
    if (geometry) {
        cudaError_t e = cudaMemsetAsync(dofSum, 0, sizeof(int), st);
        if (e != cudaSuccess) return e;
        buildGeometryKernel<<<grid, 128, 0, st>>>(S, atomMass, pos, force, dxyz, dofSum);
    }
    if (velocities) buildDynamicsKernel<<<grid, 128, 0, st>>>(S, atomMass, vel);
    return cudaGetLastError();
}

} // namespace rbk
