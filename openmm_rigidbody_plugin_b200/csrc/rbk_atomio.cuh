// rbk_atomio.cuh - device helpers shared by the kernel translation units: SoA plane access and the caller's
// atom-array formats (plain fp64 and the OpenMM-CUDA boundary formats, see rbk_device.hpp).
#pragma once
#include "rbk_device.hpp"
#include "rbk_math.cuh"

namespace rbk {
namespace {

__device__ __forceinline__ d3 loadPlane3(const double* base, size_t stride) {
    return {base[0], base[stride], base[2*stride]};
}
__device__ __forceinline__ d4 loadPlane4(const double* base, size_t stride) {
    return {base[0], base[stride], base[2*stride], base[3*stride]};
}
__device__ __forceinline__ void storePlane3(double* base, size_t stride, d3 v) {
    base[0] = v.x; base[stride] = v.y; base[2*stride] = v.z;
}
__device__ __forceinline__ void storePlane4(double* base, size_t stride, d4 v) {
    base[0] = v.w; base[stride] = v.x; base[2*stride] = v.y; base[3*stride] = v.z;
}
// OpenMM-CUDA boundary formats.  Kernels are instantiated twice: NATIVE (all three atom arrays fp64 - the
// integrator-only path, no conversion code at all) and generic (format switch per access).  With a single
// instantiation ptxas speculates the float conversions of the other formats into the fp64 path.
__device__ __forceinline__ d3 loadAtomFormat(const AtomView& A, long long i) {
    if (A.fmt == FMT_POSQ_MIXED) {
        const float4 hi = reinterpret_cast<const float4*>(A.p)[i], lo = reinterpret_cast<const float4*>(A.aux)[i];
        return {(double) hi.x + (double) lo.x, (double) hi.y + (double) lo.y, (double) hi.z + (double) lo.z};
    }
    if (A.fmt == FMT_REAL4_F64) {
        const double2* p = reinterpret_cast<const double2*>(A.p) + 2*i;
        const double2 xy = p[0];
        return {xy.x, xy.y, p[1].x};
    }
    if (A.fmt == FMT_REAL4_F32) {
        const float4 v = reinterpret_cast<const float4*>(A.p)[i];
        return {(double) v.x, (double) v.y, (double) v.z};
    }
    const long long* f = reinterpret_cast<const long long*>(A.p) + i;      // FMT_FORCE_FIXED
    const double scale = 1.0/4294967296.0;
    return {scale*(double) f[0], scale*(double) f[A.sc], scale*(double) f[2*A.sc]};
}
__device__ __forceinline__ void storeAtomFormat(const AtomView& A, long long i, d3 v) {
    if (A.fmt == FMT_POSQ_MIXED) {                             // value = (float) hi + (float) lo, charge (.w) untouched
        float* hi = reinterpret_cast<float*>(A.p) + 4*i;
        float* lo = reinterpret_cast<float*>(A.aux) + 4*i;
        const float hx = (float) v.x, hy = (float) v.y, hz = (float) v.z;
        hi[0] = hx; hi[1] = hy; hi[2] = hz;
        lo[0] = (float) (v.x - (double) hx); lo[1] = (float) (v.y - (double) hy); lo[2] = (float) (v.z - (double) hz);
    }
    else if (A.fmt == FMT_REAL4_F64) {                         // .w (charge / inverse mass) untouched
        double* p = A.p + 4*i;
        *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y);
        p[2] = v.z;
    }
    else if (A.fmt == FMT_REAL4_F32) {
        float* p = reinterpret_cast<float*>(A.p) + 4*i;
        p[0] = (float) v.x; p[1] = (float) v.y; p[2] = (float) v.z;
    }
}
template <bool NATIVE> __device__ __forceinline__ d3 loadAtom(const AtomView& A, long long i) {
    if (NATIVE || A.fmt == FMT_F64) {
        const double* p = A.p + i*A.sa;
        return {p[0], p[A.sc], p[2*A.sc]};
    }
    return loadAtomFormat(A, i);
}
template <bool NATIVE> __device__ __forceinline__ void storeAtom(const AtomView& A, long long i, d3 v) {
    if (NATIVE || A.fmt == FMT_F64) {
        double* p = A.p + i*A.sa;
        p[0] = v.x; p[A.sc] = v.y; p[2*A.sc] = v.z;
    }
    else storeAtomFormat(A, i, v);
}
// The value a later loadAtom will return for v once it has been stored in A's format (identity for fp64).
// Free atoms keep it as savedPos, so that (x - savedPos)/dt in part 2 is exactly zero without constraints,
// whatever precision the caller's position array has.
template <bool NATIVE> __device__ __forceinline__ d3 asStored(const AtomView& A, d3 v) {
    if (NATIVE) return v;
    if (A.fmt == FMT_POSQ_MIXED) {
        const float hx = (float) v.x, hy = (float) v.y, hz = (float) v.z;
        return {(double) hx + (double) (float) (v.x - (double) hx), (double) hy + (double) (float) (v.y - (double) hy),
                (double) hz + (double) (float) (v.z - (double) hz)};
    }
    if (A.fmt == FMT_REAL4_F32) return {(double) (float) v.x, (double) (float) v.y, (double) (float) v.z};
    return v;
}
// Bring atom i's entry of A towards the SM (L2 prefetch of the sectors it occupies) - issued a tile ahead of the loads by kernels
// that cannot keep those loads in flight themselves.
__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ void prefetchAtom(const AtomView& A, long long i) {
    if (A.fmt == FMT_F64) {
        const double* p = A.p + i*A.sa;
        if (A.sc == 1) { prefetchL2(p); prefetchL2(p + 2); }                  // 24 bytes, may straddle two sectors
        else { prefetchL2(p); prefetchL2(p + A.sc); prefetchL2(p + 2*A.sc); }
    }
    else if (A.fmt == FMT_POSQ_MIXED) {
        prefetchL2(reinterpret_cast<const float4*>(A.p) + i);
        prefetchL2(reinterpret_cast<const float4*>(A.aux) + i);
    }
    else if (A.fmt == FMT_REAL4_F64) prefetchL2(A.p + 4*i);
    else if (A.fmt == FMT_REAL4_F32) prefetchL2(reinterpret_cast<const float4*>(A.p) + i);
    else {
        const long long* f = reinterpret_cast<const long long*>(A.p) + i;    // FMT_FORCE_FIXED
        prefetchL2(f); prefetchL2(f + A.sc); prefetchL2(f + 2*A.sc);
    }
}
// plugin-order atom slot -> index in the caller's arrays
__device__ __forceinline__ long long atomSlot(const DeviceSystem& S, int pluginIndex) {
    return S.atomLoc ? (long long) S.atomLoc[pluginIndex] : (long long) pluginIndex;
}

} // namespace
} // namespace rbk
