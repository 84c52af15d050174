// rbk_math.cuh - device math of the rigid-body step: quaternion algebra, NO-SQUISH and exact
// free-rotor propagation.  fp64 throughout, IEEE division/sqrt (never compile with fast-math:
// symmetric tops rely on x/0 = inf semantics, cf. openmmapi/src/RigidBody.cpp:254-259).
//
// What is computed follows the reference's host arithmetic (the parity target):
//   quaternion operators      openmmapi/src/MatVec.cpp:495-549
//   uniaxial / NO-SQUISH      openmmapi/src/RigidBody.cpp:189-231
//   exact rotation            openmmapi/src/RigidBody.cpp:241-308
//   Jacobi / Carlson          openmmapi/include/internal/ellipticFunctions.h:34-245
// How it is computed is organised for the GPU (registers instead of heap objects, fixed-size
// state, reciprocal reuse); results agree with the reference to rounding level.
#pragma once
#include <cfloat>
#include <math.h>

// The same source is compiled for the device by nvcc and (for CPU unit tests of the arithmetic
// only, tests/test_device_math_host.py) for the host by g++.
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define RBK_HD __host__ __device__ __forceinline__
#define RBK_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define RBK_HD inline
#define RBK_HD_NOINLINE inline
#endif

namespace rbk {

RBK_HD double rsqrtd(double x) {
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0/sqrt(x);
#endif
}

struct d3 { double x, y, z; };
struct d4 { double w, x, y, z; };     // scalar-first quaternion (q0,q1,q2,q3)

RBK_HD d3 operator+(d3 a, d3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
RBK_HD d3 operator-(d3 a, d3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
RBK_HD d3 operator*(d3 a, double s) { return {a.x*s, a.y*s, a.z*s}; }
RBK_HD double dot(d3 a, d3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
RBK_HD d3 cross(d3 a, d3 b) { return {a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x}; }
RBK_HD d4 operator+(d4 a, d4 b) { return {a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z}; }
RBK_HD d4 operator*(d4 a, double s) { return {a.w*s, a.x*s, a.y*s, a.z*s}; }
RBK_HD d4 operator-(d4 a, d4 b) { return {a.w - b.w, a.x - b.x, a.y - b.y, a.z - b.z}; }
RBK_HD double dot(d4 a, d4 b) { return a.w*b.w + a.x*b.x + a.y*b.y + a.z*b.z; }

// B(q)v = q (x) (0,v),  C(q)v = (0,v) (x) q  and their transposes
RBK_HD d4 quatB(d4 q, d3 v) {
    return {-q.x*v.x - q.y*v.y - q.z*v.z,
             q.w*v.x - q.z*v.y + q.y*v.z,
             q.z*v.x + q.w*v.y - q.x*v.z,
            -q.y*v.x + q.x*v.y + q.w*v.z};
}
RBK_HD d4 quatC(d4 q, d3 v) {
    return {-q.x*v.x - q.y*v.y - q.z*v.z,
             q.w*v.x + q.z*v.y - q.y*v.z,
            -q.z*v.x + q.w*v.y + q.x*v.z,
             q.y*v.x - q.x*v.y + q.w*v.z};
}
RBK_HD d3 quatBt(d4 q, d4 y) {
    return {-q.x*y.w + q.w*y.x + q.z*y.y - q.y*y.z,
            -q.y*y.w - q.z*y.x + q.w*y.y + q.x*y.z,
            -q.z*y.w + q.y*y.x - q.x*y.y + q.w*y.z};
}
RBK_HD d3 quatCt(d4 q, d4 y) {
    return {-q.x*y.w + q.w*y.x - q.z*y.y + q.y*y.z,
            -q.y*y.w + q.z*y.x + q.w*y.y - q.x*y.z,
            -q.z*y.w - q.y*y.x + q.x*y.y + q.w*y.z};
}
// body frame -> space frame: A^T(q) v = C^T(q) B(q) v
RBK_HD d3 bodyToSpace(d4 q, d3 v) { return quatCt(q, quatB(q, v)); }

// sin and cos for the small angles of a single time step, without range reduction:
//   |x| <= 1/32 : truncated Maclaurin series (remainders x^9/9! and x^10/10! are < 2e-18 relative),
//   |x| <= pi/4 : the classic minimax kernels (error < 1 ulp),
//   otherwise   : the library routine.
RBK_HD void sincosStep(double x, double* s, double* c) {
    const double z = x*x;
    if (z <= 9.765625e-4) {
        *s = x + x*z*(-1.0/6.0 + z*(1.0/120.0 + z*(-1.0/5040.0)));
        *c = 1.0 + z*(-0.5 + z*(1.0/24.0 + z*(-1.0/720.0 + z*(1.0/40320.0))));
    }
    else if (z <= 0.6168502750680849) {
        const double ps = -1.66666666666666324348e-01 + z*(8.33333333332248946124e-03 + z*(-1.98412698298579493134e-04
                        + z*(2.75573137070700676789e-06 + z*(-2.50507602534068634195e-08 + z*1.58969099521155010221e-10))));
        const double pc = 4.16666666666666019037e-02 + z*(-1.38888888888741095749e-03 + z*(2.48015872894767294178e-05
                        + z*(-2.75573143513906633035e-07 + z*(2.08757232129817482790e-09 + z*(-1.13596475577881948265e-11)))));
        *s = x + x*z*ps;
        const double hz = 0.5*z, w = 1.0 - hz;
        *c = w + (((1.0 - w) - hz) + z*z*pc);
    }
    else sincos(x, s, c);
}

// permutation operators B_k q (k = 0,1,2 for principal axes 1,2,3)
template <int K> RBK_HD d4 quatPerm(d4 q) {
    if (K == 0) return {-q.x,  q.w,  q.z, -q.y};
    if (K == 1) return {-q.y, -q.z,  q.w,  q.x};
    return {-q.z,  q.y, -q.x,  q.w};
}

// sin/cos by truncated Maclaurin series, valid (remainder < 2e-18 relative) for |x| <= 1/32; no branches.
RBK_HD void sincosTiny(double x, double* s, double* c) {
    const double z = x*x;
    *s = x + x*z*(-1.0/6.0 + z*(1.0/120.0 + z*(-1.0/5040.0)));
    *c = 1.0 + z*(-0.5 + z*(1.0/24.0 + z*(-1.0/720.0 + z*(1.0/40320.0))));
}

// One uniaxial free rotation about principal axis K; quarterHinvI = h/(4 I_K).
template <int K, bool TINY> RBK_HD void uniaxialScaled(double quarterHinvI, d4& q, d4& pi) {
    const d4 Bq = quatPerm<K>(q);
    const double phi = dot(pi, Bq)*quarterHinvI;
    double s, c;
    if (TINY) sincosTiny(phi, &s, &c);
    else sincosStep(phi, &s, &c);
    const d4 Bp = quatPerm<K>(pi);
    q = q*c + Bq*s;
    pi = pi*c + Bp*s;
}
template <int K> RBK_HD void uniaxial(double h, double invIk, d4& q, d4& pi) { uniaxialScaled<K, false>(0.25*h*invIk, q, pi); }

// NO-SQUISH: n sub-steps of R3(h/2) R2(h/2) R1(h) R2(h/2) R3(h/2); axis 3 skipped for linear bodies.
// Consecutive half rotations about axis 3 of neighbouring sub-steps are flows of the same one-axis
// Hamiltonian and are applied as one rotation over h (identical map, 4n+1 instead of 5n rotations).
template <bool TINY> RBK_HD void noSquishLoop(int n, double k1, double k2, double k3, bool axis3, d4& q, d4& pi) {
    if (axis3) uniaxialScaled<2, TINY>(0.5*k3, q, pi);
    for (int i = 0; i < n; i++) {
        uniaxialScaled<1, TINY>(k2, q, pi);
        uniaxialScaled<0, TINY>(k1, q, pi);
        uniaxialScaled<1, TINY>(k2, q, pi);
        if (axis3) uniaxialScaled<2, TINY>(i == n - 1 ? 0.5*k3 : k3, q, pi);
    }
}

RBK_HD void noSquish(double dt, int n, d3 invI, d4& q, d4& pi) {
    const double h = dt/n;
    const double k1 = 0.25*h*invI.x, k2 = 0.125*h*invI.y, k3 = 0.25*h*invI.z;
    const bool axis3 = invI.z != 0.0;            // dof == 6 (linear bodies carry invI.z = 0)
    // Every rotation angle is (pi . B_k q) k with |B_k q| = |q| and both norms invariant under the rotations
    // (they act orthogonally on q and on pi), so |angle| <= |pi| |q| max(k) throughout: one test per body
    // selects the branch-free small-angle sin/cos for all 4n+1 rotations.
    const double kmax = fmax(fabs(k1), fmax(fabs(k2), fabs(k3)));
    const double bound2 = dot(pi, pi)*dot(q, q)*kmax*kmax;
    if (bound2 <= 9.5e-4) noSquishLoop<true>(n, k1, k2, k3, axis3, q, pi);
    else noSquishLoop<false>(n, k1, k2, k3, axis3, q, pi);
}

// ---------------------------------------------------------------------------------------------
// Elliptic functions
// ---------------------------------------------------------------------------------------------
RBK_HD double sgn(double x) { return x >= 0.0 ? 1.0 : -1.0; }

// sn, cn, dn (u | m) by the arithmetic-geometric-mean scale + descending Landen recurrence.
RBK_HD_NOINLINE void jacobiSnCnDn(double u, double m, double& sn, double& cn, double& dn) {
    const double EPS = DBL_EPSILON;
    if (fabs(m) > 1.0) { sn = cn = dn = nan(""); return; }
    if (fabs(m) < 2.0*EPS) { sincos(u, &sn, &cn); dn = 1.0; return; }
    if (fabs(m - 1.0) < 2.0*EPS) { sn = tanh(u); cn = 1.0/cosh(u); dn = cn; return; }
    double mu[16], nu[16];
    int n = 0;
    mu[0] = 1.0;
    nu[0] = sqrt(1.0 - m);
    const double kp = nu[0];
    while (fabs(mu[n] - nu[n]) > 4.0*EPS*fabs(mu[n] + nu[n])) {
        mu[n+1] = 0.5*(mu[n] + nu[n]);
        nu[n+1] = sqrt(mu[n]*nu[n]);
        ++n;
        if (n >= 15) { sn = cn = dn = nan(""); return; }
    }
    double s, co;
    sincos(u*mu[n], &s, &co);
    const bool small = fabs(s) < fabs(co);
    double c = mu[n]*(small ? s/co : co/s);
    double d = 1.0;
    while (n > 0) {
        n--;
        double r = (c*c)/mu[n+1];
        c = d*c;
        d = (r + nu[n])/(r + mu[n]);
    }
    const double h = rsqrtd(1.0 + c*c);            // 1/hypot(1,c); |c| <= 1 here
    if (small) {
        dn = kp/d;
        cn = dn*sgn(co)*h;
        sn = cn*c/kp;
    }
    else {
        dn = d;
        sn = sgn(s)*h;
        cn = c*sn;
    }
}

// Carlson's degenerate integral RC(x,y) by duplication (tolerance as the reference: 1e-3 -> ~1e-18 truncation)
RBK_HD double carlsonRC(double x, double y) {
    if (x < 0.0 || y < 0.0 || x + y < 5.0*DBL_MIN || x > 0.2*DBL_MAX || y > 0.2*DBL_MAX) return nan("");
    double mu, sn;
    for (int it = 0; ; ) {
        mu = (x + y + y)/3.0;
        sn = (y + mu)/mu - 2.0;
        if (fabs(sn) < 0.001) break;
        double lam = 2.0*sqrt(x)*sqrt(y) + y;
        x = (x + lam)*0.25;
        y = (y + lam)*0.25;
        if (++it == 10000) return nan("");
    }
    double s = sn*sn*(0.3 + sn*(1.0/7.0 + sn*(0.375 + sn*(9.0/22.0))));
    return (1.0 + s)*rsqrtd(mu);
}

// Carlson's RF(x,y,z) by duplication
RBK_HD_NOINLINE double carlsonRF(double x, double y, double z) {
    const double lolim = 5.0*DBL_MIN, uplim = 0.2*DBL_MAX;
    if (x < 0.0 || y < 0.0 || z < 0.0 || x + y < lolim || x + z < lolim || y + z < lolim ||
        x > uplim || y > uplim || z > uplim) return nan("");
    double mu, xd, yd, zd;
    for (int it = 0; ; ) {
        mu = (x + y + z)/3.0;
        const double rmu = 1.0/mu;
        xd = 2.0 - (mu + x)*rmu;
        yd = 2.0 - (mu + y)*rmu;
        zd = 2.0 - (mu + z)*rmu;
        if (fmax(fabs(xd), fmax(fabs(yd), fabs(zd))) < 0.001) break;
        double xr = sqrt(x), yr = sqrt(y), zr = sqrt(z);
        double lam = xr*(yr + zr) + yr*zr;
        x = (x + lam)*0.25;
        y = (y + lam)*0.25;
        z = (z + lam)*0.25;
        if (++it == 10000) return nan("");
    }
    double e2 = xd*yd - zd*zd;
    double e3 = xd*yd*zd;
    double s = 1.0 + ((1.0/24.0)*e2 - 0.1 - (3.0/44.0)*e3)*e2 + (1.0/14.0)*e3;
    return s*rsqrtd(mu);
}

// Carlson's RJ(x,y,z,p) by duplication
RBK_HD_NOINLINE double carlsonRJ(double x, double y, double z, double p) {
    const double lolim = 4.809554074311741e-103;      // (5 DBL_MIN)^(1/3)
    const double uplim = 9.901548214916537e+101;      // 0.3 (0.2 DBL_MAX)^(1/3)
    if (x < 0.0 || y < 0.0 || z < 0.0 || x + y < lolim || x + z < lolim || y + z < lolim || p < lolim ||
        x > uplim || y > uplim || z > uplim || p > uplim) return nan("");
    const double c1 = 3.0/14.0, c2 = 1.0/3.0, c3 = 3.0/22.0, c4 = 3.0/26.0;
    double sigma = 0.0, power4 = 1.0, mu, xd, yd, zd, pd;
    for (int it = 0; ; ) {
        mu = (x + y + z + p + p)*0.2;
        const double rmu = 1.0/mu;
        xd = (mu - x)*rmu;
        yd = (mu - y)*rmu;
        zd = (mu - z)*rmu;
        pd = (mu - p)*rmu;
        if (fmax(fmax(fabs(xd), fabs(yd)), fmax(fabs(zd), fabs(pd))) < 0.001) break;
        double xr = sqrt(x), yr = sqrt(y), zr = sqrt(z);
        double lam = xr*(yr + zr) + yr*zr;
        double alfa = p*(xr + yr + zr) + xr*yr*zr;
        alfa = alfa*alfa;
        double beta = p*(p + lam)*(p + lam);
        double rc = carlsonRC(alfa, beta);
        if (isnan(rc)) return nan("");
        sigma += power4*rc;
        power4 *= 0.25;
        x = (x + lam)*0.25;
        y = (y + lam)*0.25;
        z = (z + lam)*0.25;
        p = (p + lam)*0.25;
        if (++it == 10000) return nan("");
    }
    double ea = xd*(yd + zd) + yd*zd;
    double eb = xd*yd*zd;
    double ec = pd*pd;
    double e2 = ea - 3.0*ec;
    double e3 = eb + 2.0*pd*(ea - ec);
    double s1 = 1.0 + e2*(-c1 + 0.75*c3*e2 - 1.5*c4*e3);
    double s2 = eb*(0.5*c2 + pd*(-c3 - c3 + pd*c4));
    double s3 = pd*ea*(c2 - pd*c3) - c2*pd*ec;
    return 3.0*sigma + power4*(s1 + s2 + s3)/(mu*sqrt(mu));
}

// Omega(x; n, m) = Pi(n; asin x | m) - F(asin x | m) expressed through RJ
RBK_HD double omegaRJ(double x, double n, double m) {
    const double x2 = x*x;
    return (-1.0/3.0)*n*x*x2*carlsonRJ(1.0 - x2, 1.0 - m*x2, 1.0, 1.0 + n*x2);
}

RBK_HD int roundHalfIn(double x) {       // nearest integer, halves towards zero
    return x > 0.0 ? (int) ceil(x - 0.5) : (int) floor(x + 0.5);
}

// Exact torque-free rotation of an asymmetric top over dt (general path through the complete set of
// elliptic integrals).  I, invI: principal moments (descending) and inverses; q, pi updated in place.
RBK_HD_NOINLINE void exactRotationElliptic(double dt, d3 I, d3 invI, d4& q, d4& pi) {
    d3 Iw = quatBt(q, pi)*0.5;
    const d3 w0 = {invI.x*Iw.x, invI.y*Iw.y, invI.z*Iw.z};
    double Lsq = Iw.y*Iw.y + Iw.z*Iw.z;
    if (Lsq < DBL_EPSILON) { uniaxial<0>(dt, invI.x, q, pi); return; }
    Lsq += Iw.x*Iw.x;
    const double L = sqrt(Lsq);
    const double twoKr = dot(Iw, w0);
    const d4 z0 = {Iw.z, Iw.y, L - Iw.x, 0.0};
    const double r1 = Lsq - twoKr*I.z;
    const double r3 = twoKr*I.x - Lsq;
    const double l1 = r1*invI.y/(I.y - I.z);
    const double l3 = r3*invI.y/(I.x - I.y);
    const bool lowBranch = l1 < l3;
    const double lmin = (l3 < l1) ? l3 : l1;
    const double lmax = (l1 < l3) ? l3 : l1;
    const double c13 = 1.0/(I.x - I.z);
    double a0 = sgn(w0.x)*sqrt(r1*invI.x*c13);
    double a1 = sqrt(lmin);
    double a2 = sgn(w0.z)*sqrt(r3*invI.z*c13);
    const double m = lmin/lmax;
    const double K = carlsonRF(0.0, 1.0 - m, 1.0);
    const double inv2K = 0.5/K;
    double s0 = w0.y/a1, c0, u0;
    int i0;
    if (fabs(s0) < 1.0) {
        c0 = lowBranch ? w0.x/a0 : w0.z/a2;
        u0 = s0*carlsonRF(1.0 - s0*s0, 1.0 - m*s0*s0, 1.0);
        i0 = roundHalfIn(u0*inv2K);
    }
    else {
        a1 = fabs(w0.y);
        s0 = sgn(s0);
        c0 = 0.0;
        u0 = s0*K;
        i0 = 0;
    }
    const double wp = -invI.y*a0*a2/(a1*c13);
    const double u = wp*dt + u0;
    const int jump = roundHalfIn(u*inv2K) - i0;
    double sn, cn, dn, deltaF;
    jacobiSnCnDn(u, m, sn, cn, dn);
    const double alpha = I.x*a0/L;
    double eta = alpha*alpha;
    eta /= 1.0 - eta;
    const d3 Ia = {I.x*a0, I.y*a1, I.z*a2};
    if (lowBranch) {
        const double C = sqrt(m + eta);
        deltaF = u - u0 + sgn(cn)*omegaRJ(sn, eta, m) - sgn(c0)*omegaRJ(s0, eta, m)
                        + (alpha/C)*(atan(C*sn/dn) - atan(C*s0*a2/w0.z));
        if (jump != 0) deltaF += jump*2.0*omegaRJ(1.0, eta, m);
        Iw = {Ia.x*cn, Ia.y*sn, Ia.z*dn};
    }
    else {
        const double k2eta = m*eta;
        const double C = sqrt(1.0 + k2eta);
        deltaF = u - u0 + sgn(cn)*omegaRJ(sn, k2eta, m) - sgn(c0)*omegaRJ(s0, k2eta, m)
                        + (alpha/C)*(atan(C*sn/cn) - atan(C*s0/c0));
        if (jump != 0) deltaF += jump*(2.0*omegaRJ(1.0, k2eta, m) + (alpha/C)*3.14159265358979323846264338328);
        Iw = {Ia.x*dn, Ia.y*sn, Ia.z*cn};
    }
    deltaF *= 1.0 + eta;
    const double theta = (Lsq*(u - u0) + r3*deltaF)/(2.0*L*I.x*wp);
    double st, ct;
    sincos(theta, &st, &ct);
    const d4 za = {Iw.z, Iw.y, L - Iw.x, 0.0};
    const d4 zb = {-Iw.y, Iw.z, 0.0, L - Iw.x};
    const d4 z = za*ct + zb*st;
    d4 qn = z*dot(z0, q) + quatC(z, quatCt(z0, q));
    qn = qn*(1.0/sqrt(dot(qn, qn)));
    q = qn;
    pi = quatB(qn, Iw*2.0);
}

// ---------------------------------------------------------------------------------------------
// Exact torque-free rotation, fast path: Taylor series of Euler's equations.
//
// In the body frame the angular momentum l(t) obeys dl/dt = l x (l/I) (quadratic right-hand side), so
// its Taylor coefficients in s = t/dt follow from three Cauchy products per order.  The reference's
// rotation angle (RigidBody.cpp:286-302) is, after inserting its Omega/atan terms and simplifying with
// 1+eta = 1/(1-alpha^2) and alpha*cn (resp. alpha*dn) = l0/L, nothing but
//        theta(dt) = 1/2 * Int_0^dt [ L/I0 + (2T - L^2/I0) / (L - l0(t)) ] dt
// in BOTH of its branches, so theta needs only the reciprocal series of L - l0(s), integrated term by
// term.  No elliptic function, no division inside the recurrences, no data-dependent loop: ~2 K^2 FMAs.
// The series are mathematically exact; truncation is checked at run time from the size of the last
// terms and the caller falls back to exactRotationElliptic when the check fails (large |omega| dt).
// Also valid where the elliptic route is not (symmetric, spherical and linear tops).
// ---------------------------------------------------------------------------------------------
// Returns 0 when the truncation check passed (q, pi advanced), otherwise by how much the tail exceeds its bound
// (a ratio > 1, possibly inf/NaN; q, pi untouched) - what the retry path needs to pick its number of sub-steps.
//
// Choice of the reduction axis.  theta's integrand has the pole L - l0(t) = 0: with the angular momentum almost along
// +axis 1 the reciprocal series converges slowly (or not within a step).  Nothing in the derivation singles out axis 1 -
// the same formulas hold for any body axis taken as "axis 1" after a cyclic relabelling (x,y,z) -> (y,z,x) or (z,x,y),
// which is the proper rotation q -> q (x) 1/2(1,+-1,+-1,+-1) of the body frame.  Each body therefore uses the axis on which
// its angular momentum has the SMALLEST (most negative) component, so that L - l0 >= 0.42 L: at 300 K and 1 fs no water in
// 4096 then fails order 12 (0.7 % with the fixed axis; order 14 was needed before), and the nearly linear triatomics of the
// mixed benchmark that needed up to 40 sub-steps converge directly.
RBK_HD d4 quatTimesDiag(d4 q, double s) {                  // q (x) 1/2 (1, s, s, s), s = +-1
    return {0.5*(q.w - s*(q.x + q.y + q.z)), 0.5*(q.x + s*(q.w + q.y - q.z)),
            0.5*(q.y + s*(q.w - q.x + q.z)), 0.5*(q.z + s*(q.w + q.x - q.y))};
}

// KLOW > 0: also report (lowerFails) whether the truncation check of an order-KLOW series would have failed on this body -
// the same test on orders KLOW-1, KLOW, six additions - so that a caller running a ladder of orders can step down.
template <int K, int KLOW>
RBK_HD double exactRotationSeriesAxis1(double dt, d3 invI, d3 l0, d4& q, d3& lOut, bool& lowerFails);

template <int K, int KLOW = 0>
RBK_HD double exactRotationSeries(double dt, d3 invI, d4& q, d4& pi, bool* lowerFails = nullptr) {
    const d3 l = quatBt(q, pi)*0.5;
    const int j = l.y < l.x ? (l.z < l.y ? 2 : 1) : (l.z < l.x ? 2 : 0);       // argmin: the new "axis 1"
    const double s = j == 1 ? 1.0 : -1.0;
    const d3 lp = j == 0 ? l : (j == 1 ? d3{l.y, l.z, l.x} : d3{l.z, l.x, l.y});
    const d3 invIp = j == 0 ? invI : (j == 1 ? d3{invI.y, invI.z, invI.x} : d3{invI.z, invI.x, invI.y});
    d4 qp = j == 0 ? q : quatTimesDiag(q, s);
    d3 ln;
    bool low = false;
    const double excess = exactRotationSeriesAxis1<K, KLOW>(dt, invIp, lp, qp, ln, low);
    if (KLOW > 0 && lowerFails != nullptr) *lowerFails = low;
    if (excess != 0.0) return excess;
    q = j == 0 ? qp : quatTimesDiag(qp, -s);
    const d3 lb = j == 0 ? ln : (j == 1 ? d3{ln.z, ln.x, ln.y} : d3{ln.y, ln.z, ln.x});
    pi = quatB(q, lb*2.0);
    return 0.0;
}

// The series proper, reduction about axis 1 of the (relabelled) body frame: advances q and returns the new body-frame
// angular momentum in lOut.
template <int K, int KLOW>
RBK_HD double exactRotationSeriesAxis1(double dt, d3 invI, d3 l0, d4& q, d3& lOut, bool& lowerFails) {
    lowerFails = false;
    if (l0.y*l0.y + l0.z*l0.z < DBL_EPSILON) {             // rotation about axis 1 itself
        d4 pi = quatB(q, l0*2.0);
        uniaxial<0>(dt, invI.x, q, pi);
        lOut = quatBt(q, pi)*0.5;
        return 0.0;
    }
    const double Lsq = l0.y*l0.y + l0.z*l0.z + l0.x*l0.x;
    const double L = sqrt(Lsq);
    const double twoT = l0.x*l0.x*invI.x + l0.y*l0.y*invI.y + l0.z*l0.z*invI.z;
    double x[K + 1], y[K + 1], z[K + 1], r[K + 1];
    x[0] = l0.x; y[0] = l0.y; z[0] = l0.z;
    const double ca = dt*(invI.z - invI.y), cb = dt*(invI.x - invI.z), cc = dt*(invI.y - invI.x);
    r[0] = 1.0/(L - x[0]);
    double sx = x[0], sy = y[0], sz = z[0], sr = r[0];
    // Order k+1 from orders 0..k.  Each Cauchy product is summed so that the terms containing the
    // NEWEST coefficients (index k) come last: everything else of order k+1 only needs orders < k and
    // overlaps with the tail of order k, which keeps the fp64 pipe fed from a single warp.
#pragma unroll
    for (int k = 0; k < K; k++) {
        double ax = 0.0, ay = 0.0, az = 0.0, bx = 0.0, by = 0.0, bz = 0.0;
#pragma unroll
        for (int j = 1; j <= k - 1; j += 2) {
            ax = fma(y[j], z[k - j], ax);
            ay = fma(z[j], x[k - j], ay);
            az = fma(x[j], y[k - j], az);
            if (j + 1 <= k - 1) {
                bx = fma(y[j + 1], z[k - j - 1], bx);
                by = fma(z[j + 1], x[k - j - 1], by);
                bz = fma(x[j + 1], y[k - j - 1], bz);
            }
        }
        double ex, ey, ez;
        if (k == 0) { ex = y[0]*z[0]; ey = z[0]*x[0]; ez = x[0]*y[0]; }
        else {
            ex = fma(y[0], z[k], y[k]*z[0]);
            ey = fma(z[0], x[k], z[k]*x[0]);
            ez = fma(x[0], y[k], x[k]*y[0]);
        }
        const double f = 1.0/(k + 1);
        x[k + 1] = (ca*f)*((ax + bx) + ex);
        y[k + 1] = (cb*f)*((ay + by) + ey);
        z[k + 1] = (cc*f)*((az + bz) + ez);
        sx += x[k + 1]; sy += y[k + 1]; sz += z[k + 1];
        // reciprocal series: r[k+1] = r0 * sum_{j=1..k+1} x[j] r[k+1-j]; the term with the newest r (j = 1) and
        // the one with the newest x (j = k+1) come last
        double ar = 0.0, br = 0.0;
#pragma unroll
        for (int j = 2; j <= k; j += 2) {
            ar = fma(x[j], r[k + 1 - j], ar);
            if (j + 1 <= k) br = fma(x[j + 1], r[k - j], br);
        }
        double er = x[k + 1]*r[0];
        if (k >= 1) er = fma(x[1], r[k], er);
        r[k + 1] = ((ar + br) + er)*r[0];
        sr = fma(r[k + 1], 1.0/(k + 2), sr);
    }
    const double tailL = fabs(x[K]) + fabs(y[K]) + fabs(z[K]) + fabs(x[K - 1]) + fabs(y[K - 1]) + fabs(z[K - 1]);
    // truncation check on the last two orders of every series (relative to L, resp. r[0])
    const double tailR = fabs(r[K]) + fabs(r[K - 1]);
    if (KLOW > 0) {
        constexpr int J = KLOW > 0 && KLOW <= K ? KLOW : K;
        const double lowL = fabs(x[J]) + fabs(y[J]) + fabs(z[J]) + fabs(x[J - 1]) + fabs(y[J - 1]) + fabs(z[J - 1]);
        lowerFails = !(lowL <= 2.0e-16*L && fabs(r[J]) + fabs(r[J - 1]) <= 2.0e-16*fabs(r[0]));
    }
#ifndef RBK_EXPERIMENT_SKIP_CHECK
    if (!(tailL <= 2.0e-16*L && tailR <= 2.0e-16*fabs(r[0]))) return fmax(tailL/(2.0e-16*L), tailR/(2.0e-16*fabs(r[0])));
#endif
    const double theta = 0.5*dt*(L*invI.x + (twoT - Lsq*invI.x)*sr);
    double st, ct;
    sincosStep(theta, &st, &ct);
    const d4 z0 = {l0.z, l0.y, L - l0.x, 0.0};
    const d4 za = {sz, sy, L - sx, 0.0};
    const d4 zb = {-sy, sz, 0.0, L - sx};
    const d4 zz = za*ct + zb*st;
    d4 qn = zz*dot(z0, q) + quatC(zz, quatCt(z0, q));
    qn = qn*rsqrtd(dot(qn, qn));
    q = qn;
    lOut = {sx, sy, sz};
    return 0.0;
}

#ifndef RBK_SERIES_ORDER
#define RBK_SERIES_ORDER 12
#endif
constexpr int kSeriesOrder = RBK_SERIES_ORDER;

// Slow path of mode 0 (rare; kept out of line so that the fast path stays straight-line code).  The composition of
// exact flows is exact, and with n sub-steps the order-k coefficient shrinks by n^k: the number of sub-steps is taken
// from how far the failed attempt overshot its bound (excess^(1/(K-1)), with a margin), so that one retry normally
// suffices - a slow body holds up its whole warp, and with few tiles per CTA the whole launch.  Only when that would
// need more than kMaxSubSteps, or still fails, the elliptic-integral route (which needs I = 1/invI) takes over.
constexpr int kMaxSubSteps = 32;
template <int K>
RBK_HD_NOINLINE void exactRotationRetry(double dt, d3 invI, d4& q, d4& pi, double excess) {
    const double want = 1.25*exp2(log2(excess)/(K - 1));
    if (want <= (double) kMaxSubSteps) {                       // false for inf / NaN as well
        const int n = want < 2.0 ? 2 : (int) ceil(want);
        d4 q1 = q, p1 = pi;
        const double h = dt/n;
        bool ok = true;
        for (int i = 0; i < n && ok; i++) ok = exactRotationSeries<K>(h, invI, q1, p1) == 0.0;
        if (ok) { q = q1; pi = p1; return; }
    }
#ifndef RBK_EXPERIMENT_NO_ELLIPTIC
    const d3 I = {1.0/invI.x, 1.0/invI.y, 1.0/invI.z};
    exactRotationElliptic(dt, I, invI, q, pi);
#endif
}

// Mode 0 entry point used by the step: one series step over dt; if its truncation check fails (fast rotor /
// long step) the retry path above.
RBK_HD void exactRotation(double dt, d3 invI, d4& q, d4& pi) {
    const double excess = exactRotationSeries<kSeriesOrder>(dt, invI, q, pi);
    if (excess != 0.0) exactRotationRetry<kSeriesOrder>(dt, invI, q, pi, excess);
}

// The same with the series order chosen at run time from a short ladder (the hot water kernels; the choice is uniform over
// a launch, so only one of the inlined copies is ever in the instruction cache).  A body-step costs ~2 K^2 FMAs and order K
// covers |omega| dt up to ~(2e-16)^(1/K): for TIP3P water at 300 K order 11 fails its check for 1e-5 of the bodies at 1 fs
// (order 12: none; order 10: 0.5 %), order 13 for 1.5e-4 at 2 fs, order 16 for 3e-4 at 4 fs (tools/series_stats.py).
// flags: bit 0 = the check failed at this order (the retry path ran), bit 1 = it would have failed one rung lower.
constexpr int kSeriesLadder[3] = {11, 13, 16};
template <int K, int KLOW>
RBK_HD void exactRotationRung(double dt, d3 invI, d4& q, d4& pi, unsigned& flags) {
    bool low = false;
    const double excess = exactRotationSeries<K, KLOW>(dt, invI, q, pi, &low);
    flags = (excess != 0.0 ? 1u : 0u) | (low ? 2u : 0u);
    if (excess != 0.0) exactRotationRetry<K>(dt, invI, q, pi, excess);
}
RBK_HD void exactRotationLadder(int rung, double dt, d3 invI, d4& q, d4& pi, unsigned& flags) {
    if (rung <= 0) exactRotationRung<11, 0>(dt, invI, q, pi, flags);
    else if (rung == 1) exactRotationRung<13, 11>(dt, invI, q, pi, flags);
    else exactRotationRung<16, 13>(dt, invI, q, pi, flags);
}

} // namespace rbk
