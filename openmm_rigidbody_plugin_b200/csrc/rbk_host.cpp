// rbk_host.cpp - host-side rigid-body model (see rbk_host.hpp).  Compile with -ffp-contract=off.
#include "rbk_host.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace rbk {
namespace {

constexpr double kEps = DBL_EPSILON;
constexpr double kPi = 3.14159265358979323846264338328;

struct V3 {
    double x, y, z;
    double& operator[](int i) { return (&x)[i]; }
    double operator[](int i) const { return (&x)[i]; }
};
inline V3 ld(const double* p) { return V3{p[0], p[1], p[2]}; }
inline void st(double* p, const V3& a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, double s) { return V3{a.x*s, a.y*s, a.z*s}; }
inline double dot(V3 a, V3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
inline V3 cross(V3 a, V3 b) { return V3{a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x}; }
// OpenMM::Vec3 divides by multiplying with the reciprocal
inline V3 divide(V3 a, double s) { double k = 1.0/s; return a*k; }

struct M3 { V3 r[3]; };   // rows

// quaternion-matrix products, scalar-first q (openmmapi/src/MatVec.cpp:495-522)
inline void qB(const double* q, V3 v, double* o) {
    o[0] = -q[1]*v.x - q[2]*v.y - q[3]*v.z;
    o[1] =  q[0]*v.x - q[3]*v.y + q[2]*v.z;
    o[2] =  q[3]*v.x + q[0]*v.y - q[1]*v.z;
    o[3] = -q[2]*v.x + q[1]*v.y + q[0]*v.z;
}
inline void qC(const double* q, V3 v, double* o) {
    o[0] = -q[1]*v.x - q[2]*v.y - q[3]*v.z;
    o[1] =  q[0]*v.x + q[3]*v.y - q[2]*v.z;
    o[2] = -q[3]*v.x + q[0]*v.y + q[1]*v.z;
    o[3] =  q[2]*v.x - q[1]*v.y + q[0]*v.z;
}
inline V3 qBt(const double* q, const double* y) {
    return V3{-q[1]*y[0] + q[0]*y[1] + q[3]*y[2] - q[2]*y[3],
              -q[2]*y[0] - q[3]*y[1] + q[0]*y[2] + q[1]*y[3],
              -q[3]*y[0] + q[2]*y[1] - q[1]*y[2] + q[0]*y[3]};
}
inline V3 spaceToBody(const double* q, V3 v) { double t[4]; qC(q, v, t); return qBt(q, t); }   // A(q) v

// ---- analytic eigen-decomposition of a symmetric 3x3 matrix (eigenDecomposition.cpp) ----------

V3 principalMoments(const M3& A) {                      // eigenvalues, descending (:72-103)
    double offsq = A.r[0].y*A.r[0].y + A.r[0].z*A.r[0].z + A.r[1].z*A.r[1].z;
    V3 w{A.r[0].x, A.r[1].y, A.r[2].z};
    if (offsq < kEps) {
        if (w.x < w.y) std::swap(w.x, w.y);
        if (w.x < w.z) std::swap(w.x, w.z);
        if (w.y < w.z) std::swap(w.y, w.z);
        return w;
    }
    double trace = w.x + w.y + w.z;
    double mean = trace/3.0;
    V3 dev{w.x - mean, w.y - mean, w.z - mean};
    double p2 = dot(dev, dev) + 2.0*offsq;
    double p = std::sqrt(p2/6.0);
    V3 a0{A.r[0].x - mean, A.r[0].y, A.r[0].z};
    V3 a1{A.r[1].x, A.r[1].y - mean, A.r[1].z};
    V3 a2{A.r[2].x, A.r[2].y, A.r[2].z - mean};
    double det = a0.x*(a1.y*a2.z - a2.y*a1.z) - a0.y*(a1.x*a2.z - a2.x*a1.z) + a0.z*(a1.x*a2.y - a2.x*a1.y);
    double r = det*(3.0/(p*p2));
    double phi = r <= -1.0 ? kPi/3.0 : (r >= 1.0 ? 0.0 : std::acos(r)/3.0);
    double hi = mean + 2.0*p*std::cos(phi);
    double lo = mean + 2.0*p*std::cos(phi + 2.0*kPi/3.0);
    return V3{hi, trace - (hi + lo), lo};
}

void finishAxis(V3& v, const M3& a, double n1base, double n2base, double thresh) {   // (:36-68)
    double norm = dot(v, v);
    double n1 = n1base + a.r[0].x*a.r[0].x;
    double n2 = n2base + a.r[1].y*a.r[1].y;
    double error = n1*n2;
    if (n1 <= thresh) v = V3{1.0, 0.0, 0.0};
    else if (n2 <= thresh) v = V3{0.0, 1.0, 0.0};
    else if (norm < 4096.0*kEps*kEps*error) {
        double t = std::fabs(a.r[0].y);
        double f = -a.r[0].x/a.r[0].y;
        if (std::fabs(a.r[1].y) > t) { t = std::fabs(a.r[1].y); f = -a.r[0].y/a.r[1].y; }
        if (std::fabs(a.r[1].z) > t) f = -a.r[0].z/a.r[1].z;
        norm = 1.0/std::sqrt(1.0 + f*f);
        v = V3{norm, f*norm, 0.0};
    }
    else v = v*std::sqrt(1.0/norm);
}

M3 principalAxes(const M3& A, V3 w) {                   // eigenvectors as rows (:107-176)
    double tiny = 8.0*kEps*std::fabs(w.x);
    double thresh = tiny*tiny;
    M3 a = A;
    a.r[1].x = a.r[0].y; a.r[2].x = a.r[0].z; a.r[2].y = a.r[1].z;
    double n1 = a.r[0].y*a.r[0].y + a.r[0].z*a.r[0].z;
    double n2 = a.r[0].y*a.r[0].y + a.r[1].z*a.r[1].z;
    double c0 = a.r[0].y*a.r[1].z - a.r[0].z*a.r[1].y;
    double c1 = a.r[0].z*a.r[0].y - a.r[1].z*a.r[0].x;
    double c2 = a.r[0].y*a.r[0].y;
    a.r[0].x -= w.x;
    a.r[1].y -= w.x;
    V3 e0{c0 + a.r[0].z*w.x, c1 + a.r[1].z*w.x, a.r[0].x*a.r[1].y - c2};
    finishAxis(e0, a, n1, n2, thresh);
    V3 e1{c0, c1, c2};
    double gap = w.x - w.y;
    if (std::fabs(gap) > tiny) {
        a.r[0].x += gap;
        a.r[1].y += gap;
        e1 = V3{c0 + a.r[0].z*w.y, c1 + a.r[1].z*w.y, a.r[0].x*a.r[1].y - c2};
        finishAxis(e1, a, n1, n2, thresh);
    }
    else {                                              // degenerate leading pair
        a.r[0].x += w.x;
        a.r[1].y += w.x;
        bool ok = false;
        for (int i = 0; i < 3 && !ok; i++) {
            a.r[i][i] -= w.y;
            V3 col{a.r[0][i], a.r[1][i], a.r[2][i]};
            n1 = dot(col, col);
            ok = n1 > thresh;
            if (ok) {
                e1 = cross(e0, col);
                double norm = dot(e1, e1);
                ok = norm > 65536.0*kEps*kEps*n1;
                if (ok) e1 = e1*std::sqrt(1.0/norm);
            }
        }
        if (!ok) {                                      // fully degenerate: any axis orthogonal to e0
            int i = 0;
            while (i < 2 && e0[i] == 0.0) i++;
            int j = i % 3;
            double norm = 1.0/std::sqrt(e0[i]*e0[i] + e0[j]*e0[j]);
            e1[i] = e0[j]*norm;
            e1[j] = -e0[i]*norm;
            if (i + 1 < 3) e1[i+1] = 0.0;
        }
    }
    return M3{{e0, e1, cross(e0, e1)}};
}

void quaternionOf(const M3& A, double* q) {             // Shepperd (MatVec.cpp:344-373)
    double t0 = A.r[0].x, t1 = A.r[1].y, t2 = A.r[2].z;
    double cand[4] = {1.0 + t0 + t1 + t2, 1.0 + t0 - t1 - t2, 1.0 - t0 + t1 - t2, 1.0 - t0 - t1 + t2};
    int k = 0;
    for (int i = 1; i < 4; i++) if (cand[i] > cand[k]) k = i;
    double big = cand[k];
    double f = 0.5/std::sqrt(big);
    switch (k) {
    case 0: q[1] = (A.r[1].z - A.r[2].y)*f; q[2] = (A.r[2].x - A.r[0].z)*f; q[3] = (A.r[0].y - A.r[1].x)*f; break;
    case 1: q[0] = (A.r[1].z - A.r[2].y)*f; q[2] = (A.r[0].y + A.r[1].x)*f; q[3] = (A.r[0].z + A.r[2].x)*f; break;
    case 2: q[0] = (A.r[2].x - A.r[0].z)*f; q[1] = (A.r[0].y + A.r[1].x)*f; q[3] = (A.r[1].z + A.r[2].y)*f; break;
    default: q[0] = (A.r[0].y - A.r[1].x)*f; q[1] = (A.r[0].z + A.r[2].x)*f; q[2] = (A.r[1].z + A.r[2].y)*f; break;
    }
    q[k] = big*f;
}

// RigidBody.cpp:26-45
bool isCollinear(int N, const double* delta, const double* d2, V3& axis) {
    const double TOL = 1.0E-5;
    double first = d2[0], longest = d2[0];
    int jmax = 0;
    for (int j = 1; j < N; j++) if (d2[j] > first) { jmax = j; longest = d2[j]; }
    axis = divide(ld(delta + 3*jmax), std::sqrt(longest));
    bool line = true;
    for (int j = 0; line && j < N; j++) {
        double proj = dot(axis, ld(delta + 3*j));
        line = line && (d2[j] < TOL*longest || std::fabs(proj*proj/d2[j] - 1.0) < TOL);
    }
    return line;
}

// u u^T complement: (u.u) 1 - u u^T   (MatVec.cpp:555-564)
M3 complementProjector(V3 u) {
    M3 P;
    double uu = dot(u, u);
    for (int i = 0; i < 3; i++) {
        P.r[i] = V3{(-u.x)*u[i], (-u.y)*u[i], (-u.z)*u[i]};
        P.r[i][i] += uu;
    }
    return P;
}

} // namespace

// ---- RigidBodySystem::initialize (RigidBodySystem.cpp:28-114) ------------------------------------
std::string HostModel::initialize(int nAtoms, const int* labels, const double* masses, const unsigned char* virt,
                                  int nConstraints, const int* constraintAtoms, int mode) {
    if (nAtoms <= 0 || labels == nullptr || masses == nullptr) return "rbk_create: empty system";
    if (mode < 0) return "Rotation mode cannot be negative";
    numAtoms = nAtoms;
    rotationMode = mode;
    numConstraints = nConstraints;
    mass.assign(masses, masses + nAtoms);
    isVirtual.assign(nAtoms, 0);
    if (virt) for (int i = 0; i < nAtoms; i++) isVirtual[i] = virt[i] ? 1 : 0;

    // compact labels: distinct positive values -> 1..nB in ascending order of the value (sort + binary search, so
    // that sparse labels such as residue serial numbers in the billions cost O(N log N), not O(max label) memory)
    int top = *std::max_element(labels, labels + nAtoms);
    if (top < 0) return "bodyIndices has no non-negative entry";
    std::vector<int> distinct;
    for (int i = 0; i < nAtoms; i++) if (labels[i] > 0) distinct.push_back(labels[i]);
    std::sort(distinct.begin(), distinct.end());
    distinct.erase(std::unique(distinct.begin(), distinct.end()), distinct.end());
    numBodies = (int) distinct.size();
    bodyIndex.resize(nAtoms);
    for (int i = 0; i < nAtoms; i++)
        bodyIndex[i] = labels[i] > 0 ? (int) (std::lower_bound(distinct.begin(), distinct.end(), labels[i]) - distinct.begin()) + 1 : 0;

    numActualAtoms = nAtoms;
    for (int i = 0; i < nAtoms; i++) if (isVirtual[i]) numActualAtoms--;
    atomIndex.assign((size_t) numActualAtoms, 0);
    body.assign((size_t) numBodies, HostBody());
    numFree = 0;
    for (int i = 0; i < nAtoms; i++) {
        if (isVirtual[i] || mass[i] == 0.0) continue;
        if (bodyIndex[i] == 0) atomIndex[numFree++] = i;
        else body[bodyIndex[i]-1].N++;
    }
    numBodyAtoms = numActualAtoms - numFree;
    d.assign(3*(size_t) numBodyAtoms, 0.0);
    delta.assign(3*(size_t) numBodyAtoms, 0.0);
    freeInvMass.resize((size_t) numFree);
    for (int k = 0; k < numFree; k++) freeInvMass[k] = 1.0/mass[atomIndex[k]];
    int loc = 0;
    for (auto& b : body) { b.loc = loc; loc += b.N; }
    std::vector<int> fill((size_t) numBodies, 0);
    for (int i = 0; i < nAtoms; i++) {
        int ib = bodyIndex[i];
        if (ib > 0) {
            size_t slot = (size_t) numFree + body[ib-1].loc + fill[ib-1]++;
            if (slot < atomIndex.size()) atomIndex[slot] = i;
        }
    }
    // the reference counts N without virtual/massless atoms but fills all labelled atoms; a body that
    // contains such atoms is malformed there too - reject it here instead of corrupting the layout
    for (int b = 0; b < numBodies; b++)
        if (fill[b] != body[b].N) return "rigid bodies must not contain virtual sites or massless particles";
    for (int c = 0; c < nConstraints; c++) {
        int a1 = constraintAtoms[2*c], a2 = constraintAtoms[2*c+1];
        if (a1 < 0 || a1 >= nAtoms || a2 < 0 || a2 >= nAtoms) return "rbk_create: constraint atom out of range";
        if (bodyIndex[a1] != 0 || bodyIndex[a2] != 0) return "Constraints involving rigid-body atoms are not allowed";
    }
    return std::string();
}

// ---- RigidBody::buildGeometry (RigidBody.cpp:65-116) ------------------------------------------------
void HostModel::buildGeometry(HostBody& b, const double* R, const double* F) {
    const int* atom = atomIndex.data() + numFree + b.loc;
    double* dl = delta.data() + 3*(size_t) b.loc;
    double* bf = d.data() + 3*(size_t) b.loc;
    const int N = b.N;
    b.mass = 0.0;
    V3 com{0.0, 0.0, 0.0};
    for (int j = 0; j < N; j++) {
        double m = mass[atom[j]];
        b.mass += m;
        com = com + ld(R + 3*(size_t) atom[j])*m;
    }
    com = divide(com, b.mass);
    st(b.rcm, com);
    b.invMass = 1.0/b.mass;
    std::vector<double> d2((size_t) std::max(N, 1));
    for (int j = 0; j < N; j++) {
        V3 x = ld(R + 3*(size_t) atom[j]) - com;
        st(dl + 3*j, x);
        d2[j] = dot(x, x);
    }
    V3 axis;
    M3 A;
    if (isCollinear(N, dl, d2.data(), axis)) {
        double moment = 0.0;
        for (int j = 0; j < N; j++) moment += mass[atom[j]]*d2[j];
        b.I[0] = b.I[1] = moment; b.I[2] = 0.0;
        b.invI[0] = b.invI[1] = 1.0/moment; b.invI[2] = 0.0;
        // a unit vector orthogonal to the axis (RigidBody.cpp:51-58)
        int imin = axis.x < axis.y ? 0 : 1;
        if (axis.z < axis[imin]) imin = 2;
        V3 e{0.0, 0.0, 0.0};
        e[imin] = 1.0;
        M3 P = complementProjector(axis);
        V3 v{dot(P.r[0], e), dot(P.r[1], e), dot(P.r[2], e)};
        v = divide(v, std::sqrt(dot(v, v)));
        V3 w = cross(axis, v);
        for (int r = 0; r < 3; r++) A.r[r] = V3{v[r], w[r], axis[r]};     // transpose of rows (v, axis x v, axis)
        b.dof = 5;
    }
    else {
        M3 inertia{{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}};
        for (int j = 0; j < N; j++) {
            M3 P = complementProjector(ld(dl + 3*j));
            double m = mass[atom[j]];
            for (int r = 0; r < 3; r++) inertia.r[r] = inertia.r[r] + P.r[r]*m;
        }
        V3 w = principalMoments(inertia);
        st(b.I, w);
        for (int c = 0; c < 3; c++) b.invI[c] = 1.0/b.I[c];
        A = principalAxes(inertia, w);
        b.dof = 6;
    }
    quaternionOf(A, b.q);
    for (int j = 0; j < N; j++) {
        V3 x = ld(dl + 3*j);
        st(bf + 3*j, V3{dot(A.r[0], x), dot(A.r[1], x), dot(A.r[2], x)});
    }
    // resultant force and torque (RigidBody.cpp:174-183)
    V3 f{0.0, 0.0, 0.0}, t{0.0, 0.0, 0.0};
    for (int j = 0; j < N; j++) {
        V3 fj = ld(F + 3*(size_t) atom[j]);
        f = f + fj;
        t = t + cross(ld(dl + 3*j), fj);
    }
    st(b.force, f);
    st(b.tau, t);
    qC(b.q, t, b.torque);
}

// ---- RigidBody::buildDynamics (RigidBody.cpp:123-142).  p = sum m v is SET here, like the GPU-side build
// (rbk_build.cu).  The reference adds into its previous pcm (RigidBody.cpp:126-130: no reset), which doubles the
// momentum when velocities are set twice; in its own start-up protocol (setPositions with V = 0, then
// setVelocities) pcm is zero before the only call that matters, so the two agree wherever the reference is right.
void HostModel::buildDynamics(HostBody& b, const double* V) {
    const int* atom = atomIndex.data() + numFree + b.loc;
    V3 p{0.0, 0.0, 0.0};
    for (int j = 0; j < b.N; j++) p = p + ld(V + 3*(size_t) atom[j])*mass[atom[j]];
    st(b.pcm, p);
    V3 vcm = divide(p, b.mass);
    b.twoKt = dot(p, vcm);
    V3 L{0.0, 0.0, 0.0};
    for (int j = 0; j < b.N; j++) {
        V3 rel = ld(V + 3*(size_t) atom[j]) - vcm;
        V3 inBody = spaceToBody(b.q, rel)*mass[atom[j]];
        L = L + cross(ld(d.data() + 3*(size_t) (b.loc + j)), inBody);
    }
    qB(b.q, L, b.pi);
    for (int c = 0; c < 4; c++) b.pi[c] *= 2.0;
    V3 w{b.invI[0]*L.x, b.invI[1]*L.y, b.invI[2]*L.z};
    b.twoKr = dot(L, w);
}

// ---- RigidBodySystem::update (RigidBodySystem.cpp:120-142) ----------------------------------------
void HostModel::update(const double* R, const double* V, const double* F, bool geometry, bool velocities) {
    if (geometry) {
        numDOF = numFree - numConstraints;
        for (auto& b : body) { buildGeometry(b, R, F); numDOF += b.dof; }
    }
    if (velocities)
        for (auto& b : body) buildDynamics(b, V);
}

} // namespace rbk
