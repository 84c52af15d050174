/* Minimal stand-in for OpenMM's openmm/Vec3.h, written for this repo so that the
 * reference plugin's CPU sources can be compiled in place (oracle/_ref) without an
 * OpenMM install.  TEST INFRASTRUCTURE ONLY - never linked into the product library.
 *
 * Semantics follow OpenMM 7.x: three doubles, component-wise arithmetic, and division
 * by a scalar implemented as multiplication by (1.0/rhs).
 */
#ifndef RBK_SHIM_OPENMM_VEC3_H_
#define RBK_SHIM_OPENMM_VEC3_H_
#include <cassert>
#include <iosfwd>

namespace OpenMM {

class Vec3 {
public:
    Vec3() { v[0] = v[1] = v[2] = 0.0; }
    Vec3(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; }
    double operator[](int i) const { assert(i >= 0 && i < 3); return v[i]; }
    double& operator[](int i) { assert(i >= 0 && i < 3); return v[i]; }
    bool operator==(const Vec3& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
    bool operator!=(const Vec3& o) const { return !(*this == o); }
    Vec3 operator+() const { return *this; }
    Vec3 operator-() const { return Vec3(-v[0], -v[1], -v[2]); }
    Vec3 operator+(const Vec3& o) const { return Vec3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vec3& operator+=(const Vec3& o) { v[0] += o.v[0]; v[1] += o.v[1]; v[2] += o.v[2]; return *this; }
    Vec3& operator-=(const Vec3& o) { v[0] -= o.v[0]; v[1] -= o.v[1]; v[2] -= o.v[2]; return *this; }
    Vec3 operator*(double s) const { return Vec3(v[0]*s, v[1]*s, v[2]*s); }
    Vec3& operator*=(double s) { v[0] *= s; v[1] *= s; v[2] *= s; return *this; }
    Vec3 operator/(double s) const { double k = 1.0/s; return Vec3(v[0]*k, v[1]*k, v[2]*k); }
    Vec3& operator/=(double s) { double k = 1.0/s; v[0] *= k; v[1] *= k; v[2] *= k; return *this; }
    double dot(const Vec3& o) const { return v[0]*o.v[0] + v[1]*o.v[1] + v[2]*o.v[2]; }
    Vec3 cross(const Vec3& o) const {
        return Vec3(v[1]*o.v[2] - v[2]*o.v[1], v[2]*o.v[0] - v[0]*o.v[2], v[0]*o.v[1] - v[1]*o.v[0]);
    }
private:
    double v[3];
};

static inline Vec3 operator*(double s, const Vec3& a) { return a*s; }

template <class CHAR, class TRAITS>
std::basic_ostream<CHAR, TRAITS>& operator<<(std::basic_ostream<CHAR, TRAITS>& o, const Vec3& a) {
    o << '[' << a[0] << ", " << a[1] << ", " << a[2] << ']';
    return o;
}

} // namespace OpenMM
#endif
