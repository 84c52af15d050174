// Header shim standing in for OpenMM's openmm/Vec3.h (OpenMM is not installed in this build environment).
// Only used to COMPILE AND EXERCISE the plugin glue in glue/; against a real OpenMM these shim headers are
// simply left out of the include path.  Same public interface as OpenMM 7.x.
#ifndef OPENMM_VEC3_H_
#define OPENMM_VEC3_H_
#include <cassert>
#include <iosfwd>
namespace OpenMM {
class Vec3 {
public:
    Vec3() { data[0] = data[1] = data[2] = 0.0; }
    Vec3(double x, double y, double z) { data[0] = x; data[1] = y; data[2] = z; }
    double operator[](int index) const { assert(index >= 0 && index < 3); return data[index]; }
    double& operator[](int index) { assert(index >= 0 && index < 3); return data[index]; }
    bool operator==(const Vec3& rhs) const { return data[0] == rhs[0] && data[1] == rhs[1] && data[2] == rhs[2]; }
    bool operator!=(const Vec3& rhs) const { return !(*this == rhs); }
    Vec3 operator+(const Vec3& rhs) const { return Vec3(data[0]+rhs[0], data[1]+rhs[1], data[2]+rhs[2]); }
    Vec3 operator-(const Vec3& rhs) const { return Vec3(data[0]-rhs[0], data[1]-rhs[1], data[2]-rhs[2]); }
    Vec3 operator-() const { return Vec3(-data[0], -data[1], -data[2]); }
    Vec3& operator+=(const Vec3& rhs) { data[0] += rhs[0]; data[1] += rhs[1]; data[2] += rhs[2]; return *this; }
    Vec3& operator-=(const Vec3& rhs) { data[0] -= rhs[0]; data[1] -= rhs[1]; data[2] -= rhs[2]; return *this; }
    Vec3 operator*(double rhs) const { return Vec3(data[0]*rhs, data[1]*rhs, data[2]*rhs); }
    Vec3& operator*=(double rhs) { data[0] *= rhs; data[1] *= rhs; data[2] *= rhs; return *this; }
    Vec3 operator/(double rhs) const { double s = 1.0/rhs; return Vec3(data[0]*s, data[1]*s, data[2]*s); }
    double dot(const Vec3& rhs) const { return data[0]*rhs[0] + data[1]*rhs[1] + data[2]*rhs[2]; }
    Vec3 cross(const Vec3& rhs) const {
        return Vec3(data[1]*rhs[2]-data[2]*rhs[1], data[2]*rhs[0]-data[0]*rhs[2], data[0]*rhs[1]-data[1]*rhs[0]);
    }
private:
    double data[3];
};
static inline Vec3 operator*(double lhs, Vec3 rhs) { return rhs*lhs; }
}
#endif
