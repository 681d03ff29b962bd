// ORACLE (test infrastructure only — never linked into or called by the product path).
// Minimal fp64 3-vector with the operator semantics of the CGAL EPICK kernel types the
// reference uses (inc/cgalIncludesAndTypedefs.h:9-16): componentwise +,-; `a*b` style dot
// evaluated left-to-right; cross product in the standard component order.
// Compiled with -ffp-contract=off so no FMA contraction happens (bit-stable vs the CUDA
// kernels that are compiled with -fmad=false).
#pragma once
#include <cmath>

namespace orc {

struct V3 {
    double x, y, z;
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};

inline V3 mk(double x, double y, double z) { return V3{x, y, z}; }
inline V3 operator+(const V3& a, const V3& b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(const V3& a, const V3& b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(const V3& a) { return V3{-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, const V3& a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(const V3& a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(const V3& a, double s) { return V3{a.x / s, a.y / s, a.z / s}; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3& a, const V3& b)
{
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double sqlen(const V3& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double norm(const V3& a) { return std::sqrt(sqlen(a)); }

struct V2 {
    double x, y;
};
inline V2 operator+(const V2& a, const V2& b) { return V2{a.x + b.x, a.y + b.y}; }
inline V2 operator-(const V2& a, const V2& b) { return V2{a.x - b.x, a.y - b.y}; }
inline V2 operator*(double s, const V2& a) { return V2{s * a.x, s * a.y}; }
inline double cross2(const V2& a, const V2& b) { return a.x * b.y - a.y * b.x; }
inline double dot2(const V2& a, const V2& b) { return a.x * b.x + a.y * b.y; }
inline double norm2(const V2& a) { return std::sqrt(a.x * a.x + a.y * a.y); }

} // namespace orc
