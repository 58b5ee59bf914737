// glsl_shim.h — just enough of the GLSL 4.50 language surface, as C++20, to compile the REFERENCE'S OWN SHADER SOURCES
// (read from /root/reference/shaders where they lie, passed through oracle/ref_rig/glsl2cpp.py, never copied into the
// repository) into oracle/_ref/libvct_glsl_ref.so.  TEST INFRASTRUCTURE: it exists so that the CPU oracle's restatement
// of each shader can be checked, invocation by invocation and bit for bit, against the shader text the reference ships.
//
// What comes from the reference: every expression, branch, constant and call order of the shader bodies.
// What this file defines (the part OpenGL leaves to the implementation, fixed to this repository's canonical choices,
// DESIGN.md §2 "Canonical GL semantics"): IEEE fp32 evaluation without contraction (the build uses -ffp-contract=off);
// built-ins with the GLSL specification's formulas evaluated left to right (dot = (x*x' + y*y') + z*z', normalize = v /
// sqrt(dot), mix = a*(1-t) + b*t, clamp = min(max()), smoothstep, reflect = I - 2*dot(N,I)*N, ...); unorm8 image loads
// (b / 255) and stores (round to nearest even of clamp(v,0,1)*255); out-of-range image accesses return 0 / are dropped;
// ivec(vec) truncation; texture filtering: 2D material textures = LINEAR_MIPMAP_NEAREST min / NEAREST mag / REPEAT with
// the driver-supplied footprint, shadow map = LINEAR / CLAMP_TO_BORDER(1), 3D volumes = LINEAR_MIPMAP_LINEAR min /
// NEAREST mag / CLAMP_TO_BORDER(0), warp map = LINEAR / CLAMP_TO_EDGE — the sampler states the reference sets
// (src/Application.cpp:45-53, 383-389, 1094-1098; src/Graphics/GLHelper.cpp:180-189).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>

namespace glsl {

typedef unsigned int uint;

// ------------------------------------------------------------------------------------------------ vectors + swizzles
template <class V, class T, int M, int N, int A, int B = 0, int C = 0, int D = 0>
struct Swz {
    T d[M];
    operator V() const { V r; const int ix[4] = {A, B, C, D}; for (int i = 0; i < N; ++i) r[i] = d[ix[i]]; return r; }
    Swz& operator=(const V& v) { const int ix[4] = {A, B, C, D}; for (int i = 0; i < N; ++i) d[ix[i]] = v[i]; return *this; }
    Swz& operator=(const Swz& o) { return *this = V(o); }
    Swz& operator+=(const V& v) { return *this = V(*this) + v; }
    Swz& operator-=(const V& v) { return *this = V(*this) - v; }
    Swz& operator*=(const V& v) { return *this = V(*this) * v; }
    Swz& operator/=(const V& v) { return *this = V(*this) / v; }
    Swz& operator*=(T s) { return *this = V(*this) * s; }
    Swz& operator/=(T s) { return *this = V(*this) / s; }
    T operator[](int i) const { const int ix[4] = {A, B, C, D}; return d[ix[i]]; }
    // GLSL converts an integer vector to the float vector of the same size implicitly: vec3 v = ivec4_value.xyz;
    template <class F, class = std::enable_if_t<std::is_integral_v<T> && std::is_same_v<F, typename V::float_type>>>
    operator F() const { return F(V(*this)); }
};

template <class T> struct tvec2;
template <class T> struct tvec3;
template <class T> struct tvec4;

template <class T> struct tvec2 {
    typedef tvec2<float> float_type;
    union { T d[2]; struct { T x, y; }; struct { T r, g; }; struct { T s, t; }; Swz<tvec2<T>, T, 2, 2, 0, 1> xy; };
    tvec2() : d{T(0), T(0)} {}
    tvec2(const tvec2& o) : d{o.d[0], o.d[1]} {}
    tvec2& operator=(const tvec2& o) { d[0] = o.d[0]; d[1] = o.d[1]; return *this; }
    template <class S, class = std::enable_if_t<std::is_arithmetic_v<S>>> explicit tvec2(S v) : d{T(v), T(v)} {}
    template <class S1, class S2, class = std::enable_if_t<std::is_arithmetic_v<S1> && std::is_arithmetic_v<S2>>> tvec2(S1 a, S2 b) : d{T(a), T(b)} {}
    template <class U, class = std::enable_if_t<!std::is_same_v<U, T>>> tvec2(const tvec2<U>& o, std::enable_if_t<std::is_floating_point_v<T> && std::is_integral_v<U>, int> = 0) : d{T(o.d[0]), T(o.d[1])} {}          // ivec -> vec: implicit
    template <class U, class = std::enable_if_t<!std::is_same_v<U, T>>> explicit tvec2(const tvec2<U>& o, std::enable_if_t<!(std::is_floating_point_v<T> && std::is_integral_v<U>), long> = 0) : d{T(o.d[0]), T(o.d[1])} {}   // vec -> ivec: explicit, truncates
    T& operator[](int i) { return d[i]; }
    T operator[](int i) const { return d[i]; }
};
template <class T> struct tvec3 {
    typedef tvec3<float> float_type;
    union { T d[3]; struct { T x, y, z; }; struct { T r, g, b; };
            Swz<tvec2<T>, T, 3, 2, 0, 1> xy; Swz<tvec2<T>, T, 3, 2, 0, 2> xz; Swz<tvec2<T>, T, 3, 2, 1, 2> yz;
            Swz<tvec3<T>, T, 3, 3, 0, 1, 2> xyz, rgb; };
    tvec3() : d{T(0), T(0), T(0)} {}
    tvec3(const tvec3& o) : d{o.d[0], o.d[1], o.d[2]} {}
    tvec3& operator=(const tvec3& o) { d[0] = o.d[0]; d[1] = o.d[1]; d[2] = o.d[2]; return *this; }
    template <class S, class = std::enable_if_t<std::is_arithmetic_v<S>>> explicit tvec3(S v) : d{T(v), T(v), T(v)} {}
    template <class S1, class S2, class S3, class = std::enable_if_t<std::is_arithmetic_v<S1> && std::is_arithmetic_v<S2> && std::is_arithmetic_v<S3>>> tvec3(S1 a, S2 b, S3 c) : d{T(a), T(b), T(c)} {}
    template <class S, class = std::enable_if_t<std::is_arithmetic_v<S>>> tvec3(const tvec2<T>& a, S c) : d{a.d[0], a.d[1], T(c)} {}
    explicit tvec3(const tvec4<T>& a);                                     // vec3(vec4): drops w
    template <class U, class = std::enable_if_t<!std::is_same_v<U, T>>> tvec3(const tvec3<U>& o, std::enable_if_t<std::is_floating_point_v<T> && std::is_integral_v<U>, int> = 0) : d{T(o.d[0]), T(o.d[1]), T(o.d[2])} {}
    template <class U, class = std::enable_if_t<!std::is_same_v<U, T>>> explicit tvec3(const tvec3<U>& o, std::enable_if_t<!(std::is_floating_point_v<T> && std::is_integral_v<U>), long> = 0) : d{T(o.d[0]), T(o.d[1]), T(o.d[2])} {}
    T& operator[](int i) { return d[i]; }
    T operator[](int i) const { return d[i]; }
};
template <class T> struct tvec4 {
    typedef tvec4<float> float_type;
    union { T d[4]; struct { T x, y, z, w; }; struct { T r, g, b, a; };
            Swz<tvec2<T>, T, 4, 2, 0, 1> xy; Swz<tvec3<T>, T, 4, 3, 0, 1, 2> xyz, rgb; Swz<tvec4<T>, T, 4, 4, 0, 1, 2, 3> xyzw, rgba; };
    tvec4() : d{T(0), T(0), T(0), T(0)} {}
    tvec4(const tvec4& o) : d{o.d[0], o.d[1], o.d[2], o.d[3]} {}
    tvec4& operator=(const tvec4& o) { for (int i = 0; i < 4; ++i) d[i] = o.d[i]; return *this; }
    template <class S, class = std::enable_if_t<std::is_arithmetic_v<S>>> explicit tvec4(S v) : d{T(v), T(v), T(v), T(v)} {}
    template <class S1, class S2, class S3, class S4, class = std::enable_if_t<std::is_arithmetic_v<S1> && std::is_arithmetic_v<S2> && std::is_arithmetic_v<S3> && std::is_arithmetic_v<S4>>>
    tvec4(S1 a, S2 b, S3 c, S4 e) : d{T(a), T(b), T(c), T(e)} {}
    template <class U, class = std::enable_if_t<!std::is_same_v<U, T>>> tvec4(const tvec4<U>& o, std::enable_if_t<std::is_floating_point_v<T> && std::is_integral_v<U>, int> = 0) : d{T(o.d[0]), T(o.d[1]), T(o.d[2]), T(o.d[3])} {}
    template <class S, class = std::enable_if_t<std::is_arithmetic_v<S>>> tvec4(const tvec3<T>& a, S e) : d{a.d[0], a.d[1], a.d[2], T(e)} {}
    template <class S1, class S2, class = std::enable_if_t<std::is_arithmetic_v<S1> && std::is_arithmetic_v<S2>>> tvec4(const tvec2<T>& a, S1 c, S2 e) : d{a.d[0], a.d[1], T(c), T(e)} {}
    T& operator[](int i) { return d[i]; }
    T operator[](int i) const { return d[i]; }
};
template <class T> inline tvec3<T>::tvec3(const tvec4<T>& a) : d{a.d[0], a.d[1], a.d[2]} {}
typedef tvec2<float> vec2; typedef tvec3<float> vec3; typedef tvec4<float> vec4;
typedef tvec2<int> ivec2; typedef tvec3<int> ivec3; typedef tvec4<int> ivec4;
typedef tvec2<uint> uvec2; typedef tvec3<uint> uvec3; typedef tvec4<uint> uvec4;
typedef tvec2<bool> bvec2; typedef tvec3<bool> bvec3; typedef tvec4<bool> bvec4;

// component-wise operators: (V, V), (V, scalar), (scalar, V) for float and int vectors; mixed int-vector / float cases the
// shaders use are added explicitly (GLSL converts the integer operand to float)
#define GLSL_VEC_OPS(V, T, N)                                                                                                  \
    inline V operator+(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + b[i]; return r; }                \
    inline V operator-(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - b[i]; return r; }                \
    inline V operator*(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * b[i]; return r; }                \
    inline V operator/(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / b[i]; return r; }                \
    inline V operator+(const V& a, T s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] + s; return r; }                          \
    inline V operator-(const V& a, T s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] - s; return r; }                          \
    inline V operator*(const V& a, T s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] * s; return r; }                          \
    inline V operator/(const V& a, T s) { V r; for (int i = 0; i < N; ++i) r[i] = a[i] / s; return r; }                          \
    inline V operator+(T s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s + a[i]; return r; }                          \
    inline V operator-(T s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s - a[i]; return r; }                          \
    inline V operator*(T s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s * a[i]; return r; }                          \
    inline V operator/(T s, const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = s / a[i]; return r; }                          \
    inline V operator-(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; }                                  \
    inline V& operator+=(V& a, const V& b) { return a = a + b; }                                                                 \
    inline V& operator-=(V& a, const V& b) { return a = a - b; }                                                                 \
    inline V& operator*=(V& a, const V& b) { return a = a * b; }                                                                 \
    inline V& operator/=(V& a, const V& b) { return a = a / b; }                                                                 \
    inline V& operator+=(V& a, T s) { return a = a + s; }                                                                        \
    inline V& operator-=(V& a, T s) { return a = a - s; }                                                                        \
    inline V& operator*=(V& a, T s) { return a = a * s; }                                                                        \
    inline V& operator/=(V& a, T s) { return a = a / s; }
GLSL_VEC_OPS(vec2, float, 2) GLSL_VEC_OPS(vec3, float, 3) GLSL_VEC_OPS(vec4, float, 4)
GLSL_VEC_OPS(ivec2, int, 2) GLSL_VEC_OPS(ivec3, int, 3) GLSL_VEC_OPS(ivec4, int, 4)
#define GLSL_MIXED_OPS(V, IV, N)                                                                                    \
    inline V operator*(float s, const IV& a) { return s * V(a); }                                                    \
    inline V operator*(const IV& a, float s) { return V(a) * s; }                                                    \
    inline V operator/(float s, const IV& a) { return s / V(a); }                                                    \
    inline V operator/(const IV& a, float s) { return V(a) / s; }                                                    \
    inline V operator+(const V& a, const IV& b) { return a + V(b); }                                                 \
    inline V operator-(const V& a, const IV& b) { return a - V(b); }                                                 \
    inline V operator*(const V& a, const IV& b) { return a * V(b); }                                                 \
    inline V operator+(const IV& a, const V& b) { return V(a) + b; }                                                 \
    inline V operator-(const IV& a, const V& b) { return V(a) - b; }                                                 \
    inline V operator*(const IV& a, const V& b) { return V(a) * b; }
GLSL_MIXED_OPS(vec2, ivec2, 2) GLSL_MIXED_OPS(vec3, ivec3, 3)
// double literals never appear (glsl2cpp.py suffixes them), but an int scalar with a float vector does: 2 * normal - 1
#define GLSL_INT_SCALAR_OPS(V)                                                                        \
    inline V operator*(int s, const V& a) { return (float)s * a; }                                     \
    inline V operator*(const V& a, int s) { return a * (float)s; }                                     \
    inline V operator/(const V& a, int s) { return a / (float)s; }                                     \
    inline V operator/(int s, const V& a) { return (float)s / a; }                                     \
    inline V operator+(const V& a, int s) { return a + (float)s; }                                     \
    inline V operator+(int s, const V& a) { return (float)s + a; }                                     \
    inline V operator-(const V& a, int s) { return a - (float)s; }                                     \
    inline V operator-(int s, const V& a) { return (float)s - a; }
GLSL_INT_SCALAR_OPS(vec2) GLSL_INT_SCALAR_OPS(vec3) GLSL_INT_SCALAR_OPS(vec4)

// ---------------------------------------------------------------------------------------------------------- matrices
struct mat3 {
    vec3 c[3];
    mat3() { c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1); }
    mat3(const vec3& a, const vec3& b, const vec3& e) { c[0] = a; c[1] = b; c[2] = e; }
    explicit mat3(const struct mat4& m);                                   // upper-left 3x3
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
struct mat4 {
    vec4 c[4];
    mat4() { for (int i = 0; i < 4; ++i) { c[i] = vec4(0); c[i][i] = 1.0f; } }
    explicit mat4(const float* m) { for (int i = 0; i < 4; ++i) c[i] = vec4(m[4 * i], m[4 * i + 1], m[4 * i + 2], m[4 * i + 3]); }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline mat3::mat3(const mat4& m) { for (int i = 0; i < 3; ++i) c[i] = vec3(m.c[i].x, m.c[i].y, m.c[i].z); }
// M * v: sum over columns, left to right
inline vec3 operator*(const mat3& m, const vec3& v) { return (m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z; }
inline vec4 operator*(const mat4& m, const vec4& v) { return ((m.c[0] * v.x + m.c[1] * v.y) + m.c[2] * v.z) + m.c[3] * v.w; }
inline mat4 operator*(const mat4& a, const mat4& b) { mat4 r; for (int i = 0; i < 4; ++i) r.c[i] = a * b.c[i]; return r; }
inline mat4 transpose(const mat4& m) { mat4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.c[i][j] = m.c[j][i]; return r; }
// inverse(mat4): every use in the reference is mat3(transpose(inverse(model))) with an affine model matrix (last row 0 0 0 1).
// For those the upper-left 3x3 of the inverse is adj(A)/det(A) of the upper-left block A (canonical evaluation: cofactors as
// differences of two products, det = (a*c00 + b*c01) + c*c02, one division per entry) and the last column is -A^-1 t; a
// non-affine matrix falls back to the general cofactor expansion.
inline mat4 inverse(const mat4& M) {
    mat4 R;
    if (M.c[0].w == 0.0f && M.c[1].w == 0.0f && M.c[2].w == 0.0f && M.c[3].w == 1.0f) {
        const float a = M.c[0].x, b = M.c[1].x, c = M.c[2].x, d = M.c[0].y, e = M.c[1].y, f = M.c[2].y, g = M.c[0].z, h = M.c[1].z, i = M.c[2].z;
        const float c00 = e * i - f * h, c01 = f * g - d * i, c02 = d * h - e * g;
        const float c10 = c * h - b * i, c11 = a * i - c * g, c12 = b * g - a * h;
        const float c20 = b * f - c * e, c21 = c * d - a * f, c22 = a * e - b * d;
        const float det = (a * c00 + b * c01) + c * c02;
        // inverse = transpose(cofactor)/det: entry (row r, col k) = cofactor(k, r)/det; columns of R:
        R.c[0] = vec4(c00 / det, c01 / det, c02 / det, 0.0f);
        R.c[1] = vec4(c10 / det, c11 / det, c12 / det, 0.0f);
        R.c[2] = vec4(c20 / det, c21 / det, c22 / det, 0.0f);
        const vec3 t(M.c[3].x, M.c[3].y, M.c[3].z);
        const vec3 it = vec3(R.c[0].x, R.c[0].y, R.c[0].z) * t.x + vec3(R.c[1].x, R.c[1].y, R.c[1].z) * t.y + vec3(R.c[2].x, R.c[2].y, R.c[2].z) * t.z;
        R.c[3] = vec4(-it.x, -it.y, -it.z, 1.0f);
        return R;
    }
    const float* m = &M.c[0].x; float inv[16];                 // columns are contiguous: m[4*col + row]
    float mm[16]; for (int k = 0; k < 4; ++k) { mm[4 * k] = M.c[k].x; mm[4 * k + 1] = M.c[k].y; mm[4 * k + 2] = M.c[k].z; mm[4 * k + 3] = M.c[k].w; }
    m = mm;
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    for (int k = 0; k < 4; ++k) R.c[k] = vec4(inv[4 * k] / det, inv[4 * k + 1] / det, inv[4 * k + 2] / det, inv[4 * k + 3] / det);
    return R;
}
inline mat3 transpose(const mat3& m) { return mat3(vec3(m[0].x, m[1].x, m[2].x), vec3(m[0].y, m[1].y, m[2].y), vec3(m[0].z, m[1].z, m[2].z)); }

// ---------------------------------------------------------------------------------------------------------- built-ins
template <class A, class B> using arith2 = std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B>, std::conditional_t<std::is_integral_v<A> && std::is_integral_v<B>, int, float>>;
template <class A, class B> inline arith2<A, B> max(A a, B b) { typedef arith2<A, B> R; const R x = (R)a, y = (R)b; return x < y ? y : x; }
template <class A, class B> inline arith2<A, B> min(A a, B b) { typedef arith2<A, B> R; const R x = (R)a, y = (R)b; return y < x ? y : x; }
template <class A, class B, class C> inline std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B> && std::is_arithmetic_v<C>, float> clamp(A x, B lo, C hi) { return min(max((float)x, (float)lo), (float)hi); }
template <class A, class B, class C> inline std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B> && std::is_arithmetic_v<C>, float> mix(A a, B b, C t) { return (float)a * (1.0f - (float)t) + (float)b * (float)t; }
template <class A, class B> inline std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B>, float> step(A edge, B x) { return (float)x < (float)edge ? 0.0f : 1.0f; }
template <class A, class B> inline std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B>, float> pow(A x, B y) { return std::pow((float)x, (float)y); }
template <class A, class B, class C> inline std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B> && std::is_arithmetic_v<C>, float> smoothstep(A e0, B e1, C x) {
    const float t = clamp(((float)x - (float)e0) / ((float)e1 - (float)e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline float sqrt(float x) { return std::sqrt(x); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline float tan(float x) { return std::tan(x); }
inline float sin(float x) { return std::sin(x); }
inline float cos(float x) { return std::cos(x); }
inline float log2(float x) { return std::log2(x); }
inline float exp2(float x) { return std::exp2(x); }
inline float floor(float x) { return std::floor(x); }
inline float ceil(float x) { return std::ceil(x); }
inline float fract(float x) { return x - std::floor(x); }
inline float modf(float x, float& ip) { ip = std::trunc(x); return x - ip; }          // GLSL modf: fractional part, integer part through the out parameter
inline float abs(float x) { return std::fabs(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

#define GLSL_VEC_FUNCS(V, N)                                                                                                    \
    inline V max(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = max(a[i], b[i]); return r; }                   \
    inline V min(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = min(a[i], b[i]); return r; }                   \
    inline V max(const V& a, float b) { V r; for (int i = 0; i < N; ++i) r[i] = max(a[i], b); return r; }                         \
    inline V min(const V& a, float b) { V r; for (int i = 0; i < N; ++i) r[i] = min(a[i], b); return r; }                         \
    inline V clamp(const V& a, float lo, float hi) { V r; for (int i = 0; i < N; ++i) r[i] = clamp(a[i], lo, hi); return r; }     \
    inline V mix(const V& a, const V& b, float t) { V r; for (int i = 0; i < N; ++i) r[i] = mix(a[i], b[i], t); return r; }       \
    inline V mix(const V& a, const V& b, const V& t) { V r; for (int i = 0; i < N; ++i) r[i] = mix(a[i], b[i], t[i]); return r; } \
    inline V step(const V& e, const V& x) { V r; for (int i = 0; i < N; ++i) r[i] = step(e[i], x[i]); return r; }                 \
    inline V step(float e, const V& x) { V r; for (int i = 0; i < N; ++i) r[i] = step(e, x[i]); return r; }                       \
    inline V pow(const V& a, const V& b) { V r; for (int i = 0; i < N; ++i) r[i] = pow(a[i], b[i]); return r; }                   \
    inline V floor(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = floor(a[i]); return r; }                                 \
    inline V ceil(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = ceil(a[i]); return r; }                                   \
    inline V fract(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = fract(a[i]); return r; }                                 \
    inline V abs(const V& a) { V r; for (int i = 0; i < N; ++i) r[i] = abs(a[i]); return r; }                                     \
    inline V modf(const V& a, V& ip) { V r; for (int i = 0; i < N; ++i) { ip[i] = std::trunc(a[i]); r[i] = a[i] - ip[i]; } return r; } \
    inline float dot(const V& a, const V& b) { float s = a[0] * b[0]; for (int i = 1; i < N; ++i) s = s + a[i] * b[i]; return s; } \
    inline float length(const V& a) { return std::sqrt(dot(a, a)); }                                                             \
    inline float distance(const V& a, const V& b) { return length(a - b); }                                                      \
    inline V normalize(const V& a) { const float l = std::sqrt(dot(a, a)); V r; for (int i = 0; i < N; ++i) r[i] = a[i] / l; return r; } \
    inline V reflect(const V& I, const V& Nn) { return I - 2.0f * dot(Nn, I) * Nn; }
GLSL_VEC_FUNCS(vec2, 2) GLSL_VEC_FUNCS(vec3, 3) GLSL_VEC_FUNCS(vec4, 4)
inline vec3 mix(const vec3& a, const vec3& b, const tvec3<bool>& sel) { return vec3(sel[0] ? b.x : a.x, sel[1] ? b.y : a.y, sel[2] ? b.z : a.z); }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }

#define GLSL_REL(V, BV, N)                                                                                                      \
    inline BV greaterThan(const V& a, const V& b) { BV r; for (int i = 0; i < N; ++i) r[i] = a[i] > b[i]; return r; }             \
    inline BV greaterThanEqual(const V& a, const V& b) { BV r; for (int i = 0; i < N; ++i) r[i] = a[i] >= b[i]; return r; }       \
    inline BV lessThan(const V& a, const V& b) { BV r; for (int i = 0; i < N; ++i) r[i] = a[i] < b[i]; return r; }                \
    inline BV lessThanEqual(const V& a, const V& b) { BV r; for (int i = 0; i < N; ++i) r[i] = a[i] <= b[i]; return r; }          \
    inline BV equal(const V& a, const V& b) { BV r; for (int i = 0; i < N; ++i) r[i] = a[i] == b[i]; return r; }                  \
    inline BV notEqual(const V& a, const V& b) { BV r; for (int i = 0; i < N; ++i) r[i] = a[i] != b[i]; return r; }
GLSL_REL(vec2, bvec2, 2) GLSL_REL(vec3, bvec3, 3) GLSL_REL(vec4, bvec4, 4) GLSL_REL(ivec2, bvec2, 2) GLSL_REL(ivec3, bvec3, 3)
inline bool any(const bvec2& b) { return b[0] || b[1]; }
inline bool any(const bvec3& b) { return b[0] || b[1] || b[2]; }
inline bool any(const bvec4& b) { return b[0] || b[1] || b[2] || b[3]; }
inline bool all(const bvec3& b) { return b[0] && b[1] && b[2]; }

template <class T, size_t N> inline int glsl_length(const T (&)[N]) { return (int)N; }

// ------------------------------------------------------------------------------------------- unorm8 / fp16 conversion
inline float unorm8_to_float(uint b) { return (float)b / 255.0f; }
inline uint float_to_unorm8(float v) {                         // clamp, scale, round to nearest even; NaN -> 0
    if (!(v > 0.0f)) return 0u;
    if (v >= 1.0f) return 255u;
    return (uint)std::nearbyintf(v * 255.0f);
}
inline uint packUnorm4x8(const vec4& c) { return float_to_unorm8(c.x) | float_to_unorm8(c.y) << 8 | float_to_unorm8(c.z) << 16 | float_to_unorm8(c.w) << 24; }
inline vec4 unpackUnorm4x8(uint w) { return vec4(unorm8_to_float(w & 255u), unorm8_to_float((w >> 8) & 255u), unorm8_to_float((w >> 16) & 255u), unorm8_to_float(w >> 24)); }
inline float half_to_float(uint16_t h) {
    const uint32_t s = (uint32_t)(h & 0x8000u) << 16; const int e = (h >> 10) & 31; const uint32_t m = h & 0x3FFu;
    uint32_t bits;
    if (e == 0) { if (m == 0) bits = s; else { int k = 0; uint32_t mm = m; while (!(mm & 0x400u)) { mm <<= 1; ++k; } bits = s | (uint32_t)(113 - k) << 23 | (mm & 0x3FFu) << 13; } }
    else if (e == 31) bits = s | 0x7F800000u | m << 13;
    else bits = s | (uint32_t)(e + 112) << 23 | m << 13;
    float f; std::memcpy(&f, &bits, 4); return f;
}
inline uint16_t float_to_half(float f) {                       // round to nearest even
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t s = (x >> 16) & 0x8000u; const int e = (int)((x >> 23) & 255u) - 127 + 15; uint32_t m = x & 0x7FFFFFu;
    if (((x >> 23) & 255u) == 255u) return (uint16_t)(s | 0x7C00u | (m ? 0x200u : 0u));
    if (e >= 31) return (uint16_t)(s | 0x7C00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)s;
        m |= 0x800000u;
        const int shift = 14 - e; uint32_t h = m >> shift; const uint32_t rem = m & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1u))) ++h;
        return (uint16_t)(s | h);
    }
    uint32_t h = (uint32_t)e << 10 | m >> 13; const uint32_t rem = m & 0x1FFFu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
    return (uint16_t)(s | h);
}

// ------------------------------------------------------------------------------------------------------------- images
enum ImageFormat { FMT_RGBA8, FMT_RGBA16F, FMT_R32UI, FMT_RGBA16 };
struct image3D {                                               // also uimage3D: one 32-bit word per texel for r32ui
    void* data = nullptr; int w = 0, h = 0, d = 0; ImageFormat fmt = FMT_RGBA8;
    bool inside(const ivec3& p) const { return p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < w && p.y < h && p.z < d; }
    size_t at(const ivec3& p) const { return ((size_t)p.z * h + p.y) * w + p.x; }
};
typedef image3D uimage3D;
inline ivec3 imageSize(const image3D& im) { return ivec3(im.w, im.h, im.d); }
inline vec4 imageLoadF(const image3D& im, const ivec3& p) {
    if (!im.inside(p)) return vec4(0);
    if (im.fmt == FMT_RGBA8) return unpackUnorm4x8(((const uint32_t*)im.data)[im.at(p)]);
    if (im.fmt == FMT_RGBA16F) { const uint16_t* h = (const uint16_t*)im.data + 4 * im.at(p); return vec4(half_to_float(h[0]), half_to_float(h[1]), half_to_float(h[2]), half_to_float(h[3])); }
    return vec4(0);
}
inline vec4 imageLoad(const image3D& im, const ivec3& p) { return imageLoadF(im, p); }
// imageLoad on a uimage3D (r32ui view): glsl2cpp.py routes those call sites here
inline uvec4 imageLoadU(const image3D& im, const ivec3& p) { return uvec4(im.inside(p) ? ((const uint32_t*)im.data)[im.at(p)] : 0u, 0u, 0u, 1u); }
// rgba32i volume (warpPartials; uploaded as GL_RGB_INTEGER, so alpha reads 1) and r32f 2D image (warpWeights)
struct iimage3D { const int* data = nullptr; int w = 0, h = 0, d = 0, comps = 3; };
inline ivec4 imageLoadI(const iimage3D& im, const ivec3& p) {
    if (p.x < 0 || p.y < 0 || p.z < 0 || p.x >= im.w || p.y >= im.h || p.z >= im.d) return ivec4(0, 0, 0, 0);
    const int* q = im.data + (((size_t)p.z * im.h + p.y) * im.w + p.x) * im.comps;
    return ivec4(q[0], q[1], q[2], 1);
}
struct image2D { const float* data = nullptr; int w = 0, h = 0; };
inline vec4 imageLoad(const image2D& im, const ivec2& p) {
    if (p.x < 0 || p.y < 0 || p.x >= im.w || p.y >= im.h) return vec4(0);
    return vec4(im.data[(size_t)p.y * im.w + p.x], 0, 0, 1);
}
inline void imageStore(image3D& im, const ivec3& p, const vec4& v) {
    if (!im.inside(p)) return;
    if (im.fmt == FMT_RGBA8) ((uint32_t*)im.data)[im.at(p)] = packUnorm4x8(v);
    else if (im.fmt == FMT_RGBA16F) { uint16_t* h = (uint16_t*)im.data + 4 * im.at(p); for (int i = 0; i < 4; ++i) h[i] = float_to_half(v[i]); }
}
inline void imageStore(image3D& im, const ivec3& p, const uvec4& v) { if (im.inside(p)) ((uint32_t*)im.data)[im.at(p)] = v.x; }
inline uint imageAtomicCompSwap(image3D& im, const ivec3& p, uint compare, uint value) {
    if (!im.inside(p)) return 0u;
    uint32_t& w = ((uint32_t*)im.data)[im.at(p)]; const uint old = w; if (old == compare) w = value; return old;
}
inline uint imageAtomicMax(image3D& im, const ivec3& p, uint value) { if (!im.inside(p)) return 0u; uint32_t& w = ((uint32_t*)im.data)[im.at(p)]; const uint old = w; if (value > old) w = value; return old; }
inline uint imageAtomicOr(image3D& im, const ivec3& p, uint value) { if (!im.inside(p)) return 0u; uint32_t& w = ((uint32_t*)im.data)[im.at(p)]; const uint old = w; w |= value; return old; }
// buffer-variable atomics: real ones, the compute drivers run a dispatch on all host threads
inline uint atomicAdd(uint& mem, uint v) { return __atomic_fetch_add(&mem, v, __ATOMIC_RELAXED); }
inline uint atomicMax(uint& mem, uint v) {
    uint old = __atomic_load_n(&mem, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(&mem, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// ----------------------------------------------------------------------------------------------------------- samplers
// 2D: either a float depth map (shadow map: LINEAR, CLAMP_TO_BORDER border 1) or an 8-bit material texture with mips
// (LINEAR_MIPMAP_NEAREST min, NEAREST mag, REPEAT); the footprint rho^2 (squared texel-space length of the longer screen
// derivative, GL 4.5 §8.14.1) of the current fragment is supplied by the driver in `rho2`.
struct sampler2D {
    const float* depth = nullptr; int size = 0;                                  // shadow map
    const uint8_t* const* level = nullptr; int w = 0, h = 0, ch = 0, levels = 0;  // material texture
    float rho2 = 0.0f;
};
inline ivec2 textureSize(const sampler2D& s, int) { return s.depth ? ivec2(s.size, s.size) : ivec2(s.w, s.h); }
inline float shadow_texel(const sampler2D& s, int x, int y) { return (x < 0 || y < 0 || x >= s.size || y >= s.size) ? 1.0f : s.depth[(size_t)y * s.size + x]; }
inline vec4 texel2d(const sampler2D& s, int l, int x, int y) {
    const int w = s.w >> l ? s.w >> l : 1, h = s.h >> l ? s.h >> l : 1;
    x %= w; if (x < 0) x += w; y %= h; if (y < 0) y += h;
    const uint8_t* p = s.level[l] + ((size_t)y * w + x) * s.ch;
    vec4 r(0, 0, 0, 1);
    for (int c = 0; c < s.ch; ++c) r[c] = (float)p[c] / 255.0f;
    return r;
}
inline vec4 lerp4(const vec4& a, const vec4& b, float t) { const float s = 1.0f - t; vec4 r; for (int i = 0; i < 4; ++i) r[i] = a[i] * s + b[i] * t; return r; }
inline vec4 textureOffset(const sampler2D& s, const vec2& tc, const ivec2& off) {
    if (!s.depth && !s.level) return vec4(0, 0, 0, 1);          // no texture bound to the unit
    if (s.depth) {
        const float x = tc.x * (float)s.size - 0.5f, y = tc.y * (float)s.size - 0.5f;
        const float fx = std::floor(x), fy = std::floor(y);
        if (!(std::fabs(fx) < 1e9f) || !(std::fabs(fy) < 1e9f)) return vec4(1, 0, 0, 1);   // NaN / far outside: border
        const int x0 = (int)fx + off.x, y0 = (int)fy + off.y; const float ax = x - fx, ay = y - fy;
        const float t00 = shadow_texel(s, x0, y0), t10 = shadow_texel(s, x0 + 1, y0), t01 = shadow_texel(s, x0, y0 + 1), t11 = shadow_texel(s, x0 + 1, y0 + 1);
        const float top = t00 * (1.0f - ax) + t10 * ax, bot = t01 * (1.0f - ax) + t11 * ax;
        return vec4(top * (1.0f - ay) + bot * ay, 0, 0, 1);
    }
    // material texture, GL 4.5 §8.14: lambda = log2(rho) is compared through rho^2.  Magnification <=> lambda <= 0.5 (the
    // switch-over constant of the LINEAR_MIPMAP_NEAREST / NEAREST filter pair) <=> rho2 <= 2 -> NEAREST on level 0;
    // otherwise LINEAR on level ceil(lambda + 0.5) - 1 = the smallest l >= 1 with rho2 <= 2^(2l+1), clamped to the last level
    if (!(s.rho2 > 2.0f)) {
        const int x = (int)std::floor(tc.x * (float)s.w), y = (int)std::floor(tc.y * (float)s.h);
        return texel2d(s, 0, x + off.x, y + off.y);
    }
    int l = 1; float lim = 8.0f;
    while (l < s.levels - 1 && s.rho2 > lim) { ++l; lim *= 4.0f; }
    if (l > s.levels - 1) l = s.levels - 1;
    const int w = s.w >> l ? s.w >> l : 1, h = s.h >> l ? s.h >> l : 1;
    const float x = tc.x * (float)w - 0.5f, y = tc.y * (float)h - 0.5f, fx = std::floor(x), fy = std::floor(y);
    const int x0 = (int)fx + off.x, y0 = (int)fy + off.y; const float ax = x - fx, ay = y - fy;
    return lerp4(lerp4(texel2d(s, l, x0, y0), texel2d(s, l, x0 + 1, y0), ax), lerp4(texel2d(s, l, x0, y0 + 1), texel2d(s, l, x0 + 1, y0 + 1), ax), ay);
}
inline vec4 texture(const sampler2D& s, const vec2& tc) { return textureOffset(s, tc, ivec2(0, 0)); }

// 3D: RGBA8 mip pyramid (voxel volumes) or RGBA16 unorm single level (warp map)
struct sampler3D {
    const uint32_t* const* level = nullptr; int dim = 0, levels = 0;              // voxel pyramid
    const uint16_t* warp = nullptr; int wdim = 0;                                // warp map
    const uint16_t* f16 = nullptr; int fdim = 0;                                  // RGBA16F render target sampled NEAREST / CLAMP_TO_EDGE
    unsigned long long* fetches = nullptr;
};
inline ivec3 textureSize(const sampler3D& s, int) { return s.warp ? ivec3(s.wdim) : ivec3(s.dim); }
inline vec4 vol_texel(const sampler3D& s, int l, int x, int y, int z) {
    const int d = s.dim >> l ? s.dim >> l : 1;
    if (x < 0 || y < 0 || z < 0 || x >= d || y >= d || z >= d) return vec4(0);
    return unpackUnorm4x8(s.level[l][((size_t)z * d + y) * d + x]);
}
inline vec4 vol_linear(const sampler3D& s, int l, const vec3& tc) {
    const int d = s.dim >> l ? s.dim >> l : 1;
    const float x = tc.x * (float)d - 0.5f, y = tc.y * (float)d - 0.5f, z = tc.z * (float)d - 0.5f;
    const float fx = std::floor(x), fy = std::floor(y), fz = std::floor(z);
    const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz; const float ax = x - fx, ay = y - fy, az = z - fz;
    vec4 plane[2];
    for (int k = 0; k < 2; ++k)
        plane[k] = lerp4(lerp4(vol_texel(s, l, x0, y0, z0 + k), vol_texel(s, l, x0 + 1, y0, z0 + k), ax),
                         lerp4(vol_texel(s, l, x0, y0 + 1, z0 + k), vol_texel(s, l, x0 + 1, y0 + 1, z0 + k), ax), ay);
    return lerp4(plane[0], plane[1], az);
}
inline vec3 warp_texel(const sampler3D& s, int x, int y, int z) {
    const int n = s.wdim;
    x = x < 0 ? 0 : (x >= n ? n - 1 : x); y = y < 0 ? 0 : (y >= n ? n - 1 : y); z = z < 0 ? 0 : (z >= n ? n - 1 : z);
    const uint16_t* p = s.warp + 4 * (((size_t)z * n + y) * n + x);
    return vec3((float)p[0] / 65535.0f, (float)p[1] / 65535.0f, (float)p[2] / 65535.0f);
}
inline vec3 lerp3(const vec3& a, const vec3& b, float t) { const float s = 1.0f - t; return vec3(a.x * s + b.x * t, a.y * s + b.y * t, a.z * s + b.z * t); }
inline vec4 textureOffset(const sampler3D& s, const vec3& tc, const ivec3& off) {
    const int n = s.wdim;
    const float x = tc.x * (float)n - 0.5f, y = tc.y * (float)n - 0.5f, z = tc.z * (float)n - 0.5f;
    const float fx = std::floor(x), fy = std::floor(y), fz = std::floor(z);
    if (!(std::fabs(fx) < 1e9f) || !(std::fabs(fy) < 1e9f) || !(std::fabs(fz) < 1e9f)) return vec4(0, 0, 0, 1);
    const int x0 = (int)fx + off.x, y0 = (int)fy + off.y, z0 = (int)fz + off.z; const float ax = x - fx, ay = y - fy, az = z - fz;
    vec3 plane[2];
    for (int k = 0; k < 2; ++k)
        plane[k] = lerp3(lerp3(warp_texel(s, x0, y0, z0 + k), warp_texel(s, x0 + 1, y0, z0 + k), ax),
                         lerp3(warp_texel(s, x0, y0 + 1, z0 + k), warp_texel(s, x0 + 1, y0 + 1, z0 + k), ax), ay);
    return vec4(lerp3(plane[0], plane[1], az), 1.0f);
}
inline vec4 texture(const sampler3D& s, const vec3& tc) {
    if (s.f16) {                                                // warp-weight targets: NEAREST, CLAMP_TO_EDGE (Application.cpp:437-449)
        int t[3];
        for (int k = 0; k < 3; ++k) { t[k] = (int)std::floor(tc[k] * (float)s.fdim); t[k] = t[k] < 0 ? 0 : (t[k] >= s.fdim ? s.fdim - 1 : t[k]); }
        const uint16_t* p = s.f16 + 4 * (((size_t)t[2] * s.fdim + t[1]) * s.fdim + t[0]);
        return vec4(half_to_float(p[0]), half_to_float(p[1]), half_to_float(p[2]), half_to_float(p[3]));
    }
    return textureOffset(s, tc, ivec3(0, 0, 0));                 // warp map (level 0, LINEAR, CLAMP_TO_EDGE)
}
inline vec4 textureLod(const sampler3D& s, const vec3& tc, float lambda) {
    if (s.fetches) ++*s.fetches;
    const float top = (float)(s.levels - 1);
    if (lambda > top) lambda = top;
    if (!(lambda > 0.5f)) {                                     // magnification: NEAREST on level 0
        const int d = s.dim;
        return vol_texel(s, 0, (int)std::floor(tc.x * (float)d), (int)std::floor(tc.y * (float)d), (int)std::floor(tc.z * (float)d));
    }
    const float fl = std::floor(lambda); const int l0 = (int)fl;
    if (l0 >= s.levels - 1) return vol_linear(s, s.levels - 1, tc);
    return lerp4(vol_linear(s, l0, tc), vol_linear(s, l0 + 1, tc), lambda - fl);
}

// textureLodOffset on a voxel volume (filter3d.comp): explicit lod, integer texel offset; lod <= 0.5 magnifies -> NEAREST on level 0
inline vec4 textureLodOffset(const sampler3D& s, const vec3& tc, float lambda, const ivec3& off) {
    if (!(lambda > 0.5f)) {
        const int d = s.dim;
        return vol_texel(s, 0, (int)std::floor(tc.x * (float)d) + off.x, (int)std::floor(tc.y * (float)d) + off.y, (int)std::floor(tc.z * (float)d) + off.z);
    }
    const int l = lambda >= (float)(s.levels - 1) ? s.levels - 1 : (int)std::floor(lambda);
    const int d = s.dim >> l ? s.dim >> l : 1;
    return vol_linear(s, l, vec3(tc.x + (float)off.x / (float)d, tc.y + (float)off.y / (float)d, tc.z + (float)off.z / (float)d));
}
template <class T> struct ssbo_array { const T* data = nullptr; int n = 0; const T& operator[](int i) const { return data[i]; } int length() const { return n; } };
template <class T> inline int glsl_length(const ssbo_array<T>& a) { return a.n; }

struct Discard {};

}  // namespace glsl
