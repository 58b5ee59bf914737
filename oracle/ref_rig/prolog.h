// prolog.h — scaffolding that lets the reference's OWN warp example (the `#if 0` block of
// /root/reference/src/main.cpp:21-128) compile and run here without GLM/GLFW/OpenGL.  Nothing of the reference is
// stored in this repository: oracle/Makefile pipes those lines from where they lie between this prolog and
// epilog.inc into g++, and only the resulting binary lands in oracle/_ref/ (git-ignored).
//
// The block uses five things from GLM 0.9.9.5 (vcpkg pin): vec2 with component-wise float
// arithmetic, vec2(scalar), modf(vec2, vec2&) and to_string(vec2).  The shim below restates exactly those
// (GLM's vec2 operators are plain per-component binary32 operations; glm::modf is std::modf per component).
#include <cmath>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <string>

namespace glm {
struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    template <class A> explicit vec2(A s) : x((float)s), y((float)s) {}
    template <class A, class B> vec2(A a, B b) : x((float)a), y((float)b) {}
    vec2& operator*=(float s) { x *= s; y *= s; return *this; }
};
inline vec2 operator+(vec2 a, vec2 b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator-(vec2 a, vec2 b) { return vec2(a.x - b.x, a.y - b.y); }
inline vec2 operator*(vec2 a, vec2 b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator/(vec2 a, float s) { return vec2(a.x / s, a.y / s); }
inline vec2 modf(vec2 v, vec2& i) { float ix, iy; const float fx = std::modf(v.x, &ix), fy = std::modf(v.y, &iy); i = vec2(ix, iy); return vec2(fx, fy); }
inline std::string to_string(vec2 v) { std::ostringstream s; s << "vec2(" << v.x << ", " << v.y << ")"; return s.str(); }
}  // namespace glm

static std::ostringstream rig_log;          // the example narrates every lookup on cout: keep it out of the data stream
#define cout rig_log

int main(int argc, char** argv) {
