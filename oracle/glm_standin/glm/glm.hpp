// Minimal stand-in for the GLM subset the reference rasterizer/voxelizer uses.
//
// TEST INFRASTRUCTURE ONLY.  The reference pins g-truc/glm@6f14f479 as an
// un-vendored submodule (SUB/third_party/glm is an empty directory in
// /root/reference), so the reference sources cannot be compiled without some
// header named <glm/glm.hpp>.  This file is an independent restatement of the
// published GLM semantics for the handful of types/functions the reference
// touches (census: mat3 60x, vec3 88x, vec4 38x, transpose 14x, dot 9x,
// length 6x, max 1x):
//   * vec3 / vec4: plain float aggregates, component-wise arithmetic.
//   * mat3: COLUMN-major; mat3(a,b,c, d,e,f, g,h,i) makes columns (a,b,c),
//     (d,e,f), (g,h,i); m[c] is column c, m[c][r] row r of column c;
//     mat3(s) = s * identity.
//   * mat3 * mat3: Result[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]
//     accumulated left to right (matches GLM's type_mat3x3.inl expression, so
//     nvcc's FMA contraction sees the same expression tree).
//   * dot(vec3): x*x' + y*y' + z*z' left to right; length = sqrt(dot(v,v)).
// It is used only by oracle/build_ref.sh to build oracle/_ref/ from the
// reference sources where they lie; nothing in the product path includes it.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define GLM_SI_FN __host__ __device__ inline
#else
#define GLM_SI_FN inline
#endif

namespace glm {

struct vec3 {
    float x, y, z;
    GLM_SI_FN vec3() : x(0.f), y(0.f), z(0.f) {}
    GLM_SI_FN vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    GLM_SI_FN explicit vec3(float s) : x(s), y(s), z(s) {}
    GLM_SI_FN float& operator[](int i) { return (&x)[i]; }
    GLM_SI_FN const float& operator[](int i) const { return (&x)[i]; }
    GLM_SI_FN vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    GLM_SI_FN vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
    GLM_SI_FN vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    GLM_SI_FN vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
    GLM_SI_FN vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};

struct vec4 {
    float x, y, z, w;
    GLM_SI_FN vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    GLM_SI_FN vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    GLM_SI_FN float& operator[](int i) { return (&x)[i]; }
    GLM_SI_FN const float& operator[](int i) const { return (&x)[i]; }
};

GLM_SI_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLM_SI_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLM_SI_FN vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLM_SI_FN vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLM_SI_FN vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLM_SI_FN vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
GLM_SI_FN vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
GLM_SI_FN vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
GLM_SI_FN vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }

GLM_SI_FN float dot(const vec3& a, const vec3& b) {
    vec3 t = a * b;
    return t.x + t.y + t.z;
}
GLM_SI_FN float length(const vec3& a) { return sqrtf(dot(a, a)); }
GLM_SI_FN float length(const vec4& a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w); }
GLM_SI_FN vec3 max(const vec3& a, float s) {
    return vec3(a.x < s ? s : a.x, a.y < s ? s : a.y, a.z < s ? s : a.z);
}

struct mat3 {
    vec3 c[3];
    GLM_SI_FN mat3() {}
    GLM_SI_FN explicit mat3(float s) {
        c[0] = vec3(s, 0.f, 0.f);
        c[1] = vec3(0.f, s, 0.f);
        c[2] = vec3(0.f, 0.f, s);
    }
    GLM_SI_FN mat3(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2) {
        c[0] = vec3(x0, y0, z0);
        c[1] = vec3(x1, y1, z1);
        c[2] = vec3(x2, y2, z2);
    }
    GLM_SI_FN vec3& operator[](int i) { return c[i]; }
    GLM_SI_FN const vec3& operator[](int i) const { return c[i]; }
};

GLM_SI_FN mat3 transpose(const mat3& m) {
    return mat3(m[0][0], m[1][0], m[2][0],
                m[0][1], m[1][1], m[2][1],
                m[0][2], m[1][2], m[2][2]);
}

GLM_SI_FN mat3 operator*(const mat3& A, const mat3& B) {
    mat3 R;
    for (int col = 0; col < 3; ++col)
        for (int row = 0; row < 3; ++row)
            R[col][row] = A[0][row] * B[col][0] + A[1][row] * B[col][1] + A[2][row] * B[col][2];
    return R;
}
GLM_SI_FN mat3 operator*(float s, const mat3& m) {
    mat3 R;
    R[0] = m[0] * s; R[1] = m[1] * s; R[2] = m[2] * s;
    return R;
}
GLM_SI_FN mat3 operator*(const mat3& m, float s) {
    mat3 R;
    R[0] = m[0] * s; R[1] = m[1] * s; R[2] = m[2] * s;
    return R;
}
GLM_SI_FN vec3 operator*(const mat3& m, const vec3& v) {
    return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
                m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
                m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}

}  // namespace glm
