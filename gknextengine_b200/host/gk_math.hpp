// Host-side math used by the scene/UBO code above the C ABI. The reference does this with glm
// (vcpkg, not vendored under /root/reference); these are restatements of the published glm
// algorithms the reference calls: perspective (RH, depth 0..1: src/Utilities/Glm.hpp:3-4),
// lookAt, inverse, rotate, translate, scale, mat4_cast, quat(euler) and detail::toFloat16.
// Matrices are column-major m[col][row] exactly like glm, so a GkMat4 can be memcpy'd into the UBO.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#include "../../include/gkcu.h"

namespace gk {

struct vec2 { float x = 0, y = 0; };
struct vec3 {
    float x = 0, y = 0, z = 0;
    vec3() = default;
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    explicit vec3(float a) : x(a), y(a), z(a) {}
};
struct vec4 {
    float x = 0, y = 0, z = 0, w = 0;
    vec4() = default;
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
};
struct quat { float w = 1, x = 0, y = 0, z = 0; };

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }

struct mat4 {
    float m[4][4];
    mat4() { std::memset(m, 0, sizeof(m)); }
    explicit mat4(float d) { std::memset(m, 0, sizeof(m)); m[0][0] = m[1][1] = m[2][2] = m[3][3] = d; }
    float* operator[](int c) { return m[c]; }
    const float* operator[](int c) const { return m[c]; }
    GkMat4 pod() const { GkMat4 r; std::memcpy(&r, m, sizeof(r)); return r; }
};

inline mat4 operator*(const mat4& a, const mat4& b) {
    mat4 r;
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row)
            r[c][row] = a[0][row] * b[c][0] + a[1][row] * b[c][1] + a[2][row] * b[c][2] + a[3][row] * b[c][3];
    return r;
}
inline vec4 operator*(const mat4& a, const vec4& v) {
    return {a[0][0] * v.x + a[1][0] * v.y + a[2][0] * v.z + a[3][0] * v.w,
            a[0][1] * v.x + a[1][1] * v.y + a[2][1] * v.z + a[3][1] * v.w,
            a[0][2] * v.x + a[1][2] * v.y + a[2][2] * v.z + a[3][2] * v.w,
            a[0][3] * v.x + a[1][3] * v.y + a[2][3] * v.z + a[3][3] * v.w};
}

inline mat4 translate(const mat4& m, vec3 v) {
    mat4 r = m;
    for (int i = 0; i < 4; ++i) r[3][i] = m[0][i] * v.x + m[1][i] * v.y + m[2][i] * v.z + m[3][i];
    return r;
}
inline mat4 scale(const mat4& m, vec3 v) {
    mat4 r;
    for (int i = 0; i < 4; ++i) { r[0][i] = m[0][i] * v.x; r[1][i] = m[1][i] * v.y; r[2][i] = m[2][i] * v.z; r[3][i] = m[3][i]; }
    return r;
}
inline mat4 rotate(const mat4& m, float angle, vec3 v) {
    const float c = std::cos(angle), s = std::sin(angle);
    vec3 axis = normalize(v);
    vec3 temp = axis * (1.0f - c);
    float R[3][3];
    R[0][0] = c + temp.x * axis.x; R[0][1] = temp.x * axis.y + s * axis.z; R[0][2] = temp.x * axis.z - s * axis.y;
    R[1][0] = temp.y * axis.x - s * axis.z; R[1][1] = c + temp.y * axis.y; R[1][2] = temp.y * axis.z + s * axis.x;
    R[2][0] = temp.z * axis.x + s * axis.y; R[2][1] = temp.z * axis.y - s * axis.x; R[2][2] = c + temp.z * axis.z;
    mat4 r;
    for (int i = 0; i < 4; ++i) {
        r[0][i] = m[0][i] * R[0][0] + m[1][i] * R[0][1] + m[2][i] * R[0][2];
        r[1][i] = m[0][i] * R[1][0] + m[1][i] * R[1][1] + m[2][i] * R[1][2];
        r[2][i] = m[0][i] * R[2][0] + m[1][i] * R[2][1] + m[2][i] * R[2][2];
        r[3][i] = m[3][i];
    }
    return r;
}
// glm::quat(vec3 eulerAngles) (pitch, yaw, roll), as used by SceneList.cpp:216,222
inline quat quat_from_euler(vec3 e) {
    vec3 c(std::cos(e.x * 0.5f), std::cos(e.y * 0.5f), std::cos(e.z * 0.5f));
    vec3 s(std::sin(e.x * 0.5f), std::sin(e.y * 0.5f), std::sin(e.z * 0.5f));
    quat q;
    q.w = c.x * c.y * c.z + s.x * s.y * s.z;
    q.x = s.x * c.y * c.z - c.x * s.y * s.z;
    q.y = c.x * s.y * c.z + s.x * c.y * s.z;
    q.z = c.x * c.y * s.z - s.x * s.y * c.z;
    return q;
}
inline mat4 mat4_cast(quat q) {
    mat4 r(1.0f);
    float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z, qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z,
          qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    r[0][0] = 1.0f - 2.0f * (qyy + qzz); r[0][1] = 2.0f * (qxy + qwz); r[0][2] = 2.0f * (qxz - qwy);
    r[1][0] = 2.0f * (qxy - qwz); r[1][1] = 1.0f - 2.0f * (qxx + qzz); r[1][2] = 2.0f * (qyz + qwx);
    r[2][0] = 2.0f * (qxz + qwy); r[2][1] = 2.0f * (qyz - qwx); r[2][2] = 1.0f - 2.0f * (qxx + qyy);
    return r;
}
inline mat4 perspective_rh_zo(float fovy, float aspect, float zNear, float zFar) {
    const float t = std::tan(fovy / 2.0f);
    mat4 r;
    r[0][0] = 1.0f / (aspect * t);
    r[1][1] = 1.0f / t;
    r[2][2] = zFar / (zNear - zFar);
    r[2][3] = -1.0f;
    r[3][2] = -(zFar * zNear) / (zFar - zNear);
    return r;
}
inline mat4 look_at_rh(vec3 eye, vec3 center, vec3 up) {
    vec3 f = normalize(center - eye), s = normalize(cross(f, up)), u = cross(s, f);
    mat4 r(1.0f);
    r[0][0] = s.x; r[1][0] = s.y; r[2][0] = s.z;
    r[0][1] = u.x; r[1][1] = u.y; r[2][1] = u.z;
    r[0][2] = -f.x; r[1][2] = -f.y; r[2][2] = -f.z;
    r[3][0] = -dot(s, eye); r[3][1] = -dot(u, eye); r[3][2] = dot(f, eye);
    return r;
}
// glm::inverse(mat4): cofactor expansion with shared 2x2 sub-determinants.
inline mat4 inverse(const mat4& m) {
    float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3], c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    float c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3], c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2], c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    float c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3], c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2], c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    float c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1], c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    float F0[4] = {c00, c00, c02, c03}, F1[4] = {c04, c04, c06, c07}, F2[4] = {c08, c08, c10, c11};
    float F3[4] = {c12, c12, c14, c15}, F4[4] = {c16, c16, c18, c19}, F5[4] = {c20, c20, c22, c23};
    float V0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, V1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
    float V2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, V3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
    const float SA[4] = {+1, -1, +1, -1}, SB[4] = {-1, +1, -1, +1};
    mat4 inv;
    for (int i = 0; i < 4; ++i) {
        inv[0][i] = (V1[i] * F0[i] - V2[i] * F1[i] + V3[i] * F2[i]) * SA[i];
        inv[1][i] = (V0[i] * F0[i] - V2[i] * F3[i] + V3[i] * F4[i]) * SB[i];
        inv[2][i] = (V0[i] * F1[i] - V1[i] * F3[i] + V3[i] * F5[i]) * SA[i];
        inv[3][i] = (V0[i] * F2[i] - V1[i] * F4[i] + V2[i] * F5[i]) * SB[i];
    }
    float det = (m[0][0] * inv[0][0] + m[0][1] * inv[1][0]) + (m[0][2] * inv[2][0] + m[0][3] * inv[3][0]);
    float ood = 1.0f / det;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) inv[c][r] *= ood;
    return inv;
}

// glm::detail::toFloat16 (detail/type_half.inl), used by MakeVertex (src/Assets/Vertex.hpp:80-99):
// round-half-up in magnitude on mantissa bit 12, not ties-to-even.
inline uint16_t to_float16(float f) {
    uint32_t bits;
    std::memcpy(&bits, &f, 4);
    int i = (int)bits;
    int s = (i >> 16) & 0x00008000;
    int e = ((i >> 23) & 0x000000ff) - (127 - 15);
    int m = i & 0x007fffff;
    if (e <= 0) {
        if (e < -10) return (uint16_t)s;
        m = (m | 0x00800000) >> (1 - e);
        if (m & 0x00001000) m += 0x00002000;
        return (uint16_t)(s | (m >> 13));
    } else if (e == 0xff - (127 - 15)) {
        if (m == 0) return (uint16_t)(s | 0x7c00);
        m >>= 13;
        return (uint16_t)(s | 0x7c00 | m | (m == 0));
    } else {
        if (m & 0x00001000) {
            m += 0x00002000;
            if (m & 0x00800000) { m = 0; e += 1; }
        }
        if (e > 30) return (uint16_t)(s | 0x7c00);
        return (uint16_t)(s | (e << 10) | (m >> 13));
    }
}

}  // namespace gk
