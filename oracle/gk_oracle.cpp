// gk_oracle.cpp — scalar CPU restatement of gkNextEngine's path-tracing frame. TEST INFRASTRUCTURE ONLY
// (see gk_oracle.h for who may use it and for the parity-pinning status).
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
// Build: g++ -O2 -std=c++17 -ffp-contract=off  (no FMA contraction: include/gk_detmath.h explains why).
// Intersector: a native BVH2 + watertight triangle test by default; with -DGKO_TINYBVH the reference's own
// vendored tinybvh (src/ThirdParty/tinybvh/tiny_bvh.h, used like src/Assets/CPUAccelerationStructure.cpp:172-284).
#include "gk_oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/gk_detmath.h"

#ifdef GKO_TINYBVH
#define TINYBVH_IMPLEMENTATION
#include "tiny_bvh.h"
#endif

namespace {

// ----------------------------------------------------------------------------------------------- small vectors
struct f2 { float x, y; };
struct f3 { float x, y, z; };
struct f4 { float x, y, z, w; };
struct u4 { uint32_t x, y, z, w; };

inline f3 F3(float a, float b, float c) { return {a, b, c}; }
inline f4 F4(float a, float b, float c, float d) { return {a, b, c, d}; }
inline f3 operator+(f3 a, f3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline f3 operator-(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline f3 operator-(f3 a) { return {-a.x, -a.y, -a.z}; }
inline f3 operator*(f3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline f3 operator*(float s, f3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline f3 operator*(f3 a, f3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline f4 operator+(f4 a, f4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline f4 operator-(f4 a, f4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline f4 operator*(f4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
inline f4 operator*(f4 a, f4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline f4 operator/(f4 a, float s) { return {a.x / s, a.y / s, a.z / s, a.w / s}; }
inline f2 operator+(f2 a, f2 b) { return {a.x + b.x, a.y + b.y}; }
inline f2 operator-(f2 a, f2 b) { return {a.x - b.x, a.y - b.y}; }
inline f2 operator*(f2 a, float s) { return {a.x * s, a.y * s}; }
inline float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline f3 cross(f3 a, f3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(f3 a) { return sqrtf(dot(a, a)); }
inline float length(f2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
inline f3 normalize(f3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
inline f3 xyz(f4 a) { return {a.x, a.y, a.z}; }
inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
inline float saturate(float v) { return clampf(v, 0.0f, 1.0f); }
inline f3 lerp3(f3 a, f3 b, float t) { return a + (b - a) * t; }
inline f4 lerp4(f4 a, f4 b, float t) { return a + (b - a) * t; }

// mul(M, v) with glm column-major storage m[col][row]
inline f4 mul(const GkMat4& M, f4 v) {
    f4 r;
    r.x = M.m[0][0] * v.x + M.m[1][0] * v.y + M.m[2][0] * v.z + M.m[3][0] * v.w;
    r.y = M.m[0][1] * v.x + M.m[1][1] * v.y + M.m[2][1] * v.z + M.m[3][1] * v.w;
    r.z = M.m[0][2] * v.x + M.m[1][2] * v.y + M.m[2][2] * v.z + M.m[3][2] * v.w;
    r.w = M.m[0][3] * v.x + M.m[1][3] * v.y + M.m[2][3] * v.z + M.m[3][3] * v.w;
    return r;
}
inline GkMat4 matmul(const GkMat4& a, const GkMat4& b) {
    GkMat4 r;
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row)
            r.m[c][row] = a.m[0][row] * b.m[c][0] + a.m[1][row] * b.m[c][1] + a.m[2][row] * b.m[c][2] + a.m[3][row] * b.m[c][3];
    return r;
}

// fp16 storage: R16G16B16A16_SFLOAT stores round to nearest even; loads are exact.
inline uint16_t f2h(float f) { _Float16 h = (_Float16)f; uint16_t b; memcpy(&b, &h, 2); return b; }
inline float h2f(uint16_t b) { _Float16 h; memcpy(&h, &b, 2); return (float)h; }
inline uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// ----------------------------------------------------------------------------------------------- PreProcessor.slang
constexpr float M_PI_F = 3.14159265358979323846f;
constexpr float M_TWO_PI_F = 6.283185307179586476925f;
constexpr float M_PI_4_F = 0.785398163397448309616f;
constexpr float M_1_PI_F = 0.318309886183790671538f;
constexpr float M_1_OVER_TWO_PI_F = 0.159154943091895335768f;
constexpr float EPS = 1e-3f;                     // PreProcessor.slang:13
constexpr float NEARzero = 1e-35f;               // PreProcessor.slang:17
constexpr float PT_MAX_TRACE_DISTANCE = 1000.f;  // Shading.slang:15
constexpr float PRIMARY_T_MAX = 2000.f;          // far plane of the visibility raster (Engine.cpp:747)
constexpr float COS_QUARTER_DEGREE = 0.99999048f; // cos(0.25f / 180.f * M_PI), Shading.slang:897,1138
inline bool isZERO(float x) { return x > -NEARzero && x < NEARzero; }
inline f3 to_local(f3 v, f3 t, f3 b, f3 n) { return {dot(t, v), dot(b, v), dot(n, v)}; }          // PreProcessor.slang:1
inline f3 to_world(f3 v, f3 T, f3 B, f3 N) {                                                      // PreProcessor.slang:2
    return {v.x * T.x + v.y * B.x + v.z * N.x, v.x * T.y + v.y * B.y + v.z * N.y, v.x * T.z + v.y * B.z + v.z * N.z};
}

// ----------------------------------------------------------------------------------------------- Const_Func.slang
inline float pow5(float x) { return x * x * x * x * x; }  // :4
inline float Schlick(float cosine, float refractionIndex) {  // :8-14
    float r0 = (1 - refractionIndex) / (1 + refractionIndex);
    r0 *= r0;
    return r0 + (1 - r0) * pow5(1 - cosine);
}
inline void ONB(f3 n, f3& b1, f3& b2) {  // :17-22
    float signZ = n.z < 0.f ? -1.f : 1.f;
    float a = -1.0f / (signZ + n.z);
    b2 = F3(n.x * n.y * a, signZ + n.y * n.y * a, -n.y);
    b1 = F3(1.0f + signZ * n.x * n.x * a, signZ * b2.x, -signZ * n.x);
}
inline f3 AlignWithNormal(f3 ray, f3 normal) {  // :44-49
    f3 T, B;
    ONB(normal, T, B);
    return to_world(ray, T, B, normal);
}
inline void pcg4d(u4& v) {  // :227-233
    v.x = v.x * 1664525u + 1013904223u; v.y = v.y * 1664525u + 1013904223u;
    v.z = v.z * 1664525u + 1013904223u; v.w = v.w * 1664525u + 1013904223u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    v.x ^= v.x >> 16u; v.y ^= v.y >> 16u; v.z ^= v.z >> 16u; v.w ^= v.w >> 16u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
}
inline float uint_to_float(uint32_t x) { return asfloat(0x3f800000u | (x >> 9)) - 1.0f; }  // :235
inline u4 InitRandomSeed(uint32_t v0, uint32_t v1, uint32_t frame) { return {v0, v1, frame, 0}; }  // :237-240
inline float RandomFloat(u4& v) { pcg4d(v); return uint_to_float(v.x); }  // :242-246
inline f2 RandomFloat2(u4& v) { pcg4d(v); return {uint_to_float(v.x), uint_to_float(v.y)}; }  // :248-252
inline f2 concentric_sample_disk(f2 offset) {  // :260-275
    offset = offset + (offset - f2{1, 1});
    if (isZERO(offset.x) && isZERO(offset.y)) return {0, 0};
    float theta;
    if (fabsf(offset.x) > fabsf(offset.y)) {
        theta = M_PI_4_F * offset.y / offset.x;
        return f2{gk_det_cosf(theta), gk_det_sinf(theta)} * offset.x;
    }
    float cos_theta = gk_det_sinf(M_PI_4_F * offset.x / offset.y);
    return f2{cos_theta, sqrtf(1.f - cos_theta * cos_theta)} * offset.y;
}
inline f2 RandomInUnitDisk(u4& seed) { return concentric_sample_disk(RandomFloat2(seed)); }  // :277-280
inline f3 RandomInCone(u4& seed, float cos_theta) {  // :282-289
    const f2 u = RandomFloat2(seed);
    float phi = M_TWO_PI_F * u.x;
    cos_theta = 1.0f + u.y * (cos_theta - 1.f);
    float s, c;
    gk_det_sincosf(phi, &s, &c);
    float r = sqrtf(1.0f - cos_theta * cos_theta);
    return F3(r * c, r * s, cos_theta);
}
inline f3 RandomInHemiSphere1(u4& seed) {  // :299-305
    const f2 u = RandomFloat2(seed);
    float phi = M_TWO_PI_F * u.x;
    float s, c;
    gk_det_sincosf(phi, &s, &c);
    float r = sqrtf(u.y);
    return F3(r * c, r * s, sqrtf(1.0f - u.y));
}
inline f3 ggx_sample_vndf(f2 alpha, f3 wi_, f2 uv) {  // :310-328 (Eto & Tokuyoshi 2023 bounded VNDF)
    f3 wi = normalize(F3(wi_.x * alpha.x, wi_.y * alpha.y, wi_.z));
    float b = wi.z;
    if (wi_.z > 0.f) {
        float a = saturate(fminf(alpha.x, alpha.y));
        float awiz_s = a * wi_.z / (1.0f + length(f2{wi_.x, wi_.y}));
        b *= ((1.0f - a * a) / (1.0f + awiz_s * awiz_s));
    }
    float z = fmaf(1.0f - uv.y, 1.0f + b, -b);
    float phi = M_TWO_PI_F * uv.x;
    float s, c;
    gk_det_sincosf(phi, &s, &c);
    float r = sqrtf(saturate(1.0f - z * z));
    f3 o_std = F3(r * c, r * s, z);
    f3 m_std = wi + o_std;
    return normalize(F3(m_std.x * alpha.x, m_std.y * alpha.y, m_std.z));
}
inline f3 ggxSampling(u4& RandomSeed, float roughness, f3 normal) {  // :330-340
    f3 tangent, bitangent;
    ONB(normal, tangent, bitangent);
    f3 wm_ = ggx_sample_vndf(f2{roughness * roughness, roughness * roughness}, to_local(normal, tangent, bitangent, normal), RandomFloat2(RandomSeed));
    return to_world(wm_, tangent, bitangent, normal);
}
inline f3 reflect3(f3 i, f3 n) { return i - n * (2.0f * dot(n, i)); }  // HLSL reflect
inline f3 refract3(f3 i, f3 n, float eta) {  // HLSL refract
    float ni = dot(n, i);
    float k = 1.0f - eta * eta * (1.0f - ni * ni);
    if (k < 0.0f) return F3(0, 0, 0);
    return i * eta - n * (eta * ni + sqrtf(k));
}

// tonemappers, Const_Func.slang:51-68, 82-126
inline f3 LinearToST2084UE(f3 lin) {
    const float m1 = 0.1593017578125f, m2 = 78.84375f, c1 = 0.8359375f, c2 = 18.8515625f, c3 = 18.6875f, C = 10000.f;
    float in[3] = {lin.x, lin.y, lin.z}, out[3];
    for (int i = 0; i < 3; ++i) {
        float L = in[i] / C;
        float Lm = gk_det_powf(L, m1);
        float N1 = (c1 + c2 * Lm), N2 = (1.0f + c3 * Lm);
        float N = N1 * (1.0f / N2);
        out[i] = gk_det_powf(N, m2);
    }
    return F3(out[0], out[1], out[2]);
}
inline float W_f(float x, float e0, float e1) {
    if (x <= e0) return 0;
    if (x >= e1) return 1;
    float a = (x - e0) / (e1 - e0);
    return a * a * (3 - 2 * a);
}
inline float H_f(float x, float e0, float e1) {
    if (x <= e0) return 0;
    if (x >= e1) return 1;
    return (x - e0) / (e1 - e0);
}
inline float GranTurismoTonemapper(float x) {
    const float e = 2.71828f;
    float P = 1, a = 0.7f, m = 0.22f, l = 0.4f, c = 1.33f, b = 0;
    float l0 = (P - m) * l / a;
    float L_x = m + a * (x - m);
    float T_x = m * gk_det_powf(x / m, c) + b;
    float S0 = m + l0;
    float S1 = m + a * l0;
    float C2 = a * P / (P - S1);
    float S_x = P - (P - S1) * gk_det_powf(e, -(C2 * (x - S0) / P));
    float w0_x = 1 - W_f(x, 0, m);
    float w2_x = H_f(x, m + l0, m + l0);
    float w1_x = 1 - w0_x - w2_x;
    return T_x * w0_x + L_x * w1_x + S_x * w2_x;
}
inline f3 GT_Tonemapping(f3 x) { return F3(GranTurismoTonemapper(x.x), GranTurismoTonemapper(x.y), GranTurismoTonemapper(x.z)); }

// ----------------------------------------------------------------------------------------------- scene data
struct Vertex {  // BasicTypes.slang:238-247
    f3 Position{}, Normal{};
    f4 Tangent{};
    f2 TexCoord{};
    uint32_t MaterialIndex = 0;
    f3 LocalNormal{};
    f4 LocalTangent{};
};

inline Vertex UnpackVertexV2(const GkGPUVertex& g) {  // Const_Func.slang:356-368
    Vertex v;
    v.Position = F3(h2f(g.posx), h2f(g.posy), h2f(g.posz));
    v.Normal = F3(h2f(g.normalx), h2f(g.normaly), h2f(g.normalz));
    v.Tangent = F4(h2f(g.tangentx), h2f(g.tangenty), h2f(g.tangentz), (float)((g.tangentw >> 8) & 0xFF));
    v.TexCoord = f2{h2f(g.texcoordx), h2f(g.texcoordy)};
    v.MaterialIndex = g.tangentw & 0xFF;
    return v;
}
inline uint32_t FetchMaterialId(const GkNodeProxy& node, uint32_t offset) { return offset < 16 ? node.matId[offset] : 0; }  // :370-424

struct Hit { float t = 0, u = 0, v = 0; uint32_t inst = 0xFFFFFFFFu, prim = 0; };

struct Tri { f3 v0, v1, v2; };

// ----------------------------------------------------------------------------------------------- native intersector
// BVH2 with float boxes, binned SAH. Closest hit is defined order-independently: among all triangles with
// tmin < t < tmax the minimum of (t, instance, primitive) wins, so any conservative traversal gives the same answer.
struct Box { f3 lo{1e30f, 1e30f, 1e30f}, hi{-1e30f, -1e30f, -1e30f};
    void grow(f3 p) { lo = F3(fminf(lo.x, p.x), fminf(lo.y, p.y), fminf(lo.z, p.z)); hi = F3(fmaxf(hi.x, p.x), fmaxf(hi.y, p.y), fmaxf(hi.z, p.z)); }
    void grow(const Box& b) { grow(b.lo); grow(b.hi); }
    float area() const { f3 d = hi - lo; return d.x * d.y + d.y * d.z + d.z * d.x; }
};
struct BNode { Box box; uint32_t left = 0, right = 0, first = 0, count = 0; };
struct Bvh {
    std::vector<BNode> nodes;
    std::vector<uint32_t> order;
    void build(const std::vector<Box>& boxes) {
        uint32_t n = (uint32_t)boxes.size();
        order.resize(n);
        for (uint32_t i = 0; i < n; ++i) order[i] = i;
        nodes.clear();
        nodes.reserve(2 * n + 1);
        nodes.emplace_back();
        if (n == 0) return;
        std::vector<f3> cent(n);
        for (uint32_t i = 0; i < n; ++i) cent[i] = (boxes[i].lo + boxes[i].hi) * 0.5f;
        subdivide(0, 0, n, boxes, cent);
    }
    void subdivide(uint32_t ni, uint32_t first, uint32_t count, const std::vector<Box>& boxes, const std::vector<f3>& cent) {
        Box b, cb;
        for (uint32_t i = first; i < first + count; ++i) { b.grow(boxes[order[i]]); cb.grow(cent[order[i]]); }
        nodes[ni].box = b;
        nodes[ni].first = first;
        nodes[ni].count = count;
        if (count <= 2) return;
        const int BINS = 8;
        int bestAxis = -1, bestSplit = 0;
        float bestCost = 1e30f;
        for (int axis = 0; axis < 3; ++axis) {
            float lo = (&cb.lo.x)[axis], hi = (&cb.hi.x)[axis];
            if (hi <= lo) continue;
            Box bb[BINS];
            uint32_t bc[BINS] = {};
            float scale = BINS / (hi - lo);
            for (uint32_t i = first; i < first + count; ++i) {
                int k = std::min(BINS - 1, (int)(((&cent[order[i]].x)[axis] - lo) * scale));
                bb[k].grow(boxes[order[i]]);
                bc[k]++;
            }
            float la[BINS - 1], ra[BINS - 1];
            uint32_t lc[BINS - 1], rc[BINS - 1];
            Box l, r;
            uint32_t ls = 0, rs = 0;
            for (int k = 0; k < BINS - 1; ++k) {
                ls += bc[k]; if (bc[k]) l.grow(bb[k]); lc[k] = ls; la[k] = ls ? l.area() : 0;
                rs += bc[BINS - 1 - k]; if (bc[BINS - 1 - k]) r.grow(bb[BINS - 1 - k]); rc[BINS - 2 - k] = rs; ra[BINS - 2 - k] = rs ? r.area() : 0;
            }
            for (int k = 0; k < BINS - 1; ++k) {
                if (!lc[k] || !rc[k]) continue;
                float cost = lc[k] * la[k] + rc[k] * ra[k];
                if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestSplit = k; }
            }
        }
        uint32_t mid;
        if (bestAxis < 0) {
            if (count <= 4) return;
            mid = first + count / 2;
        } else {
            if (count <= 4 && bestCost >= count * b.area()) return;
            float lo = (&cb.lo.x)[bestAxis], hi = (&cb.hi.x)[bestAxis];
            float scale = 8 / (hi - lo);
            auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](uint32_t id) {
                return std::min(7, (int)(((&cent[id].x)[bestAxis] - lo) * scale)) <= bestSplit;
            });
            mid = (uint32_t)(it - order.begin());
            if (mid == first || mid == first + count) mid = first + count / 2;
        }
        uint32_t l = (uint32_t)nodes.size();
        nodes.emplace_back();
        nodes.emplace_back();
        nodes[ni].left = l;
        nodes[ni].right = l + 1;
        nodes[ni].count = 0;
        subdivide(l, first, mid - first, boxes, cent);
        subdivide(l + 1, mid, first + count - mid, boxes, cent);
    }
};

// Conservative slab test: true when [tn, tf] padded by a few ulps overlaps (tmin, tbest].
inline bool slab(const Box& b, f3 o, f3 id, float tmin, float tbest) {
    float tx0 = (b.lo.x - o.x) * id.x, tx1 = (b.hi.x - o.x) * id.x;
    float ty0 = (b.lo.y - o.y) * id.y, ty1 = (b.hi.y - o.y) * id.y;
    float tz0 = (b.lo.z - o.z) * id.z, tz1 = (b.hi.z - o.z) * id.z;
    // A ray that lies IN a face plane of the box (direction component 0, origin on the plane) gives a slab interval that ends at
    // exactly 0; every axis interval is therefore widened by 1e-14 world units (|id| * 1e-14 in t), far below fp32 resolution
    // unless the origin sits on the plane: the test is inclusive, so every triangle whose box the ray touches is tested and the
    // closest hit does not depend on the tree (nor on how far a quantised tree rounds its boxes outward).
    const float px = fabsf(id.x) * 1e-14f, py = fabsf(id.y) * 1e-14f, pz = fabsf(id.z) * 1e-14f;
    float tn = fmaxf(fmaxf(fminf(tx0, tx1) - px, fminf(ty0, ty1) - py), fmaxf(fminf(tz0, tz1) - pz, tmin));
    float tf = fminf(fminf(fmaxf(tx0, tx1) + px, fmaxf(ty0, ty1) + py), fminf(fmaxf(tz0, tz1) + pz, tbest));
    return tn * 0.9999990f - 1e-7f <= tf * 1.0000010f + 1e-7f;
}

// Watertight ray/triangle test (Woop, Benthin, Wald 2013). Per-ray constants:
struct RayShear { int kx, ky, kz; float Sx, Sy, Sz; };
inline RayShear make_shear(f3 d) {
    RayShear r;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    r.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    r.kx = r.kz + 1; if (r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if (r.ky == 3) r.ky = 0;
    const float dd[3] = {d.x, d.y, d.z};
    if (dd[r.kz] < 0.0f) std::swap(r.kx, r.ky);
    r.Sx = dd[r.kx] / dd[r.kz];
    r.Sy = dd[r.ky] / dd[r.kz];
    r.Sz = 1.0f / dd[r.kz];
    return r;
}
// returns true and (t, b1, b2) = (distance, weight of v1, weight of v2); no culling, no range test
inline bool tri_test(const Tri& tr, f3 o, const RayShear& s, float& t, float& b1, float& b2) {
    const f3 A = tr.v0 - o, B = tr.v1 - o, C = tr.v2 - o;
    const float a[3] = {A.x, A.y, A.z}, b[3] = {B.x, B.y, B.z}, c[3] = {C.x, C.y, C.z};
    const float Ax = a[s.kx] - s.Sx * a[s.kz], Ay = a[s.ky] - s.Sy * a[s.kz];
    const float Bx = b[s.kx] - s.Sx * b[s.kz], By = b[s.ky] - s.Sy * b[s.kz];
    const float Cx = c[s.kx] - s.Sx * c[s.kz], Cy = c[s.ky] - s.Sy * c[s.kz];
    float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
        V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
        W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float Az = s.Sz * a[s.kz], Bz = s.Sz * b[s.kz], Cz = s.Sz * c[s.kz];
    const float T = U * Az + V * Bz + W * Cz;
    const float inv = 1.0f / det;
    t = T * inv;
    b1 = V * inv;
    b2 = W * inv;
    return true;
}

struct Blas {
    std::vector<Tri> tris;  // in primitive order
    Bvh bvh;
    Box bounds;
#ifdef GKO_TINYBVH
    std::vector<tinybvh::bvhvec4> tverts;
    tinybvh::BVH tbvh;
#endif
};

// world -> object, rows of the 3x4 matrix; same adjugate formula as the CUDA instance setup.
struct Instance { float w2o[3][4]; float o2w[3][4]; uint32_t blas; bool active; Box box; };

inline void invert_affine(const GkMat4& M, float w2o[3][4], float o2w[3][4]) {
    // o2w row r, col c = M.m[c][r]
    float a[3][3], t[3];
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) { a[r][c] = M.m[c][r]; o2w[r][c] = M.m[c][r]; }
        t[r] = M.m[3][r];
        o2w[r][3] = t[r];
    }
    float c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
    float c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
    float c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
    float det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
    float id = 1.0f / det;
    float inv[3][3];
    inv[0][0] = c00 * id;
    inv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
    inv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
    inv[1][0] = c01 * id;
    inv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
    inv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
    inv[2][0] = c02 * id;
    inv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
    inv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) w2o[r][c] = inv[r][c];
        w2o[r][3] = -(inv[r][0] * t[0] + inv[r][1] * t[1] + inv[r][2] * t[2]);
    }
}
inline f3 xform_point(const float m[3][4], f3 p) {
    return F3(m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3], m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3],
              m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3]);
}
inline f3 xform_vec(const float m[3][4], f3 p) {
    return F3(m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z, m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z,
              m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z);
}
inline f3 safe_inv(f3 d) {
    auto inv = [](float x) { return 1.0f / (fabsf(x) > 1e-30f ? x : (x < 0 ? -1e-30f : 1e-30f)); };
    return F3(inv(d.x), inv(d.y), inv(d.z));
}

struct Texture { uint32_t w = 0, h = 0; std::vector<f4> texels; };

}  // namespace

struct GkoScene {
    std::vector<GkGPUVertex> vertices;
    std::vector<uint16_t> simple;
    std::vector<uint32_t> indices;
    std::vector<GkModelData> offsets;
    uint32_t model_count = 0;
    std::vector<GkMaterial> materials;
    std::vector<GkNodeProxy> nodes;
    std::vector<Blas> blas;
    std::vector<Instance> instances;
    Bvh tlas;
    std::vector<Texture> textures;
    std::vector<GkAmbientCube> cubes;
    std::vector<GkVoxelData> voxels;
    std::vector<GkPageIndex> pages;
    std::vector<GkLightObject> lights;
    std::vector<GkSphericalHarmonics> sky_sh;
    std::vector<uint32_t> raster_indices, reorders;  // the raster pass's index buffer and GPUScene.Reorders (Scene.cpp:240-250)
#ifdef GKO_TINYBVH
    tinybvh::BVH ttlas;
    std::vector<tinybvh::BLASInstance> tinst;
    std::vector<uint32_t> tinst_to_node;
    std::vector<tinybvh::BVHBase*> tblas_list;
#endif

    GkMaterial GetMaterial(int idx) const {  // BasicTypes.slang:376-379
        int i = std::min(GK_MAX_MATERIALS - 1, std::max(0, idx));
        if (i >= (int)materials.size()) { GkMaterial z{}; return z; }
        return materials[i];
    }

    void build_blas(uint32_t m) {
        Blas& b = blas[m];
        const GkModelData& md = offsets[m * 10];
        uint32_t total = 0;
        for (int s = 0; s < 10; ++s) total += offsets[m * 10 + s].indexCount;
        uint32_t ntri = total / 3;
        b.tris.resize(ntri);
        std::vector<Box> boxes(ntri);
        b.bounds = Box();
        for (uint32_t p = 0; p < ntri; ++p) {
            f3 v[3];
            for (int k = 0; k < 3; ++k) {
                uint32_t vi = md.vertexOffset + indices[md.indexOffset + p * 3 + k];
                v[k] = F3(h2f(simple[vi * 4 + 0]), h2f(simple[vi * 4 + 1]), h2f(simple[vi * 4 + 2]));
                boxes[p].grow(v[k]);
            }
            b.tris[p] = {v[0], v[1], v[2]};
            b.bounds.grow(boxes[p]);
        }
        b.bvh.build(boxes);
#ifdef GKO_TINYBVH
        b.tverts.resize(ntri * 3);
        for (uint32_t p = 0; p < ntri; ++p) {
            b.tverts[p * 3 + 0] = tinybvh::bvhvec4(b.tris[p].v0.x, b.tris[p].v0.y, b.tris[p].v0.z, 0);
            b.tverts[p * 3 + 1] = tinybvh::bvhvec4(b.tris[p].v1.x, b.tris[p].v1.y, b.tris[p].v1.z, 0);
            b.tverts[p * 3 + 2] = tinybvh::bvhvec4(b.tris[p].v2.x, b.tris[p].v2.y, b.tris[p].v2.z, 0);
        }
        if (ntri) b.tbvh.Build(b.tverts.data(), ntri);  // CPUAccelerationStructure.cpp:225
#endif
    }

    void build_tlas() {
        instances.resize(nodes.size());
        std::vector<Box> boxes(nodes.size());
        for (size_t i = 0; i < nodes.size(); ++i) {
            Instance& in = instances[i];
            const GkNodeProxy& n = nodes[i];
            in.blas = n.modelId / 10;  // RayTraceBaseRenderer.cpp:178
            in.active = n.visible && !n.nort && in.blas < blas.size() && !blas[in.blas].tris.empty();
            invert_affine(n.worldTS, in.w2o, in.o2w);
            Box wb;
            if (in.active) {
                const Box& lb = blas[in.blas].bounds;
                for (int c = 0; c < 8; ++c) {
                    f3 p = F3(c & 1 ? lb.hi.x : lb.lo.x, c & 2 ? lb.hi.y : lb.lo.y, c & 4 ? lb.hi.z : lb.lo.z);
                    wb.grow(xform_point(in.o2w, p));
                }
                f3 pad = (wb.hi - wb.lo) * 1e-6f + F3(1e-6f, 1e-6f, 1e-6f);
                wb.lo = wb.lo - pad;
                wb.hi = wb.hi + pad;
            } else {
                wb.lo = F3(1e30f, 1e30f, 1e30f);
                wb.hi = F3(1e30f, 1e30f, 1e30f);
            }
            in.box = wb;
            boxes[i] = wb;
        }
        tlas.build(boxes);
#ifdef GKO_TINYBVH
        tinst.clear();
        tinst_to_node.clear();
        tblas_list.clear();
        for (auto& b : blas) tblas_list.push_back(&b.tbvh);
        for (size_t i = 0; i < nodes.size(); ++i) {
            if (!instances[i].active) continue;
            tinybvh::BLASInstance bi;
            bi.blasIdx = instances[i].blas;
            float* tf = (float*)bi.transform;  // row-major 4x4 = transpose(worldTS), CPUAccelerationStructure.cpp:250-257
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) tf[r * 4 + c] = nodes[i].worldTS.m[c][r];
            tinst.push_back(bi);
            tinst_to_node.push_back((uint32_t)i);
        }
        if (!tinst.empty()) ttlas.Build(tinst.data(), (int)tinst.size(), tblas_list.data(), (int)tblas_list.size());  // :272
#endif
    }

#ifndef GKO_TINYBVH
    void blas_closest(const Blas& b, uint32_t inst, f3 o, f3 d, float tmin, float tmax, Hit& best) const {
        if (b.tris.empty()) return;
        const f3 id = safe_inv(d);
        const RayShear sh = make_shear(d);
        uint32_t stack[64];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const BNode& n = b.bvh.nodes[stack[--sp]];
            float tb = best.inst == 0xFFFFFFFFu ? tmax : best.t;
            if (!slab(n.box, o, id, tmin, tb)) continue;
            if (n.count) {
                for (uint32_t i = n.first; i < n.first + n.count; ++i) {
                    uint32_t p = b.bvh.order[i];
                    float t, b1, b2;
                    if (!tri_test(b.tris[p], o, sh, t, b1, b2)) continue;
                    if (!(t > tmin && t < tmax)) continue;
                    bool better = best.inst == 0xFFFFFFFFu || t < best.t ||
                                  (t == best.t && (inst < best.inst || (inst == best.inst && p < best.prim)));
                    if (better) { best.t = t; best.u = b1; best.v = b2; best.inst = inst; best.prim = p; }
                }
            } else {
                stack[sp++] = n.left;
                stack[sp++] = n.right;
            }
        }
    }
    bool blas_any(const Blas& b, f3 o, f3 d, float tmin, float tmax) const {
        if (b.tris.empty()) return false;
        const f3 id = safe_inv(d);
        const RayShear sh = make_shear(d);
        uint32_t stack[64];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const BNode& n = b.bvh.nodes[stack[--sp]];
            if (!slab(n.box, o, id, tmin, tmax)) continue;
            if (n.count) {
                for (uint32_t i = n.first; i < n.first + n.count; ++i) {
                    float t, b1, b2;
                    if (tri_test(b.tris[b.bvh.order[i]], o, sh, t, b1, b2) && t > tmin && t < tmax) return true;
                }
            } else {
                stack[sp++] = n.left;
                stack[sp++] = n.right;
            }
        }
        return false;
    }
#endif

    // RayQuery closest hit: TMin < t < TMax, no culling, mask 0xFF (Shading.slang:784-801)
    Hit closest(f3 o, f3 d, float tmin, float tmax) const {
        Hit best;
#ifdef GKO_TINYBVH
        if (tinst.empty()) return best;
        // tinybvh::Ray has no tmin and normalises D (tiny_bvh.h:556-572): advance the origin by tmin instead.
        float len = length(d);
        f3 dn = d * (1.0f / len);
        f3 o2 = o + dn * (tmin * len);
        tinybvh::Ray ray(tinybvh::bvhvec3(o2.x, o2.y, o2.z), tinybvh::bvhvec3(dn.x, dn.y, dn.z), (tmax - tmin) * len);
        ttlas.Intersect(ray);
        if (ray.hit.t < (tmax - tmin) * len) {
            best.t = ray.hit.t / len + tmin;
            best.u = ray.hit.u;
            best.v = ray.hit.v;
            best.inst = tinst_to_node[ray.hit.inst];
            best.prim = ray.hit.prim;
        }
        return best;
#else
        if (tlas.order.empty()) return best;
        const f3 id = safe_inv(d);
        uint32_t stack[64];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const BNode& n = tlas.nodes[stack[--sp]];
            float tb = best.inst == 0xFFFFFFFFu ? tmax : best.t;
            if (!slab(n.box, o, id, tmin, tb)) continue;
            if (n.count) {
                for (uint32_t i = n.first; i < n.first + n.count; ++i) {
                    uint32_t ii = tlas.order[i];
                    const Instance& in = instances[ii];
                    if (!in.active) continue;
                    float tb2 = best.inst == 0xFFFFFFFFu ? tmax : best.t;
                    if (!slab(in.box, o, id, tmin, tb2)) continue;
                    blas_closest(blas[in.blas], ii, xform_point(in.w2o, o), xform_vec(in.w2o, d), tmin, tmax, best);
                }
            } else {
                stack[sp++] = n.left;
                stack[sp++] = n.right;
            }
        }
        return best;
#endif
    }

    // RayQuery any hit: ACCEPT_FIRST_HIT_AND_END_SEARCH (Shading.slang:737-757)
    bool occluded(f3 o, f3 d, float tmin, float tmax) const {
#ifdef GKO_TINYBVH
        if (tinst.empty()) return false;
        float len = length(d);
        f3 dn = d * (1.0f / len);
        f3 o2 = o + dn * (tmin * len);
        tinybvh::Ray ray(tinybvh::bvhvec3(o2.x, o2.y, o2.z), tinybvh::bvhvec3(dn.x, dn.y, dn.z), (tmax - tmin) * len);
        return ttlas.IsOccluded(ray);
#else
        if (tlas.order.empty()) return false;
        const f3 id = safe_inv(d);
        uint32_t stack[64];
        int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const BNode& n = tlas.nodes[stack[--sp]];
            if (!slab(n.box, o, id, tmin, tmax)) continue;
            if (n.count) {
                for (uint32_t i = n.first; i < n.first + n.count; ++i) {
                    const Instance& in = instances[tlas.order[i]];
                    if (!in.active || !slab(in.box, o, id, tmin, tmax)) continue;
                    if (blas_any(blas[in.blas], xform_point(in.w2o, o), xform_vec(in.w2o, d), tmin, tmax)) return true;
                }
            } else {
                stack[sp++] = n.left;
                stack[sp++] = n.right;
            }
        }
        return false;
#endif
    }

    // Sampler2D.SampleLevel(uv, 0): linear filter, repeat addressing (src/Vulkan/Sampler.hpp:9-26)
    f4 sample_texture(uint32_t id, f2 uv) const {
        if (id >= textures.size() || textures[id].w == 0) return F4(0, 0, 0, 0);
        const Texture& tx = textures[id];
        float x = uv.x * (float)tx.w - 0.5f, y = uv.y * (float)tx.h - 0.5f;
        float fx = floorf(x), fy = floorf(y);
        float ax = x - fx, ay = y - fy;
        int x0 = (int)fx, y0 = (int)fy;
        auto wrap = [](int v, int n) { int r = v % n; return r < 0 ? r + n : r; };
        int xa = wrap(x0, (int)tx.w), xb = wrap(x0 + 1, (int)tx.w), ya = wrap(y0, (int)tx.h), yb = wrap(y0 + 1, (int)tx.h);
        f4 t00 = tx.texels[(size_t)ya * tx.w + xa], t10 = tx.texels[(size_t)ya * tx.w + xb];
        f4 t01 = tx.texels[(size_t)yb * tx.w + xa], t11 = tx.texels[(size_t)yb * tx.w + xb];
        return lerp4(lerp4(t00, t10, ax), lerp4(t01, t11, ax), ay);
    }

    f4 SampleIBLV2(uint32_t skyIdx, f3 direction, float rotate, float roughness) const {  // Shading.slang:194-199
        (void)roughness;  // the path tracer always passes 0 -> mip 0
        f3 d = normalize(direction);
        f2 t = f2{(gk_det_atan2f(d.x, d.z) + M_PI_F * rotate) * M_1_OVER_TWO_PI_F, gk_det_acosf(d.y) * M_1_PI_F};
        f4 s = sample_texture(skyIdx, t);
        return F4(fminf(10.f, s.x), fminf(10.f, s.y), fminf(10.f, s.z), fminf(1.f, s.w));
    }

    // AmbientCube.slang:71-78, 81-89, 178-223, 290-382
    static f4 unpackRGB10A2(uint32_t p) {
        float r = float(p & 0x3FF) / 1023.0f, g = float((p >> 10) & 0x3FF) / 1023.0f, b = float((p >> 20) & 0x3FF) / 1023.0f;
        return F4(r, g, b, 0.0f) * 512.f;
    }
    static f4 sampleAmbientCubeHL2_Full(const GkAmbientCube& cube, f3 normal) {
        f4 color = F4(0, 0, 0, 0);
        float skyvis = 0.0f;
        float wx = fmaxf(normal.x, 0.0f), wnx = fmaxf(-normal.x, 0.0f), wy = fmaxf(normal.y, 0.0f), wny = fmaxf(-normal.y, 0.0f),
              wz = fmaxf(normal.z, 0.0f), wnz = fmaxf(-normal.z, 0.0f);
        float sum = wx + wnx + wy + wny + wz + wnz;
        uint32_t s0 = cube.skyVisibility_pznzpyny, s1 = cube.skyVisibility_pxnxs0s1;
        skyvis += float(s0 & 0xFF) * wz;
        skyvis += float((s0 >> 8) & 0xFF) * wnz;
        skyvis += float((s0 >> 16) & 0xFF) * wy;
        skyvis += float((s0 >> 24) & 0xFF) * wny;
        skyvis += float(s1 & 0xFF) * wx;
        skyvis += float((s1 >> 8) & 0xFF) * wnx;
        color = color + unpackRGB10A2(cube.PosX_D) * wx; color = color + unpackRGB10A2(cube.NegX_D) * wnx;
        color = color + unpackRGB10A2(cube.PosY_D) * wy; color = color + unpackRGB10A2(cube.NegY_D) * wny;
        color = color + unpackRGB10A2(cube.PosZ_D) * wz; color = color + unpackRGB10A2(cube.NegZ_D) * wnz;
        color = color + unpackRGB10A2(cube.PosX) * wx; color = color + unpackRGB10A2(cube.NegX) * wnx;
        color = color + unpackRGB10A2(cube.PosY) * wy; color = color + unpackRGB10A2(cube.NegY) * wny;
        color = color + unpackRGB10A2(cube.PosZ) * wz; color = color + unpackRGB10A2(cube.NegZ) * wnz;
        float k = (sum > 0.0f) ? (1.0f / sum) : 1.0f;
        color.x *= k; color.y *= k; color.z *= k;
        skyvis *= k;
        color.w = skyvis / 255.0f;
        return color;
    }
    static f4 sampleAmbientCubeHL2_DI(const GkAmbientCube& cube, f3 normal) {  // AmbientCube.slang:128-151
        f4 color = F4(0, 0, 0, 0);
        float wx = fmaxf(normal.x, 0.0f), wnx = fmaxf(-normal.x, 0.0f), wy = fmaxf(normal.y, 0.0f), wny = fmaxf(-normal.y, 0.0f),
              wz = fmaxf(normal.z, 0.0f), wnz = fmaxf(-normal.z, 0.0f);
        float sum = wx + wnx + wy + wny + wz + wnz;
        color = color + unpackRGB10A2(cube.PosX_D) * wx; color = color + unpackRGB10A2(cube.NegX_D) * wnx;
        color = color + unpackRGB10A2(cube.PosY_D) * wy; color = color + unpackRGB10A2(cube.NegY_D) * wny;
        color = color + unpackRGB10A2(cube.PosZ_D) * wz; color = color + unpackRGB10A2(cube.NegZ_D) * wnz;
        return color * ((sum > 0.0f) ? (1.0f / sum) : 1.0f);
    }
    static uint32_t packRGB10A2(f4 color) {  // AmbientCube.slang:58-69
        f4 c = F4(clampf(color.x / 512.f, 0.0f, 1.0f), clampf(color.y / 512.f, 0.0f, 1.0f), clampf(color.z / 512.f, 0.0f, 1.0f),
                  clampf(color.w / 512.f, 0.0f, 1.0f));
        uint32_t r = (uint32_t)(c.x * 1023.0f), g = (uint32_t)(c.y * 1023.0f), b = (uint32_t)(c.z * 1023.0f), a = (uint32_t)(c.w * 3.0f);
        return r | (g << 10) | (b << 20) | (a << 30);
    }
    static uint32_t LerpPackedColorAlt(uint32_t c0, f4 c1, float t) { return packRGB10A2(lerp4(unpackRGB10A2(c0), c1, t)); }  // :112-119

    // Shading.slang:146-192: EvaluateSH + SampleIBLRough
    f4 SampleIBLRough(uint32_t skyIdx, f3 normal, float rotate) const {
        if (skyIdx >= sky_sh.size()) return F4(0, 0, 0, 1);
        const float angle = (1.0f - rotate) * 3.14159265358979323846f;
        float sinAngle, cosAngle;
        gk_det_sincosf(angle, &sinAngle, &cosAngle);
        const f3 n = F3(normal.x * cosAngle + normal.z * sinAngle, normal.y, -normal.x * sinAngle + normal.z * cosAngle);
        const float SH_C0 = 0.282095f, SH_C1 = 0.488603f, SH_C2 = 1.092548f, SH_C3 = 0.315392f, SH_C4 = 0.546274f;
        float basis[9];
        basis[0] = SH_C0;
        basis[1] = -SH_C1 * n.y;
        basis[2] = SH_C1 * n.z;
        basis[3] = -SH_C1 * n.x;
        basis[4] = SH_C2 * n.x * n.y;
        basis[5] = -SH_C2 * n.y * n.z;
        basis[6] = SH_C3 * (3.f * n.y * n.y - 1.0f);
        basis[7] = -SH_C2 * n.x * n.z;
        basis[8] = SH_C4 * (n.x * n.x - n.z * n.z);
        const GkSphericalHarmonics& sh = sky_sh[skyIdx];
        f3 color = F3(0, 0, 0);
        for (int i = 0; i < 9; ++i) {
            color.x += sh.coefficients[0][i] * basis[i];
            color.y += sh.coefficients[1][i] * basis[i];
            color.z += sh.coefficients[2][i] * basis[i];
        }
        return F4(color.x, color.y, color.z, 1.0f);
    }

    bool TraceRayVertex(f3 rayOrigin, f3 rayDir, float maxDistance, Vertex& outVertex, GkNodeProxy& outNode) const {
        Hit h = closest(rayOrigin, rayDir, EPS, maxDistance);
        if (h.inst == 0xFFFFFFFFu) return false;
        f3 Bary = F3(1.0f - h.u - h.v, h.u, h.v);
        outNode = nodes[h.inst];
        const GkModelData& model = offsets[outNode.modelId];
        const uint32_t indexOffset = model.indexOffset + h.prim * 3;
        const uint32_t vertexOffset = model.vertexOffset;
        const Vertex v0 = UnpackVertexV2(vertices[vertexOffset + indices[indexOffset]]);
        const Vertex v1 = UnpackVertexV2(vertices[vertexOffset + indices[indexOffset + 1]]);
        const Vertex v2 = UnpackVertexV2(vertices[vertexOffset + indices[indexOffset + 2]]);
        outVertex.MaterialIndex = FetchMaterialId(outNode, v0.MaterialIndex);
        // Mix(a,b,c,bary) = a + (b-a)*bary.y + (c-a)*bary.z (PreProcessor.slang:24)
        f3 n = v0.Normal + (v1.Normal - v0.Normal) * Bary.y + (v2.Normal - v0.Normal) * Bary.z;
        // mul(WorldToObject4x3, n).xyz = transpose(world-to-object 3x3) * n
        const Instance& in = instances[h.inst];
        f3 nw = F3(in.w2o[0][0] * n.x + in.w2o[1][0] * n.y + in.w2o[2][0] * n.z, in.w2o[0][1] * n.x + in.w2o[1][1] * n.y + in.w2o[2][1] * n.z,
                   in.w2o[0][2] * n.x + in.w2o[1][2] * n.y + in.w2o[2][2] * n.z);
        outVertex.Normal = normalize(nw);
        outVertex.TexCoord = v0.TexCoord + (v1.TexCoord - v0.TexCoord) * Bary.y + (v2.TexCoord - v0.TexCoord) * Bary.z;
        outVertex.Position = rayOrigin + rayDir * h.t;
        return true;
    }

    // ------------------------------------------------------------------------------------------ probe grid
    static f3 probe_origin(uint32_t gIdx) {  // Bake.HwAmbientCube.comp.slang:12-17
        const uint32_t y = gIdx / (GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY);
        const uint32_t z = (gIdx - y * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY) / GK_CUBE_SIZE_XY;
        const uint32_t x = gIdx - y * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY - z * GK_CUBE_SIZE_XY;
        const f3 CUBE_OFFSET = F3(-GK_CUBE_SIZE_XY / 2, -1.375f, -GK_CUBE_SIZE_XY / 2) * 0.25f;
        return F3((float)x, (float)y, (float)z) * 0.25f + CUBE_OFFSET;
    }
    // TraceRay of the CPU baker, src/Assets/CPUAccelerationStructure.cpp:42-66 (normalised direction, no tmin, face normal
    // through the world matrix, material of vertex 0), on the fp16 geometry.
    bool VoxTrace(f3 origin, f3 dir, float dist, float& outT, f3& outNormal, uint32_t& outMat) const {
        const f3 d = normalize(dir);
        Hit h = closest(origin, d, 0.0f, dist);
        if (h.inst == 0xFFFFFFFFu) return false;
        const GkNodeProxy& node = nodes[h.inst];
        const GkModelData& off = offsets[node.modelId];
        const Vertex v0 = UnpackVertexV2(vertices[off.vertexOffset + indices[off.indexOffset + h.prim * 3]]);
        const Vertex v1 = UnpackVertexV2(vertices[off.vertexOffset + indices[off.indexOffset + h.prim * 3 + 1]]);
        const Vertex v2 = UnpackVertexV2(vertices[off.vertexOffset + indices[off.indexOffset + h.prim * 3 + 2]]);
        const f3 n = normalize(cross(v1.Position - v0.Position, v2.Position - v1.Position));
        const GkMat4& M = node.worldTS;
        outNormal = F3(M.m[0][0] * n.x + M.m[1][0] * n.y + M.m[2][0] * n.z, M.m[0][1] * n.x + M.m[1][1] * n.y + M.m[2][1] * n.z,
                       M.m[0][2] * n.x + M.m[1][2] * n.y + M.m[2][2] * n.z);
        outT = h.t;
        outMat = FetchMaterialId(node, v0.MaterialIndex);
        return true;
    }
    void InsideGeometry(f3 origin, f3 rayDir, uint32_t& matId, float& distance) const {  // :85-110
        float t; f3 nrm; uint32_t mat;
        if (VoxTrace(origin, rayDir, 0.25f * 64, t, nrm, mat)) {
            distance = t;
            if (distance <= 0.25f) {
                matId = mat;
                if (dot(nrm, rayDir) > 0.0f || GetMaterial((int)mat).MaterialModel == GK_MATERIAL_DIFFUSE_LIGHT) distance = 0.0f;
            }
        }
    }
    float DetectDistance(f3 origin, f3 rayDir) const {  // :72-83
        float t; f3 nrm; uint32_t mat;
        return VoxTrace(origin, rayDir, 0.25f * 64, t, nrm, mat) ? t : 255.0f;
    }
    void VoxelizeCube(uint32_t gIdx) {  // :112-159
        const f3 origin = probe_origin(gIdx);
        GkVoxelData cube{};
        float distPY = 255.0f, distNY = 255.0f, distPX = 255.0f, distNX = 255.0f, distPZ = 255.0f, distNZ = 255.0f;
        const bool have = !instances.empty();
        if (have) {
            InsideGeometry(origin, F3(0, 1, 0), cube.matId, distPY);
            InsideGeometry(origin, F3(0, -1, 0), cube.matId, distNY);
            InsideGeometry(origin, F3(1, 0, 0), cube.matId, distPX);
            InsideGeometry(origin, F3(-1, 0, 0), cube.matId, distNX);
            InsideGeometry(origin, F3(0, 0, 1), cube.matId, distPZ);
            InsideGeometry(origin, F3(0, 0, -1), cube.matId, distNZ);
        }
        float minDist = fminf(fminf(fminf(distPY, distNY), fminf(distPX, distNX)), fminf(distPZ, distNZ));
        if (minDist > 254.0f && have) {
            const f3 dirs[8] = {F3(1, 1, 1), F3(-1, 1, 1), F3(-1, -1, 1), F3(-1, 1, 1), F3(1, 1, -1), F3(-1, 1, -1), F3(-1, -1, -1), F3(-1, 1, -1)};
            for (const f3& d : dirs) minDist = fminf(minDist, DetectDistance(origin, d));
        }
        distPY = clampf(distPY / 0.25f, 0.0f, 1.0f); distNY = clampf(distNY / 0.25f, 0.0f, 1.0f);
        distPX = clampf(distPX / 0.25f, 0.0f, 1.0f); distNX = clampf(distNX / 0.25f, 0.0f, 1.0f);
        distPZ = clampf(distPZ / 0.25f, 0.0f, 1.0f); distNZ = clampf(distNZ / 0.25f, 0.0f, 1.0f);
        const float inside = distPY * distNY * distPX * distNX * distPZ * distNZ;
        cube.distanceToSolid_gg_z01 = ((uint32_t)(minDist / 0.25f) & 0xFF) | (((uint32_t)(inside * 255.0f) & 0xFF) << 8) |
                                      (((uint32_t)(distPZ * 255.0f) & 0xFF) << 16) | (((uint32_t)(distNZ * 255.0f) & 0xFF) << 24);
        cube.distanceToSolid_x01_y01 = ((uint32_t)(distPX * 255.0f) & 0xFF) | (((uint32_t)(distNX * 255.0f) & 0xFF) << 8) |
                                       (((uint32_t)(distPY * 255.0f) & 0xFF) << 16) | (((uint32_t)(distNY * 255.0f) & 0xFF) << 24);
        voxels[gIdx] = cube;
    }
    void UpdatePageIndex() {  // FCPUPageIndex::UpdateData / GetPage, :499-545
        pages.assign((size_t)GK_PAGE_COUNT * GK_PAGE_COUNT, GkPageIndex{});
        for (uint32_t gIdx = 0; gIdx < voxels.size(); ++gIdx) {
            if (voxels[gIdx].matId == 0) continue;
            const f3 rel = probe_origin(gIdx) - F3(-512.0f, 0.0f, -512.0f);
            const int px = std::min(std::max((int)(rel.x / 16.0f), 0), GK_PAGE_COUNT - 1), pz = std::min(std::max((int)(rel.z / 16.0f), 0), GK_PAGE_COUNT - 1);
            pages[pz * GK_PAGE_COUNT + px].voxelCount += 1;
        }
    }
    // FGpuProbeGenerator::FaceTask, AmbientCube.slang:479-553
    void FaceTask(const GkUniformBufferObject& Camera, f3 origin, f3 basis, uint32_t iterate, uint32_t& DirectLight, uint32_t& IndirectLight, uint32_t& SkyVisOut,
                  uint32_t& SunVisOut, const GkAmbientCube* snapshot, uint32_t first, uint32_t count) const {
        static const f2 grid3x3[9] = {{-0.667f, -0.667f}, {0.0f, -0.667f}, {0.667f, -0.667f}, {-0.667f, 0.0f}, {0.0f, 0.0f},
                                      {0.667f, 0.0f},     {-0.667f, 0.667f}, {0.0f, 0.667f},  {0.667f, 0.667f}};
        static const f2 grid4x4[16] = {{-0.75f, -0.75f}, {-0.25f, -0.75f}, {0.25f, -0.75f}, {0.75f, -0.75f}, {-0.75f, -0.25f}, {-0.25f, -0.25f},
                                       {0.25f, -0.25f},  {0.75f, -0.25f},  {-0.75f, 0.25f}, {-0.25f, 0.25f}, {0.25f, 0.25f},   {0.75f, 0.25f},
                                       {-0.75f, 0.75f},  {-0.25f, 0.75f},  {0.25f, 0.75f},  {0.75f, 0.75f}};
        origin = origin + basis * 0.25f * 0.25f;
        f4 directColor = F4(0, 0, 0, 0), bounceColor = F4(0, 0, 0, 0);
        float skyVisibility = 0.0f;
        const f2 offset = grid3x3[iterate % 9u] * 0.25f;
        for (uint32_t i = 0; i < 16; i++) {
            const f2 g = grid4x4[i] + offset;
            f3 hemiVec = normalize(F3(g.x, g.y, 1.0f));
            f3 rayDir = AlignWithNormal(hemiVec, basis);
            Vertex hitVertex;
            GkNodeProxy hitNode;
            if (TraceRayVertex(origin, rayDir, 20.0f /* FAST_MAX_TRACE_DISTANCE */, hitVertex, hitNode)) {
                GkMaterial hitMaterial = GetMaterial((int)hitVertex.MaterialIndex);
                f4 outAlbedo = F4(hitMaterial.Diffuse.x, hitMaterial.Diffuse.y, hitMaterial.Diffuse.z, hitMaterial.Diffuse.w);
                if (hitMaterial.DiffuseTextureId >= 0) outAlbedo = outAlbedo * sample_texture((uint32_t)hitMaterial.DiffuseTextureId, hitVertex.TexCoord);
                bounceColor = bounceColor + outAlbedo * interpolateAmbientCubesT(hitVertex.Position, hitVertex.Normal, true, snapshot, first, count) * 1.25f;
            } else {
                directColor = directColor + SampleIBLRough(Camera.SkyIdx, rayDir, Camera.SkyRotation) * (Camera.HasSky ? Camera.SkyIntensity : 0.0f);
                skyVisibility += 1.0f;
            }
        }
        directColor = directColor / 16.0f;
        bounceColor = bounceColor / 16.0f;
        if (Camera.LightCount > 0 && !lights.empty()) {
            const GkLightObject& light = lights[0];
            const GkMaterial lm = GetMaterial((int)light.lightMatIdx);
            const f4 lightPower = F4(lm.Diffuse.x, lm.Diffuse.y, lm.Diffuse.z, lm.Diffuse.w);
            const f3 lightPos = lerp3(F3(light.p1.x, light.p1.y, light.p1.z), F3(light.p3.x, light.p3.y, light.p3.z), 0.5f);
            // TraceSegment(origin, lightPos, CUBE_UNIT * 0.5), Shading.slang:759-783
            const f3 dir = lightPos - origin;
            const float len = length(dir);
            const float lightAtten = occluded(origin, F3(dir.x / len, dir.y / len, dir.z / len), 0.25f * 0.5f, len - 0.25f * 0.5f) ? 0.0f : 1.0f;
            const f3 lightDir = normalize(lightPos - origin);
            const float ndotl = clampf(dot(basis, lightDir), 0.0f, 1.0f);
            const float distance = length(lightPos - origin);
            const float attenuation = ndotl * light.normal_area.w / (distance * distance * 3.14159f);
            directColor = directColor + lightPower * attenuation * lightAtten;
        }
        if (Camera.HasSun) {
            const f3 sunDir = F3(Camera.SunDirection.x, Camera.SunDirection.y, Camera.SunDirection.z);
            const float sunAtten = occluded(origin, sunDir, EPS, PT_MAX_TRACE_DISTANCE) ? 0.0f : 1.0f;
            const float ndotl = clampf(dot(basis, sunDir), 0.0f, 1.0f);
            SunVisOut = sunAtten > 0.0f ? 1 : 0;
            const f4 sunColor = F4(Camera.SunColor.x, Camera.SunColor.y, Camera.SunColor.z, Camera.SunColor.w);
            directColor = directColor + sunColor * sunAtten * ndotl * 1.0f * 0.25f;
        }
        const float currWeight = 0.125f;
        const float a = (float)SkyVisOut, b = 255.0f * skyVisibility / 16.0f;
        SkyVisOut = (uint32_t)(a + (b - a) * currWeight);
        DirectLight = LerpPackedColorAlt(DirectLight, directColor, currWeight);
        IndirectLight = LerpPackedColorAlt(IndirectLight, bounceColor, currWeight);
    }
    // FGpuProbeGenerator::RenderAmbientCubeOnly, AmbientCube.slang:673-692
    void BakeCube(const GkUniformBufferObject& Camera, uint32_t CubeIdx, const GkAmbientCube* snapshot, uint32_t first, uint32_t count) {
        if ((voxels[CubeIdx].distanceToSolid_gg_z01 & 0xFFu) >= 8u) return;
        const f3 origin = probe_origin(CubeIdx);
        const uint32_t iterate = voxels[CubeIdx].age;
        voxels[CubeIdx].age = iterate + 1;
        GkAmbientCube c = snapshot[CubeIdx - first];
        uint32_t s0[4] = {c.skyVisibility_pznzpyny & 0xFF, (c.skyVisibility_pznzpyny >> 8) & 0xFF, (c.skyVisibility_pznzpyny >> 16) & 0xFF, c.skyVisibility_pznzpyny >> 24};
        uint32_t s1[4] = {c.skyVisibility_pxnxs0s1 & 0xFF, (c.skyVisibility_pxnxs0s1 >> 8) & 0xFF, (c.skyVisibility_pxnxs0s1 >> 16) & 0xFF, c.skyVisibility_pxnxs0s1 >> 24};
        uint32_t sunvis = 0;
        FaceTask(Camera, origin, F3(0, 1, 0), iterate, c.PosY_D, c.PosY, s0[2], s1[2], snapshot, first, count);
        FaceTask(Camera, origin, F3(0, -1, 0), iterate, c.NegY_D, c.NegY, s0[3], sunvis, snapshot, first, count);
        FaceTask(Camera, origin, F3(1, 0, 0), iterate, c.PosX_D, c.PosX, s1[0], sunvis, snapshot, first, count);
        FaceTask(Camera, origin, F3(-1, 0, 0), iterate, c.NegX_D, c.NegX, s1[1], sunvis, snapshot, first, count);
        FaceTask(Camera, origin, F3(0, 0, 1), iterate, c.PosZ_D, c.PosZ, s0[0], sunvis, snapshot, first, count);
        FaceTask(Camera, origin, F3(0, 0, -1), iterate, c.NegZ_D, c.NegZ, s0[1], sunvis, snapshot, first, count);
        c.skyVisibility_pznzpyny = (s0[0] & 0xFF) | ((s0[1] & 0xFF) << 8) | ((s0[2] & 0xFF) << 16) | ((s0[3] & 0xFF) << 24);
        c.skyVisibility_pxnxs0s1 = (s1[0] & 0xFF) | ((s1[1] & 0xFF) << 8) | ((s1[2] & 0xFF) << 16) | ((s1[3] & 0xFF) << 24);
        cubes[CubeIdx] = c;
    }

    f4 interpolateAmbientCubesV2(f3 inPos, f3 normal) const {
        return interpolateAmbientCubesT(inPos, normal, false, nullptr, 0, 0);
    }
    // di: DIAmbientCubeSampler instead of FullAmbientCubeSampler; snapshot: pre-dispatch copies of cubes[first .. first + count)
    // (the bake reads the range it is updating from there, see gk_shading.cuh SnapshotCubes).
    f4 interpolateAmbientCubesT(f3 inPos, f3 normal, bool di, const GkAmbientCube* snapshot, uint32_t first, uint32_t count) const {
        const f3 CUBE_OFFSET = F3(-GK_CUBE_SIZE_XY / 2, -1.375f, -GK_CUBE_SIZE_XY / 2) * 0.25f;
        f3 nearpos = (inPos - CUBE_OFFSET) * (1.0f / 0.25f);
        if (nearpos.x < 0 || nearpos.y < 0 || nearpos.z < 0 || nearpos.x > GK_CUBE_SIZE_XY - 2 || nearpos.y > GK_CUBE_SIZE_Z - 2 ||
            nearpos.z > GK_CUBE_SIZE_XY - 2)
            return F4(0, 0, 0, 1);
        if (cubes.empty()) return F4(0, 0, 0, 0);  // zero probe state: every probe is skipped at :327
        int bx = (int)floorf(nearpos.x), by = (int)floorf(nearpos.y), bz = (int)floorf(nearpos.z);
        f3 fr = F3(nearpos.x - floorf(nearpos.x), nearpos.y - floorf(nearpos.y), nearpos.z - floorf(nearpos.z));
        float totalWeight = 0.0f;
        f4 result = F4(0, 0, 0, 0);
        for (int i = 0; i < 8; i++) {
            int ox = i & 1, oy = (i >> 1) & 1, oz = (i >> 2) & 1;
            int idx = (by + oy) * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY + (bz + oz) * GK_CUBE_SIZE_XY + (bx + ox);
            const uint32_t rel = (uint32_t)idx - first;
            const GkAmbientCube& cube = (snapshot && rel < count) ? snapshot[rel] : cubes[idx];
            const GkVoxelData& voxel = voxels[idx];
            uint32_t u0 = voxel.distanceToSolid_gg_z01, u1 = voxel.distanceToSolid_x01_y01;
            float d0y = float((u0 >> 8) & 0xFF) / 255.0f, distPZ = float((u0 >> 16) & 0xFF) / 255.0f, distNZ = float((u0 >> 24) & 0xFF) / 255.0f;
            float distPX = float(u1 & 0xFF) / 255.0f, distNX = float((u1 >> 8) & 0xFF) / 255.0f, distPY = float((u1 >> 16) & 0xFF) / 255.0f,
                  distNY = float((u1 >> 24) & 0xFF) / 255.0f;
            if (d0y < 0.01f) continue;
            f3 probeToLocal = fr - F3((float)ox, (float)oy, (float)oz);
            float distanceToProbe = length(probeToLocal);
            f3 dir = normalize(probeToLocal);
            float hitLength = sqrtf(fmaxf(dir.x, 0.0f)) * distPX + sqrtf(fmaxf(-dir.x, 0.0f)) * distNX + sqrtf(fmaxf(dir.y, 0.0f)) * distPY +
                              sqrtf(fmaxf(-dir.y, 0.0f)) * distNY + sqrtf(fmaxf(dir.z, 0.0f)) * distPZ + sqrtf(fmaxf(-dir.z, 0.0f)) * distNZ;
            if (distanceToProbe > hitLength + 0.05f) continue;
            float wx = ox == 0 ? (1.0f - fr.x) : fr.x, wy = oy == 0 ? (1.0f - fr.y) : fr.y, wz = oz == 0 ? (1.0f - fr.z) : fr.z;
            float weight = wx * wy * wz;
            result = result + (di ? sampleAmbientCubeHL2_DI(cube, normal) : sampleAmbientCubeHL2_Full(cube, normal)) * weight;
            totalWeight += weight;
        }
        return totalWeight > 0.0f ? result / totalWeight : F4(0, 0, 0, 0);
    }
};

struct GkoFrame {
    uint32_t w = 0, h = 0;
    // fp16x4 images
    std::vector<uint16_t> accum_diffuse, accum_specular, accum_albedo, prev_diffuse, prev_specular, prev_albedo, single_diffuse,
        single_specular, albedo, normal, denoised;
    std::vector<uint32_t> objectid0, objectid1, primitiveid;
    std::vector<float> motion;  // RG32F
    std::vector<float> prev_depth;
    std::vector<uint16_t> moment;
    std::vector<uint32_t> vis;    // RT_MINIGBUFFER: the rasterised visibility buffer (primary_visibility = 1)
    bool use_vis = false;
    std::vector<float> radiance;  // multi-GPU split: FinalColor | FinalReflection | direct planes (float4 per pixel)
    std::atomic<uint64_t> primary_rays{0}, closest_rays{0}, any_rays{0};
    double rt_seconds = 0;

    bool inb(int x, int y) const { return x >= 0 && y >= 0 && x < (int)w && y < (int)h; }
    // robust image access: OOB loads return 0, OOB stores are dropped
    f4 load4(const std::vector<uint16_t>& img, int x, int y) const {
        if (!inb(x, y)) return F4(0, 0, 0, 0);
        const uint16_t* p = &img[((size_t)y * w + x) * 4];
        return F4(h2f(p[0]), h2f(p[1]), h2f(p[2]), h2f(p[3]));
    }
    void store4(std::vector<uint16_t>& img, int x, int y, f4 v) {
        if (!inb(x, y)) return;
        uint16_t* p = &img[((size_t)y * w + x) * 4];
        p[0] = f2h(v.x); p[1] = f2h(v.y); p[2] = f2h(v.z); p[3] = f2h(v.w);
    }
    void store3(std::vector<uint16_t>& img, int x, int y, f3 v) {  // image[ipos].rgb = v keeps alpha
        if (!inb(x, y)) return;
        uint16_t* p = &img[((size_t)y * w + x) * 4];
        p[0] = f2h(v.x); p[1] = f2h(v.y); p[2] = f2h(v.z);
    }
    uint32_t loadu(const std::vector<uint32_t>& img, int x, int y) const { return inb(x, y) ? img[(size_t)y * w + x] : 0u; }
    f2 loadmv(int x, int y) const { return inb(x, y) ? f2{motion[((size_t)y * w + x) * 2], motion[((size_t)y * w + x) * 2 + 1]} : f2{0, 0}; }
};

namespace {

// ----------------------------------------------------------------------------------------------- hit resolve
// GeneralFunc.slang:6-56
Vertex get_material_dataV2(const GkoScene& S, uint32_t nodeIdxPlus1, uint32_t primitive_index, f3 ray_origin, f3 ray_direction) {
    const GkNodeProxy& proxy = S.nodes[nodeIdxPlus1 - 1];
    const GkModelData& model = S.offsets[proxy.modelId];
    Vertex result;
    f3 positions[3], normals[3], tangents[3], localNormals[3];
    f4 localTangents[3];
    f2 tex_coords[3];
    uint32_t matid = 0;
    for (int i = 0; i != 3; ++i) {
        const Vertex v = UnpackVertexV2(S.vertices[model.vertexOffset + S.indices[model.indexOffset + primitive_index * 3 + i]]);
        positions[i] = xyz(mul(proxy.worldTS, F4(v.Position.x, v.Position.y, v.Position.z, 1)));
        normals[i] = xyz(mul(proxy.worldTS, F4(v.Normal.x, v.Normal.y, v.Normal.z, 0)));
        tangents[i] = xyz(mul(proxy.worldTS, F4(v.Tangent.x, v.Tangent.y, v.Tangent.z, 0)));
        localNormals[i] = v.Normal;
        localTangents[i] = v.Tangent;
        tex_coords[i] = v.TexCoord;
        if (i == 0) matid = v.MaterialIndex;
    }
    f3 edges[2] = {positions[1] - positions[0], positions[2] - positions[0]};
    f3 ray_cross_edge_1 = cross(ray_direction, edges[1]);
    float rcp_det_edges_direction = 1.0f / dot(edges[0], ray_cross_edge_1);
    f3 ray_to_0 = ray_origin - positions[0];
    float det_0_dir_edge_1 = dot(ray_to_0, ray_cross_edge_1);
    float by = rcp_det_edges_direction * det_0_dir_edge_1;
    f3 edge_0_cross_0 = cross(edges[0], ray_to_0);
    float det_dir_edge_0_0 = dot(ray_direction, edge_0_cross_0);
    float bz = -rcp_det_edges_direction * det_dir_edge_0_0;
    float bx = 1.0f - (by + bz);
    auto mix3 = [&](const f3* a) { return a[0] * bx + (a[1] * by + a[2] * bz); };
    result.Position = mix3(positions);
    result.Normal = normalize(mix3(normals));
    f3 tg = normalize(mix3(tangents));
    result.Tangent = F4(tg.x, tg.y, tg.z, 0);
    result.TexCoord = tex_coords[0] * bx + (tex_coords[1] * by + tex_coords[2] * bz);
    result.LocalNormal = mix3(localNormals);
    result.LocalTangent = localTangents[0] * bx + (localTangents[1] * by + localTangents[2] * bz);
    result.MaterialIndex = matid;
    return result;
}

struct Renderer {  // FPathTracingRendererV2, Shading.slang:995-1261, configured as in Core.PathTracing.comp.slang:15-20
    const GkoScene& S;
    GkoFrame& F;
    const GkUniformBufferObject& Camera;
    float ExitProbability = 0.5f;
    bool ExitAfterFirst = false;
    float HitNormalOffset = 0.001f;
    int SampleDownscale = 1;

    int px = 0, py = 0;
    u4 RandomSeed_{};
    Vertex hitPrimaryVertex_;
    GkNodeProxy hitPrimaryNode_{};
    f3 primaryRayDir_{};
    float primaryRayLength_ = 0;
    f2 primaryPixelOffset_{};
    f4 primaryAlbedo_{};
    uint32_t primaryPrim_ = 0xFFFFFFFFu;
    uint64_t n_primary = 0, n_closest = 0, n_any = 0;

    Renderer(const GkoScene& s, GkoFrame& f, const GkUniformBufferObject& c) : S(s), F(f), Camera(c) { ExitAfterFirst = c.FastGather != 0; }

    // FHardwareRayTracerV2::TraceRay, Shading.slang:784-826
    bool TraceRay(f3 rayOrigin, f3 rayDir, float maxDistance, Vertex& outVertex, GkNodeProxy& outNode) {
        n_closest++;
        return S.TraceRayVertex(rayOrigin, rayDir, maxDistance, outVertex, outNode);
    }
    bool TraceOcclusion(f3 rayOrigin, f3 rayDir) {  // Shading.slang:737-757
        n_any++;
        return S.occluded(rayOrigin, rayDir, EPS, PT_MAX_TRACE_DISTANCE);
    }

    // Primary visibility: the reference reads a visibility buffer rasterised from the same camera (Rast.VisibilityPass.vert.slang:9-23,
    // read at Shading.slang:348-349). The rasteriser samples coverage and depth at the pixel CENTRE, so the stand-in for a GPU without one
    // is the closest hit of the ray through the pixel centre; the shading ray that TracePrimaryRay then intersects with that triangle's
    // plane stays the pixel-corner ray of the text (:350-356). Returns (node index + 1, primitive) in the vis-buffer convention, 0 = none.
    void CastVisibility(int x, int y, f3 origin, uint32_t& node1, uint32_t& prim) {
        if (F.use_vis) {  // Shading.slang:348-349: unpack the visibility buffer the raster pass wrote
            const uint32_t v = F.vis[(size_t)y * F.w + x];
            node1 = (v >> 17) & 0x7FFFu;
            prim = v & 0x1FFFFu;
            return;
        }
        f2 uv = f2{((float)x + 0.5f) / (float)F.w * 2.0f - 1.0f, ((float)y + 0.5f) / (float)F.h * 2.0f - 1.0f};
        f4 target = mul(Camera.ProjectionInverse, F4(uv.x, uv.y, 1, 1));
        f3 tn = normalize(xyz(target));
        f4 dir = mul(Camera.ModelViewInverse, F4(tn.x, tn.y, tn.z, 0));
        f3 ray_dir = normalize(xyz(dir));
        n_primary++;
        Hit h = S.closest(origin, ray_dir, EPS, PRIMARY_T_MAX);
        if (h.inst == 0xFFFFFFFFu) { node1 = 0; prim = 0; return; }
        node1 = h.inst + 1;
        prim = h.prim;
    }

    // FVisibilityBufferRayCasterV2::TracePrimaryRay, Shading.slang:338-494
    bool TracePrimaryRay(Vertex& outVertex, GkNodeProxy& outNode, f3& outRayDir, float& outRayLength, f2& outPixelOffset) {
        const int isx = (int)F.w, isy = (int)F.h;
        f4 origin = mul(Camera.ModelViewInverse, F4(0, 0, 0, 1));
        uint32_t vbx, vby;
        CastVisibility(px, py, xyz(origin), vbx, vby);
        f2 uv = f2{(float)px / (float)isx * 2.0f - 1.0f, (float)py / (float)isy * 2.0f - 1.0f};
        f4 target = mul(Camera.ProjectionInverse, F4(uv.x, uv.y, 1, 1));
        f3 tn = normalize(xyz(target));
        f4 dir = mul(Camera.ModelViewInverse, F4(tn.x, tn.y, tn.z, 0));
        f3 ray_dir = normalize(xyz(dir));
        if (vbx == 0) { outRayDir = ray_dir; return false; }
        Vertex initialVertex = get_material_dataV2(S, vbx, vby, xyz(origin), ray_dir);
        float vertexDistance = length(initialVertex.Position - xyz(origin));
        float cocRadius = 0.0f;
        if (fabsf(vertexDistance - Camera.FocusDistance) > 0.001f)
            cocRadius = (Camera.Aperture * fabsf(vertexDistance - Camera.FocusDistance)) / vertexDistance;
        f2 pixelOffset = f2{0, 0};
        if (cocRadius > 0.001f) {
            f2 diskSample = RandomInUnitDisk(RandomSeed_);
            f3 right = normalize(cross(ray_dir, F3(0, 1, 0)));
            f3 circleEdge = initialVertex.Position + right * cocRadius;
            f4 centerProj = mul(Camera.ViewProjection, F4(initialVertex.Position.x, initialVertex.Position.y, initialVertex.Position.z, 1.0f));
            f4 edgeProj = mul(Camera.ViewProjection, F4(circleEdge.x, circleEdge.y, circleEdge.z, 1.0f));
            f2 cp = f2{centerProj.x / centerProj.w, centerProj.y / centerProj.w}, ep = f2{edgeProj.x / edgeProj.w, edgeProj.y / edgeProj.w};
            float screenSpaceRadius = length(ep - cp);
            pixelOffset = f2{diskSample.x * screenSpaceRadius * (float)isx * 0.5f, diskSample.y * screenSpaceRadius * (float)isy * 0.5f};
        }
        outPixelOffset = pixelOffset;
        int ox = px + (int)pixelOffset.x, oy = py + (int)pixelOffset.y;
        ox = std::min(std::max(ox, 0), isx - 1);
        oy = std::min(std::max(oy, 0), isy - 1);
        uint32_t fvx = vbx, fvy = vby;
        if (ox != px || oy != py) CastVisibility(ox, oy, xyz(origin), fvx, fvy);
        f2 finalUV = f2{(float)ox / (float)isx * 2.0f - 1.0f, (float)oy / (float)isy * 2.0f - 1.0f};
        f4 finalTarget = mul(Camera.ProjectionInverse, F4(finalUV.x, finalUV.y, 1, 1));
        f3 ftn = normalize(xyz(finalTarget));
        f4 finalDir = mul(Camera.ModelViewInverse, F4(ftn.x, ftn.y, ftn.z, 0));
        f3 final_ray_dir = normalize(xyz(finalDir));
        auto useInitial = [&]() {
            outVertex.Position = initialVertex.Position;
            outVertex.Normal = normalize(initialVertex.Normal);
            outNode = S.nodes[vbx - 1];
            outVertex.MaterialIndex = FetchMaterialId(outNode, initialVertex.MaterialIndex);
            outVertex.TexCoord = initialVertex.TexCoord;
            outVertex.Tangent = initialVertex.Tangent;
            outVertex.LocalNormal = initialVertex.LocalNormal;
            outVertex.LocalTangent = initialVertex.LocalTangent;
            outRayDir = ray_dir;
            primaryPrim_ = vby;
        };
        if (fvx == 0) {
            useInitial();
        } else {
            Vertex finalVertex = get_material_dataV2(S, fvx, fvy, xyz(origin), final_ray_dir);
            float distanceToFocus = fabsf(vertexDistance - Camera.FocusDistance);
            float focusThreshold = Camera.FocusDistance * 0.1f;
            bool shouldCancelOffset = false;
            if (distanceToFocus < focusThreshold) {
                float finalVertexDistance = length(finalVertex.Position - xyz(origin));
                float depthDifference = vertexDistance - finalVertexDistance;
                float depthThreshold = Camera.FocusDistance * 0.05f;
                if (depthDifference > depthThreshold) shouldCancelOffset = true;
            }
            if (shouldCancelOffset) {
                useInitial();
            } else {
                outNode = S.nodes[fvx - 1];
                outVertex.Position = finalVertex.Position;
                outVertex.Normal = normalize(finalVertex.Normal);
                outVertex.MaterialIndex = FetchMaterialId(outNode, finalVertex.MaterialIndex);
                outVertex.TexCoord = finalVertex.TexCoord;
                outVertex.Tangent = finalVertex.Tangent;
                outVertex.LocalNormal = finalVertex.LocalNormal;
                outVertex.LocalTangent = finalVertex.LocalTangent;
                outRayDir = final_ray_dir;
                primaryPrim_ = fvy;
            }
        }
        outRayLength = length(outVertex.Position - xyz(origin));
        return true;
    }

    // Common.CalculateMotionVectorV2, Shading.slang:50-72 (side effect on MOTIONMOMENT)
    f2 CalculateMotionVectorV2(const GkNodeProxy& node, const Vertex& v) {
        f4 P = F4(v.Position.x, v.Position.y, v.Position.z, 1);
        f4 currFrameHPos = mul(Camera.ViewProjectionUnJit, P);
        f2 currfpos = f2{currFrameHPos.x / currFrameHPos.w * 0.5f, currFrameHPos.y / currFrameHPos.w * 0.5f};
        f4 prevFrameHPos = mul(matmul(Camera.PrevViewProjectionUnJit, node.combinedPrevTS), P);
        f2 prevfpos = f2{prevFrameHPos.x / prevFrameHPos.w * 0.5f, prevFrameHPos.y / prevFrameHPos.w * 0.5f};
        f4 prevPos = mul(node.combinedPrevTS, P);
        uint16_t& mm = F.moment[(size_t)py * F.w + px];
        if (length(v.Position - xyz(prevPos)) > 0.02f) mm = (uint16_t)Camera.TemporalFrames;
        else mm = mm > 0 ? mm - 1 : 0;
        return prevfpos - currfpos;
    }

    // Shading.slang:74-80
    static f3 mikkTSpace(f3 vNt, f3 normal, f4 tangent) {
        vNt = vNt * 2.0f - F3(1.0f, 1.0f, 1.0f);
        f3 vB = cross(normal, xyz(tangent)) * tangent.w;
        return normalize(xyz(tangent) * vNt.x + vB * vNt.y + normal * vNt.z);
    }
    // Common.FetchGBufferV2, Shading.slang:82-112
    void FetchGBufferV2(Vertex& inVertex, const GkMaterial& m, const GkNodeProxy& proxy, f4& outAlbedo, f4& outNormal) {
        outAlbedo = F4(m.Diffuse.x, m.Diffuse.y, m.Diffuse.z, m.Diffuse.w);
        if (m.DiffuseTextureId >= 0) outAlbedo = outAlbedo * S.sample_texture((uint32_t)m.DiffuseTextureId, inVertex.TexCoord);
        float roughness = m.Fuzziness;
        if (m.MRATextureId >= 0) roughness = roughness * S.sample_texture((uint32_t)m.MRATextureId, inVertex.TexCoord).y;
        if (m.NormalTextureId >= 0) {  // normal map in the local frame, then to world space through the node matrix
            f4 t = S.sample_texture((uint32_t)m.NormalTextureId, inVertex.TexCoord);
            f3 nl = mikkTSpace(F3(t.x, t.y, t.z), inVertex.LocalNormal, inVertex.LocalTangent);
            f3 nw = normalize(xyz(mul(proxy.worldTS, F4(nl.x, nl.y, nl.z, 0.0f))));
            outNormal = F4(nw.x, nw.y, nw.z, roughness);
            inVertex.Normal = nw;
        } else {
            outNormal = F4(inVertex.Normal.x, inVertex.Normal.y, inVertex.Normal.z, roughness);
        }
    }

    void Init(int x, int y, uint32_t frameSeed) {  // Shading.slang:1226-1241
        px = x;
        py = y;
        RandomSeed_ = InitRandomSeed((uint32_t)x, (uint32_t)y, frameSeed);
    }

    f4 direct_{};  // what PrimaryHit / the emissive early-out / DirectIlluminate put into SINGLE_DIFFUSE un-normalised

    bool PrimaryHit(bool write_gbuffer) {  // Shading.slang:1187-1223
        if (!TracePrimaryRay(hitPrimaryVertex_, hitPrimaryNode_, primaryRayDir_, primaryRayLength_, primaryPixelOffset_)) {
            f4 skyColor = Camera.HasSky ? S.SampleIBLV2(Camera.SkyIdx, primaryRayDir_, Camera.SkyRotation, 0) * Camera.SkyIntensity : F4(0, 0, 0, 0);
            direct_ = skyColor;
            if (write_gbuffer) {
                size_t o = (size_t)py * F.w + px;
                F.motion[o * 2] = 0; F.motion[o * 2 + 1] = 0;
                F.store4(F.albedo, px, py, F4(1, 1, 1, 1));
                F.store4(F.normal, px, py, F4(0, 1, 0, 1));
                F.objectid0[o] = GK_OBJECTID_MISS;
                F.primitiveid[o] = 0xFFFFFFFFu;
            }
            return false;
        }
        f4 clipPos = mul(Camera.ViewProjection, F4(hitPrimaryVertex_.Position.x, hitPrimaryVertex_.Position.y, hitPrimaryVertex_.Position.z, 1.0f));
        float ndcDepth = clipPos.z / clipPos.w;
        GkMaterial material = S.GetMaterial((int)hitPrimaryVertex_.MaterialIndex);
        f4 gbuffer = F4(0, 0, 0, 0);
        if (write_gbuffer) {
            f2 motion = CalculateMotionVectorV2(hitPrimaryNode_, hitPrimaryVertex_);
            FetchGBufferV2(hitPrimaryVertex_, material, hitPrimaryNode_, primaryAlbedo_, gbuffer);
            size_t o = (size_t)py * F.w + px;
            F.store4(F.albedo, px, py, primaryAlbedo_);
            F.store4(F.normal, px, py, gbuffer);
            F.objectid0[o] = hitPrimaryNode_.instanceId;
            F.primitiveid[o] = primaryPrim_;
            F.motion[o * 2] = motion.x * (float)F.w;
            F.motion[o * 2 + 1] = motion.y * (float)F.h;
            F.prev_depth[o] = ndcDepth;
        } else {
            FetchGBufferV2(hitPrimaryVertex_, material, hitPrimaryNode_, primaryAlbedo_, gbuffer);
        }
        if (material.MaterialModel != GK_MATERIAL_DIELECTRIC) hitPrimaryVertex_.Position = hitPrimaryVertex_.Position - primaryRayDir_ * 0.03f;
        return true;
    }

    // Shading.slang:1003-1082
    bool GetRayColor(Vertex& inVertex, f3& rayDir, f4& rayColor, bool& hitReflect, bool& hitMetal) {
        GkMaterial mat = S.GetMaterial((int)inVertex.MaterialIndex);
        const float startPosOffset = mat.MaterialModel == GK_MATERIAL_DIELECTRIC ? 0.0f : 1.0f;
        float roughness = mat.Fuzziness;
        const float dotValue = dot(rayDir, inVertex.Normal);
        const bool BackFace = dotValue > 0;
        const f3 outwardNormal = BackFace ? -inVertex.Normal : inVertex.Normal;
        const float niOverNt = BackFace ? mat.RefractionIndex2 : (1 / mat.RefractionIndex2);
        const float cosine = dotValue > 0 ? mat.RefractionIndex * dotValue : -dotValue;
        const float reflectProb = Schlick(cosine, mat.RefractionIndex);
        const float metalProb = mat.Metalness;
        bool chanceReflect = (RandomFloat(RandomSeed_) < reflectProb);
        bool chanceMetal = (RandomFloat(RandomSeed_) < metalProb);
        bool chanceGGX = chanceReflect || chanceMetal;
        const f3 trace_next = chanceGGX ? reflect3(rayDir, outwardNormal) : outwardNormal;
        // ?: short-circuits for a scalar condition (HLSL 2021 / Slang): exactly one RNG draw
        f3 trace_dir = chanceGGX ? ggxSampling(RandomSeed_, sqrtf(roughness), trace_next) : AlignWithNormal(RandomInHemiSphere1(RandomSeed_), trace_next);
        hitReflect = chanceGGX;
        hitMetal = chanceMetal;
        if (mat.MaterialModel == GK_MATERIAL_DIELECTRIC) {
            if (!chanceReflect) trace_dir = refract3(rayDir, outwardNormal, niOverNt);
        }
        rayDir = trace_dir;
        GkNodeProxy hitNode;
        if (TraceRay(inVertex.Position + inVertex.Normal * HitNormalOffset * startPosOffset, trace_dir, PT_MAX_TRACE_DISTANCE, inVertex, hitNode)) {
            GkMaterial hitMat = S.GetMaterial((int)inVertex.MaterialIndex);
            f4 outAlbedo = F4(hitMat.Diffuse.x, hitMat.Diffuse.y, hitMat.Diffuse.z, hitMat.Diffuse.w);
            if (hitMat.DiffuseTextureId >= 0) outAlbedo = outAlbedo * S.sample_texture((uint32_t)hitMat.DiffuseTextureId, inVertex.TexCoord);
            if (hitMat.MaterialModel == GK_MATERIAL_DIFFUSE_LIGHT || !chanceReflect) rayColor = rayColor * outAlbedo;
            if (BackFace && mat.MaterialModel != GK_MATERIAL_DIELECTRIC) {
                rayColor = F4(0.5f, 0.5f, 0.5f, 0);
                return true;
            }
            if (hitMat.MaterialModel == GK_MATERIAL_DIFFUSE_LIGHT) return true;
            return false;
        } else {
            f4 skyColor = Camera.HasSky ? S.SampleIBLV2(Camera.SkyIdx, trace_dir, Camera.SkyRotation, 0) * Camera.SkyIntensity : F4(0, 0, 0, 0);
            rayColor = rayColor * skyColor;
            return true;
        }
    }

    // FHardwareDirectIlluminatorV2::DirectIlluminate, Shading.slang:886-908
    f4 DirectIlluminate(f3 position, f3 normal) {
        float shadow = 0.0f;
        const f3 lightVector = F3(Camera.SunDirection.x, Camera.SunDirection.y, Camera.SunDirection.z);
        const float d = fmaxf(dot(lightVector, normalize(normal)), 0.0f) * M_1_PI_F;
        if (Camera.HasSun) {
            const f3 lightVectorCone = AlignWithNormal(RandomInCone(RandomSeed_, COS_QUARTER_DEGREE), lightVector);
            shadow = 1;
            if (TraceOcclusion(position, lightVectorCone)) shadow = 0;
        }
        return F4(Camera.SunColor.x, Camera.SunColor.y, Camera.SunColor.z, Camera.SunColor.w) * d * shadow;
    }

    // The sample loop of FPathTracingRendererV2::Render, Shading.slang:1085-1173: samples [s0, s1) of `sample`.
    // Returns false when the primary hit is emissive (direct_ = mat.Diffuse, :1093-1099).
    bool RenderSamples(int s0, int s1, f4& FinalColor, f4& FinalReflection) {
        GkMaterial mat = S.GetMaterial((int)hitPrimaryVertex_.MaterialIndex);
        if (mat.MaterialModel == GK_MATERIAL_DIFFUSE_LIGHT) {
            direct_ = F4(mat.Diffuse.x, mat.Diffuse.y, mat.Diffuse.z, mat.Diffuse.w);
            return false;
        }
        for (int i = s0; i < s1; i++) {
            f4 RayColor = F4(1, 1, 1, 1);
            f4 eye = mul(Camera.ModelViewInverse, F4(0, 0, 0, 1));
            f3 ray_dir = normalize(hitPrimaryVertex_.Position - xyz(eye));
            f3 direction = ray_dir;
            Vertex vertexStart = hitPrimaryVertex_;
            uint32_t maxBounces = mat.MaterialModel == GK_MATERIAL_DIELECTRIC ? Camera.MaxNumberOfBounces : Camera.NumberOfBounces;
            bool hitReflect = false, hitMetal = false;
            bool exitFirst = GetRayColor(vertexStart, direction, RayColor, hitReflect, hitMetal);
            if (!exitFirst) {
                if (ExitAfterFirst && mat.MaterialModel != GK_MATERIAL_DIELECTRIC) {
                    RayColor = RayColor * S.interpolateAmbientCubesV2(vertexStart.Position, vertexStart.Normal);
                } else {
                    for (uint32_t b = 1; b < maxBounces; ++b) {
                        bool hitReflectDontCare = false, hitMetalDontCare = false;
                        // the reference passes hitMetalDontCare twice (:1130), so hitReflectDontCare stays false
                        if (GetRayColor(vertexStart, direction, RayColor, hitMetalDontCare, hitMetalDontCare)) break;
                        if (!hitReflectDontCare && Camera.HasSun && (RandomFloat(RandomSeed_) < 0.5f)) {
                            const f3 lightVector = F3(Camera.SunDirection.x, Camera.SunDirection.y, Camera.SunDirection.z);
                            const f3 lightVectorCone = AlignWithNormal(RandomInCone(RandomSeed_, COS_QUARTER_DEGREE), lightVector);
                            if (!TraceOcclusion(vertexStart.Position + vertexStart.Normal * 0.0f, lightVectorCone)) {
                                RayColor = RayColor * F4(Camera.SunColor.x, Camera.SunColor.y, Camera.SunColor.z, Camera.SunColor.w);
                                break;
                            }
                        }
                        const bool earlyExit = b > 0 && (mat.MaterialModel != GK_MATERIAL_DIELECTRIC) && (RandomFloat(RandomSeed_) < ExitProbability);
                        if (b == maxBounces - 1 || earlyExit) {
                            RayColor = RayColor * S.interpolateAmbientCubesV2(vertexStart.Position, vertexStart.Normal);
                            break;
                        }
                    }
                }
            }
            if (hitMetal) RayColor = RayColor * primaryAlbedo_;
            if (hitReflect) FinalReflection = FinalReflection + RayColor;
            else FinalColor = FinalColor + RayColor;
        }
        return true;
    }
};

// Core.PathTracing.comp.slang:5-31 for one pixel. spp_world > 1 simulates every rank of the sample split in rank order.
void RenderPixel(const GkoScene& S, GkoFrame& F, const GkUniformBufferObject& ubo, const GkcuFrameParams& P, int x, int y) {
    const uint32_t world = std::max(1u, P.spp_world);
    size_t o = (size_t)y * F.w + x;
    const bool single_rank = P.skip_post != 0;  // trace only rank P.spp_rank and leave sums in F.radiance
    // Util.BufferClear.comp.slang:5-14
    F.store4(F.single_diffuse, x, y, F4(0, 0, 0, 0));
    F.store4(F.single_specular, x, y, F4(0, 0, 0, 0));
    F.prev_depth[o] = 100000.0f;
    f4 FinalColor = F4(0, 0, 0, 0), FinalReflection = F4(0, 0, 0, 0), direct = F4(0, 0, 0, 0);
    int sample = 1;
    bool sampled = false, hit = false;
    for (uint32_t r = 0; r < world; ++r) {
        if (single_rank && r != P.spp_rank) continue;
        Renderer R(S, F, ubo);
        R.Init(x, y, world == 1 ? ubo.TotalFrames : ubo.TotalFrames * world + r);
        bool rhit = R.PrimaryHit(r == 0 || single_rank);  // every rank keeps its own (identical) G-buffer; rank 0 owns the direct term
        if (r == 0 || single_rank) { hit = rhit; direct = R.direct_; }
        if (rhit) {
            int sampleMultiplier = F.moment[o] > 0 ? 4 : 1;  // Core.PathTracing.comp.slang:28
            sample = std::max(1, (int)ubo.NumberOfSamples / R.SampleDownscale * sampleMultiplier);
            int s0 = (int)((uint64_t)r * sample / world), s1 = (int)((uint64_t)(r + 1) * sample / world);
            bool rsampled = R.RenderSamples(s0, s1, FinalColor, FinalReflection);
            if (single_rank && r != 0) sampled = rsampled;
            if (r == 0) {
                sampled = rsampled;
                // world == 1: exactly Shading.slang:1178-1180 (the RNG has consumed all samples by now)
                direct = rsampled ? R.DirectIlluminate(R.hitPrimaryVertex_.Position, R.hitPrimaryVertex_.Normal) : R.direct_;
            }
        }
        F.primary_rays += R.n_primary;
        F.closest_rays += R.n_closest;
        F.any_rays += R.n_any;
    }
    if (single_rank) {
        const size_t n = (size_t)F.w * F.h;
        float* rc = &F.radiance[o * 4];
        float* rr = &F.radiance[(n + o) * 4];
        float* rd = &F.radiance[(2 * n + o) * 4];
        f4 z = F4(0, 0, 0, 0);
        f4 c = (hit && sampled) ? FinalColor : z, rf = (hit && sampled) ? FinalReflection : z, d = P.spp_rank == 0 ? direct : z;
        rc[0] = c.x; rc[1] = c.y; rc[2] = c.z; rc[3] = c.w;
        rr[0] = rf.x; rr[1] = rf.y; rr[2] = rf.z; rr[3] = rf.w;
        rd[0] = d.x; rd[1] = d.y; rd[2] = d.z; rd[3] = d.w;
        return;
    }
    if (!hit) {  // Shading.slang:1195 writes only SINGLE_DIFFUSE
        F.store4(F.single_diffuse, x, y, direct);
        return;
    }
    if (!sampled) {  // emissive primary, :1093-1099
        F.store4(F.single_diffuse, x, y, direct);
        F.store4(F.single_specular, x, y, F4(0, 0, 0, 0));
        return;
    }
    FinalColor = FinalColor / float(sample);  // :1175-1176
    FinalReflection = FinalReflection / float(sample);
    FinalColor = FinalColor + direct;  // :1180
    F.store4(F.single_diffuse, x, y, FinalColor);
    F.store4(F.single_specular, x, y, FinalReflection);
}

// ----------------------------------------------------------------------------------------------- ReProject
float calculateWeight(float centerDist, bool sameObject, bool isCenter, f3 Normal, f3 CenterNormal) {  // Process.ReProject.comp.slang:32-47
    if (isCenter) return 0.4f;
    if (!sameObject) return 0.0f;
    float normalDot = clampf(dot(Normal, CenterNormal), 0.0f, 1.0f);
    const float normalThreshold = 0.98f;
    if (normalDot < normalThreshold) return 0.0f;
    const float normalWeight = (normalDot - normalThreshold) / (1.0f - normalThreshold);
    return normalWeight * 2.0f / (centerDist * 1.5f + 4.0f);
}
f3 FilterHistoryColor(const GkoFrame& F, const std::vector<uint16_t>& prev, f2 subpixel, int px, int py, uint32_t cur, uint32_t p0, uint32_t p1,
                      uint32_t p2, uint32_t p3, f4 currFrameSample) {  // :49-68
    f3 c = xyz(currFrameSample);
    f3 h0 = cur == p0 ? xyz(F.load4(prev, px, py)) : c;
    f3 h1 = cur == p1 ? xyz(F.load4(prev, px + 1, py)) : c;
    f3 h2 = cur == p2 ? xyz(F.load4(prev, px, py + 1)) : c;
    f3 h3 = cur == p3 ? xyz(F.load4(prev, px + 1, py + 1)) : c;
    f3 history = lerp3(lerp3(h0, h1, subpixel.x), lerp3(h2, h3, subpixel.x), subpixel.y);
    return F3(clampf(history.x, 0.0f, 1600.0f), clampf(history.y, 0.0f, 1600.0f), clampf(history.z, 0.0f, 1600.0f));
}

void ReProjectPixel(GkoFrame& F, const GkcuFrameParams& P, int x, int y) {  // Process.ReProject.comp.slang:71-223
    const float k512 = clampf(1.0f / 512.0f, 0.0f, 1.0f);
    if (P.progressive) {
        F.store4(F.accum_diffuse, x, y, lerp4(F.load4(F.prev_diffuse, x, y), F.load4(F.single_diffuse, x, y), k512));
        F.store4(F.accum_specular, x, y, lerp4(F.load4(F.prev_specular, x, y), F.load4(F.single_specular, x, y), k512));
        F.store4(F.accum_albedo, x, y, lerp4(F.load4(F.prev_albedo, x, y), F.load4(F.albedo, x, y), k512));
        return;
    }
    f4 currDiffuseSample = F.load4(F.single_diffuse, x, y), currSpecularSample = F.load4(F.single_specular, x, y), currAlbedoSample = F.load4(F.albedo, x, y);
    bool useHistory = true;
    f2 motion = F.loadmv(x, y);
    int pvx = (int)floorf((float)x + motion.x), pvy = (int)floorf((float)y + motion.y);
    // all(previpos < isize) compares int with uint: negative coordinates convert to huge unsigned values and fail
    const bool inside = ((uint32_t)pvx < F.w) && ((uint32_t)pvy < F.h) && pvx >= -1 && pvy >= -1;
    uint32_t cur = F.loadu(F.objectid0, x, y);
    uint32_t prev0 = F.loadu(F.objectid1, pvx, pvy);
    const bool miss = cur == 65535;
    uint32_t moment = F.moment[(size_t)y * F.w + x];
    if (miss || !inside || moment > 0) useHistory = false;
    if (useHistory) {
        const int SAMPLE_RANGE = P.fast_reproject ? 0 : 2;
        const f3 CenterNormal = xyz(F.load4(F.normal, x, y));
        f4 spatialDiffuseSample = F4(0, 0, 0, 0), spatialSpecularSample = F4(0, 0, 0, 0);
        float totalWeight = 0.0f, motionInRange = 0;
        for (int yy = -SAMPLE_RANGE; yy <= SAMPLE_RANGE; yy++) {
            for (int xx = -SAMPLE_RANGE; xx <= SAMPLE_RANGE; xx++) {
                const int sx = x + xx, sy = y + yy;
                const uint32_t primitiveIdx = F.loadu(F.objectid0, sx, sy);
                const f3 worldNormal = xyz(F.load4(F.normal, sx, sy));
                const bool isCenter = (xx == 0 && yy == 0);
                const bool sameObject = (primitiveIdx == cur);
                const float centerDist = length(f2{(float)xx, (float)yy});
                const float weight = calculateWeight(centerDist, sameObject, isCenter, worldNormal, CenterNormal);
                spatialDiffuseSample = spatialDiffuseSample + F.load4(F.single_diffuse, sx, sy) * weight;
                spatialSpecularSample = spatialSpecularSample + F.load4(F.single_specular, sx, sy) * weight;
                totalWeight += weight;
                motionInRange += length(F.loadmv(x + xx * 2, y + yy * 2));
            }
        }
        spatialDiffuseSample = spatialDiffuseSample / totalWeight;
        spatialSpecularSample = spatialSpecularSample / totalWeight;
        uint32_t prev1 = F.loadu(F.objectid1, pvx + 1, pvy), prev2 = F.loadu(F.objectid1, pvx, pvy + 1), prev3 = F.loadu(F.objectid1, pvx + 1, pvy + 1);
        if (fabsf(motionInRange) < 0.1f) prev0 = prev1 = prev2 = prev3 = cur;
        f2 pos = f2{(float)x + motion.x, (float)y + motion.y};
        f2 subpixel = f2{pos.x - floorf(pos.x), pos.y - floorf(pos.y)};
        f3 historyDiffuse = FilterHistoryColor(F, F.prev_diffuse, subpixel, pvx, pvy, cur, prev0, prev1, prev2, prev3, spatialDiffuseSample);
        f3 historySpecular = FilterHistoryColor(F, F.prev_specular, subpixel, pvx, pvy, cur, prev0, prev1, prev2, prev3, spatialSpecularSample);
        f3 historyAlbedo = FilterHistoryColor(F, F.prev_albedo, subpixel, pvx, pvy, cur, prev0, prev1, prev2, prev3, currAlbedoSample);
        float currKeep = clampf(1.0f / (float)std::max(1, P.temporal_frames), 0.0f, 1.0f);
        F.store3(F.accum_diffuse, x, y, lerp3(historyDiffuse, xyz(currDiffuseSample), currKeep));
        F.store3(F.accum_specular, x, y, lerp3(historySpecular, xyz(currSpecularSample), currKeep));
        F.store3(F.accum_albedo, x, y, lerp3(historyAlbedo, xyz(currAlbedoSample), currKeep));
    } else {
        F.store4(F.accum_diffuse, x, y, currDiffuseSample);
        F.store4(F.accum_specular, x, y, currSpecularSample);
        F.store4(F.accum_albedo, x, y, currAlbedoSample);
    }
}

// ----------------------------------------------------------------------------------------------- DenoiseJBF
void DenoisePixel(GkoFrame& F, const GkUniformBufferObject& Camera, int dx, int dy) {  // Process.DenoiseJBF.comp.slang:75-182
    const int WINDOW_R = 5;
    const int gx = (int)Camera.ViewportRect.x, gy = (int)Camera.ViewportRect.y;
    const int x = dx + gx, y = dy + gy;
    f3 Total = F3(0, 0, 0);
    if (Camera.BFSize > 0) {
        auto FetchDiffuse = [&](int ax, int ay) { return xyz(F.load4(F.accum_diffuse, ax, ay)) + F3(0.001f, 0.001f, 0.001f); };
        auto FetchSpec = [&](int ax, int ay) { return xyz(F.load4(F.accum_specular, ax, ay)) + F3(0.001f, 0.001f, 0.001f); };
        const f3 CenterColor = FetchDiffuse(x, y);
        const f3 CenterSpec = FetchSpec(x, y);
        const float CenterLum = dot(CenterColor, F3(0.212671f, 0.715160f, 0.072169f));
        const float sigma = Camera.BFSigma, sigmaL = Camera.BFSigmaLum * 100.0f;
        float Weight = 0;
        // the tile is indexed [row][col]; tap (i, j) is row + i, col + j (:49-73, 133-141); the centre is not a tap
        for (int i = -WINDOW_R; i <= WINDOW_R; i += 2) {
            for (int j = -WINDOW_R; j <= WINDOW_R; j += 2) {
                const f3 Ci = FetchDiffuse(x + j, y + i);
                const float lumi = dot(Ci, F3(0.212671f, 0.715160f, 0.072169f));
                const float dist = clampf(float(i * i + j * j) / float(WINDOW_R * WINDOW_R), 0.0f, 1.0f);
                const float dlum = (CenterLum - lumi) * (CenterLum - lumi);
                const float Fi = gk_det_expf(-dist * dist / (2.0f * sigma * sigma));
                const float Li = gk_det_expf(-dlum * dlum / (2.0f * sigmaL * sigmaL));
                Total = Total + Ci * Fi * Li;
                Weight += Fi * Li;
            }
        }
        Total = F3(Total.x / Weight, Total.y / Weight, Total.z / Weight);
        if (!Camera.DebugDraw_Lighting) Total = Total * xyz(F.load4(F.accum_albedo, x, y)) + CenterSpec;
    } else {
        if (Camera.DebugDraw_Lighting) Total = xyz(F.load4(F.accum_diffuse, x, y)) * F3(0.5f, 0.5f, 0.5f) + xyz(F.load4(F.accum_specular, x, y));
        else Total = xyz(F.load4(F.accum_diffuse, x, y)) * xyz(F.load4(F.accum_albedo, x, y)) + xyz(F.load4(F.accum_specular, x, y));
    }
    if (Camera.HDR) {
        Total = F3(Total.x / 2000.0f, Total.y / 2000.0f, Total.z / 2000.0f);
        Total = GT_Tonemapping(Total);
        Total = Total * 2000.0f;
        f3 pw = Total * Camera.PaperWhiteNit;
        f3 st = LinearToST2084UE(F3(pw.x / 230.0f, pw.y / 230.0f, pw.z / 230.0f));
        F.store4(F.denoised, x, y, F4(st.x, st.y, st.z, 1.0f));
    } else {
        f3 pw = Total * Camera.PaperWhiteNit;
        f3 t = GT_Tonemapping(F3(pw.x / 40000.0f, pw.y / 40000.0f, pw.z / 40000.0f));
        F.store4(F.denoised, x, y, F4(t.x, t.y, t.z, 1.0f));
    }
}

// Work items are 64-pixel row segments handed out by an atomic counter, so even a 1-row band uses every core.
template <class Fn>
void parallel_pixels(uint32_t y0, uint32_t y1, uint32_t width, int threads, Fn fn) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const uint32_t seg = 64, segs_per_row = (width + seg - 1) / seg;
    const uint64_t total = (uint64_t)(y1 > y0 ? y1 - y0 : 0) * segs_per_row;
    std::atomic<uint64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            uint64_t i = next.fetch_add(1);
            if (i >= total) break;
            int y = (int)(y0 + i / segs_per_row);
            uint32_t x0 = (uint32_t)(i % segs_per_row) * seg;
            for (uint32_t x = x0; x < std::min(x0 + seg, width); ++x) fn((int)x, y);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
}

}  // namespace

// ------------------------------------------------------------------------------------------------- C API
template <class Fn>
static void parallel_range(uint32_t n, int threads, Fn fn) {
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint32_t> next{0};
    auto worker = [&]() {
        for (;;) {
            uint32_t i0 = next.fetch_add(256);
            if (i0 >= n) break;
            for (uint32_t i = i0; i < std::min(i0 + 256u, n); ++i) fn(i);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
}

extern "C" {

const char* gko_intersector_name(void) {
#ifdef GKO_TINYBVH
    return "tinybvh-1.3.8";
#else
    return "native-watertight";
#endif
}

GkoScene* gko_scene_create(const GkcuSceneDesc* d) {
    GkoScene* s = new GkoScene();
    s->vertices.assign(d->vertices, d->vertices + d->vertex_count);
    s->simple.assign(d->simple_vertices, d->simple_vertices + (size_t)d->vertex_count * 4);
    s->indices.assign(d->indices, d->indices + d->index_count);
    s->offsets.assign(d->offsets, d->offsets + d->offset_count);
    s->model_count = d->model_count;
    s->blas.resize(d->model_count);
    for (uint32_t m = 0; m < d->model_count; ++m) s->build_blas(m);
    return s;
}
void gko_scene_destroy(GkoScene* s) { delete s; }
void gko_update_nodes(GkoScene* s, const GkNodeProxy* nodes, uint32_t count) {
    s->nodes.assign(nodes, nodes + count);
    s->build_tlas();
}
void gko_update_materials(GkoScene* s, const GkMaterial* m, uint32_t count) { s->materials.assign(m, m + count); }
void gko_refit_model(GkoScene* s, uint32_t model, const uint16_t* sv, uint32_t vertex_count) {
    const GkModelData& md = s->offsets[model * 10];
    memcpy(&s->simple[(size_t)md.vertexOffset * 4], sv, (size_t)vertex_count * 8);
    s->build_blas(model);
    if (!s->nodes.empty()) s->build_tlas();
}
void gko_set_sky(GkoScene* s, uint32_t id, const float* rgba, uint32_t w, uint32_t h) {
    if (s->textures.size() <= id) s->textures.resize(id + 1);
    Texture& t = s->textures[id];
    t.w = w;
    t.h = h;
    t.texels.resize((size_t)w * h);
    memcpy(t.texels.data(), rgba, (size_t)w * h * 16);
}
void gko_set_probes(GkoScene* s, const GkAmbientCube* cubes, const GkVoxelData* voxels) {
    size_t n = (size_t)GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_Z;
    if (!cubes || !voxels) { s->cubes.clear(); s->voxels.clear(); return; }
    s->cubes.assign(cubes, cubes + n);
    s->voxels.assign(voxels, voxels + n);
}

void gko_set_lights(GkoScene* s, const GkLightObject* lights, uint32_t count) { s->lights.assign(lights, lights + count); }
void gko_set_sky_sh(GkoScene* s, uint32_t id, const GkSphericalHarmonics* sh) {
    if (s->sky_sh.size() <= id) s->sky_sh.resize(id + 1, GkSphericalHarmonics{});
    s->sky_sh[id] = *sh;
}
void gko_probes_voxelize(GkoScene* s, int threads) {
    const uint32_t n = GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_Z;
    s->voxels.assign(n, GkVoxelData{});
    s->cubes.assign(n, GkAmbientCube{});
    parallel_range(n, threads, [&](uint32_t i) { s->VoxelizeCube(i); });
    s->UpdatePageIndex();
}
void gko_probes_bake(GkoScene* s, const GkUniformBufferObject* ubo, uint32_t first, uint32_t count, int threads) {
    if (s->cubes.empty() || count == 0) return;
    const std::vector<GkAmbientCube> snapshot(s->cubes.begin() + first, s->cubes.begin() + first + count);
    parallel_range(count, threads, [&](uint32_t i) { s->BakeCube(*ubo, first + i, snapshot.data(), first, count); });
}
const GkAmbientCube* gko_probes_cubes(GkoScene* s) { return s->cubes.data(); }
const GkVoxelData* gko_probes_voxels(GkoScene* s) { return s->voxels.data(); }
const GkPageIndex* gko_probes_pages(GkoScene* s) { return s->pages.data(); }
uint32_t gko_pack_rgb10a2(const float c[4]) { return GkoScene::packRGB10A2(F4(c[0], c[1], c[2], c[3])); }
void gko_unpack_rgb10a2(uint32_t p, float out[4]) { f4 c = GkoScene::unpackRGB10A2(p); out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w; }
uint32_t gko_get_material_data(GkoScene* s, uint32_t node_plus1, uint32_t prim, const float ro[3], const float rd[3], float out[18]) {
    Vertex v = get_material_dataV2(*s, node_plus1, prim, F3(ro[0], ro[1], ro[2]), F3(rd[0], rd[1], rd[2]));
    const float o[18] = {v.Position.x, v.Position.y, v.Position.z, v.Normal.x, v.Normal.y, v.Normal.z, v.Tangent.x, v.Tangent.y, v.Tangent.z, v.TexCoord.x, v.TexCoord.y,
                         v.LocalNormal.x, v.LocalNormal.y, v.LocalNormal.z, v.LocalTangent.x, v.LocalTangent.y, v.LocalTangent.z, v.LocalTangent.w};
    for (int i = 0; i < 18; ++i) out[i] = o[i];
    return v.MaterialIndex;
}
uint32_t gko_fetch_material_id(GkoScene* s, uint32_t node, uint32_t offset) { return FetchMaterialId(s->nodes[node], offset); }
void gko_mikktspace(const float vnt[3], const float n[3], const float t[4], float out[3]) {
    f3 r = Renderer::mikkTSpace(F3(vnt[0], vnt[1], vnt[2]), F3(n[0], n[1], n[2]), F4(t[0], t[1], t[2], t[3]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float gko_calculate_weight(float centerDist, int sameObject, int isCenter, const float n[3], const float cn[3]) {
    return calculateWeight(centerDist, sameObject != 0, isCenter != 0, F3(n[0], n[1], n[2]), F3(cn[0], cn[1], cn[2]));
}
uint32_t gko_lerp_packed_color_alt(uint32_t c0, const float c1[4], float t) { return GkoScene::LerpPackedColorAlt(c0, F4(c1[0], c1[1], c1[2], c1[3]), t); }
void gko_sample_ambient_cube(int which, const uint32_t cube14[14], const float n[3], float out[4]) {  /* which: 0 DI, 2 Full */
    GkAmbientCube c;
    memcpy(&c, cube14, sizeof(c));
    f4 r = which == 0 ? GkoScene::sampleAmbientCubeHL2_DI(c, F3(n[0], n[1], n[2])) : GkoScene::sampleAmbientCubeHL2_Full(c, F3(n[0], n[1], n[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
void gko_sample_ibl_rough(GkoScene* s, uint32_t id, const float dir[3], float rotate, float out[4]) {
    f4 c = s->SampleIBLRough(id, F3(dir[0], dir[1], dir[2]), rotate); out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w;
}

// ---- services for oracle/ref_pathtrace.cpp (see gk_oracle.h)
void gko_scene_arrays(GkoScene* s, GkoSceneArrays* out) {
    out->nodes = s->nodes.data(); out->offsets = s->offsets.data(); out->vertices = s->vertices.data(); out->indices = s->indices.data();
    out->materials = s->materials.data(); out->material_count = (uint32_t)s->materials.size(); out->node_count = (uint32_t)s->nodes.size();
    out->cubes = s->cubes.empty() ? nullptr : s->cubes.data(); out->voxels = s->voxels.empty() ? nullptr : s->voxels.data();
    out->lights = s->lights.data(); out->light_count = (uint32_t)s->lights.size(); out->sky_sh = s->sky_sh.data(); out->sky_sh_count = (uint32_t)s->sky_sh.size();
}
void gko_trace_one(GkoScene* s, const GkcuRay* r, GkcuHit* hit, uint32_t any_hit) {
    if (any_hit) {
        bool occ = s->occluded(F3(r->ox, r->oy, r->oz), F3(r->dx, r->dy, r->dz), r->tmin, r->tmax);
        *hit = {0, 0, 0, occ ? 0u : 0xFFFFFFFFu, 0};
    } else {
        Hit h = s->closest(F3(r->ox, r->oy, r->oz), F3(r->dx, r->dy, r->dz), r->tmin, r->tmax);
        *hit = {h.t, h.u, h.v, h.inst, h.prim};
    }
}
void gko_interpolate_ambient_cubes(GkoScene* s, int which, const float pos[3], const float n[3], float out[4]) {
    f4 c = s->interpolateAmbientCubesT(F3(pos[0], pos[1], pos[2]), F3(n[0], n[1], n[2]), which == 0, nullptr, 0, 0);
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w;
}
int gko_vox_trace(GkoScene* s, const float o[3], const float d[3], float dist, float* t, float normal[3], uint32_t* material) {
    f3 n;
    bool hit = s->VoxTrace(F3(o[0], o[1], o[2]), F3(d[0], d[1], d[2]), dist, *t, n, *material);
    if (hit) { normal[0] = n.x; normal[1] = n.y; normal[2] = n.z; }
    return hit ? 1 : 0;
}
void gko_instance_w2o(GkoScene* s, uint32_t inst, float out[12]) { std::memcpy(out, s->instances[inst].w2o, 48); }
void gko_sample_texture(GkoScene* s, uint32_t id, float u, float v, float out[4]) { f4 c = s->sample_texture(id, f2{u, v}); out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w; }
void gko_cast_visibility_rows(GkoScene* s, const GkUniformBufferObject* ubo, uint32_t w, uint32_t h, uint32_t y0, uint32_t y1, int threads, uint32_t* packed) {
    GkoFrame F;
    F.w = w; F.h = h;
    const f4 origin = mul(ubo->ModelViewInverse, F4(0, 0, 0, 1));
    parallel_pixels(y0, std::min(y1, h), w, threads, [&](int x, int y) {
        Renderer R(*s, F, *ubo);
        uint32_t node1, prim;
        R.CastVisibility(x, y, xyz(origin), node1, prim);
        packed[(size_t)y * w + x] = (node1 << 17) | (prim & 0x1FFFFu);
    });
}
void gko_cast_visibility(GkoScene* s, const GkUniformBufferObject* ubo, uint32_t w, uint32_t h, uint32_t* packed) {
    gko_cast_visibility_rows(s, ubo, w, h, 0, h, 1, packed);
}


// ------------------------------------------------------------------------------------------------- visibility front end
// "gpu cull": Task.GpuCull.comp.slang:199-227 with IsAABBInFrustum (:127-197) and IsOccludedVolume (:56-115), one NodeProxy at a time.
// "visibility pass": Rast.VisibilityPass.{vert,frag}.slang under the Vulkan rasterisation rules (the fixed-function unit of the reference): near clip,
// perspective divide, viewport transform, 8 sub-pixel bits, back faces culled by the sign of the area (cull BACK, front face counter-clockwise,
// PipelineCommon/CommonComputePipeline.cpp:201-202), coverage at pixel centres with the top-left rule, depth z/w interpolated in screen space, LESS (:221),
// depth ties to the smaller id. Restated independently of csrc/gk_raster.cu from the same specification; both must agree bit for bit.
namespace {
struct DrawCmdO { uint32_t indexCount, instanceCount, firstIndex; int32_t vertexOffset; uint32_t firstInstance; };

GkMat4 MatMul(const GkMat4& a, const GkMat4& b) {
    GkMat4 r;
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row)
            r.m[c][row] = a.m[0][row] * b.m[c][0] + a.m[1][row] * b.m[c][1] + a.m[2][row] * b.m[c][2] + a.m[3][row] * b.m[c][3];
    return r;
}

bool AABBInFrustum(f3 mn, f3 mx, const GkMat4& wvp) {
    f4 clip[8];
    for (int i = 0; i < 8; ++i) clip[i] = mul(wvp, F4(i & 1 ? mx.x : mn.x, i & 2 ? mx.y : mn.y, i & 4 ? mx.z : mn.z, 1.0f));
    auto all_outside = [&](auto inside) {
        for (int i = 0; i < 8; ++i)
            if (inside(clip[i])) return false;
        return true;
    };
    if (all_outside([](f4 c) { return c.x >= -c.w; })) return false;
    if (all_outside([](f4 c) { return c.x <= c.w; })) return false;
    if (all_outside([](f4 c) { return c.y >= -c.w; })) return false;
    if (all_outside([](f4 c) { return c.y <= c.w; })) return false;
    if (all_outside([](f4 c) { return c.z >= -c.w; })) return false;
    if (all_outside([](f4 c) { return c.z <= c.w; })) return false;
    return true;
}

bool OccludedVolume(f3 mn, f3 mx, const GkMat4& wvp, const float* prev_depth, uint32_t w, uint32_t h) {
    const int GRID = 6;
    const float GRID_M1 = GRID - 1;
    for (int x = 0; x < GRID; x++)
        for (int y = 0; y < GRID; y++)
            for (int z = 0; z < GRID; z++) {
                const f3 g = F3(mn.x + (mx.x - mn.x) * (float(x) / GRID_M1), mn.y + (mx.y - mn.y) * (float(y) / GRID_M1), mn.z + (mx.z - mn.z) * (float(z) / GRID_M1));
                const f4 c = mul(wvp, F4(g.x, g.y, g.z, 1.0f));
                const f3 ndc = F3(c.x / c.w, c.y / c.w, c.z / c.w);
                if (fabsf(ndc.x) > 1.0f || fabsf(ndc.y) > 1.0f || ndc.z < 0.0f || ndc.z > 1.0f) return false;
                const float u = ndc.x * 0.5f + 0.5f, v = ndc.y * 0.5f + 0.5f;
                const int px = (int)(u * float(w)), py = (int)(v * float(h));
                if (px < 0 || px >= (int)w || py < 0 || py >= (int)h) return false;
                if (prev_depth[(size_t)py * w + px] > ndc.z - 0.0001f) return false;
            }
    return true;
}

struct Frag { uint64_t key; };

int64_t SnapFixed(float v) {
    const float s = floorf(v * 256.0f + 0.5f);
    const float lim = 1073741824.0f;  // 2^30
    if (!(s > -lim)) return -(int64_t(1) << 30);
    if (!(s < lim)) return int64_t(1) << 30;
    return (int64_t)s;
}
int64_t CeilDiv256(int64_t a) { return a >= 0 ? (a + 255) / 256 : -((-a) / 256); }
int64_t FloorDiv256(int64_t a) { return a >= 0 ? a / 256 : -((-a + 255) / 256); }

// one triangle given in clip space, all z >= 0
void RasterClipped(const f4* c, uint32_t w, uint32_t h, uint32_t id, std::vector<uint64_t>& keys) {
    int64_t X[3], Y[3];
    float Z[3];
    for (int k = 0; k < 3; ++k) {
        const float nx = c[k].x / c[k].w, ny = c[k].y / c[k].w;
        Z[k] = c[k].z / c[k].w;
        X[k] = SnapFixed((nx + 1.0f) * ((float)w * 0.5f));
        Y[k] = SnapFixed((ny + 1.0f) * ((float)h * 0.5f));
    }
    const int64_t area2 = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0]);
    if (area2 >= 0) return;  // back-facing (or degenerate): the Vulkan area -area2 / 2 must be positive for a counter-clockwise front face
    std::swap(X[1], X[2]); std::swap(Y[1], Y[2]); std::swap(Z[1], Z[2]);  // edge functions positive inside
    const int64_t x0 = std::max<int64_t>(0, CeilDiv256(std::min({X[0], X[1], X[2]}) - 128)), x1 = std::min<int64_t>((int64_t)w - 1, FloorDiv256(std::max({X[0], X[1], X[2]}) - 128));
    const int64_t y0 = std::max<int64_t>(0, CeilDiv256(std::min({Y[0], Y[1], Y[2]}) - 128)), y1 = std::min<int64_t>((int64_t)h - 1, FloorDiv256(std::max({Y[0], Y[1], Y[2]}) - 128));
    for (int64_t py = y0; py <= y1; ++py)
        for (int64_t px = x0; px <= x1; ++px) {
            const int64_t Px = px * 256 + 128, Py = py * 256 + 128;
            int64_t E[3];
            bool inside = true;
            for (int i = 0; i < 3 && inside; ++i) {
                const int a = (i + 1) % 3, b = (i + 2) % 3;
                const int64_t dx = X[b] - X[a], dy = Y[b] - Y[a];
                E[i] = dx * (Py - Y[a]) - dy * (Px - X[a]);
                const bool topleft = dy < 0 || (dy == 0 && dx > 0);
                if (E[i] < 0 || (E[i] == 0 && !topleft)) inside = false;
            }
            if (!inside) continue;
            const double area = (double)(E[0] + E[1] + E[2]);
            const double zs = ((double)E[0] * (double)Z[0] + (double)E[1] * (double)Z[1]) + (double)E[2] * (double)Z[2];
            const float z = (float)(zs / area);
            if (!(z >= 0.0f && z < 1.0f)) continue;
            uint32_t zb;
            std::memcpy(&zb, &z, 4);
            const uint64_t key = ((uint64_t)zb << 32) | id;
            uint64_t& dst = keys[(size_t)py * w + px];
            if (key < dst) dst = key;
        }
}
}  // namespace

void gko_set_raster_buffers(GkoScene* s, const uint32_t* raster_indices, uint32_t index_count, const uint32_t* reorders, uint32_t reorder_count) {
    s->raster_indices.assign(raster_indices, raster_indices + index_count);
    s->reorders.assign(reorders, reorders + reorder_count);
}

// cmds: count x VkDrawIndexedIndirectCommand (5 words), stats: GPUDrivenStat (8 words)
void gko_cull(GkoScene* s, const GkUniformBufferObject* ubo, const float* prev_depth, uint32_t w, uint32_t h, uint32_t* cmds, uint32_t* stats) {
    GkGPUDrivenStat st{};
    for (size_t i = 0; i < s->nodes.size(); ++i) {
        const GkNodeProxy& node = s->nodes[i];
        const GkModelData& model = s->offsets[node.modelId];
        const GkMat4 mvp = MatMul(ubo->ViewProjection, node.worldTS);
        const f3 mn = F3(model.localAabbMin.x, model.localAabbMin.y, model.localAabbMin.z), mx = F3(model.localAabbMax.x, model.localAabbMax.y, model.localAabbMax.z);
        bool shouldDraw = node.visible > 0 && AABBInFrustum(mn, mx, mvp);
        if (shouldDraw) {
            st.ProcessedCount += 1;
            st.TriangleCount += model.indexCount / 3;
            if (OccludedVolume(mn, mx, mvp, prev_depth, w, h)) {
                st.CulledCount += 1;
                st.CulledTriangleCount += model.indexCount / 3;
                shouldDraw = false;
            }
        }
        DrawCmdO c;
        c.instanceCount = shouldDraw ? 1 : 0;
        c.firstInstance = shouldDraw ? (uint32_t)i : 0;
        c.indexCount = shouldDraw ? model.indexCount : 0;
        c.firstIndex = shouldDraw ? model.indexOffset : 0;
        c.vertexOffset = shouldDraw ? (int32_t)model.reorderOffset : 0;
        std::memcpy(cmds + i * 5, &c, 20);
    }
    std::memcpy(stats, &st, sizeof(st));
}

void gko_visibility_pass(GkoScene* s, const GkUniformBufferObject* ubo, const uint32_t* cmds, uint32_t count, uint32_t w, uint32_t h, uint32_t* vis) {
    std::vector<uint64_t> keys((size_t)w * h, (uint64_t)0x3F800000u << 32);
    for (uint32_t b = 0; b < count; ++b) {
        DrawCmdO cmd;
        std::memcpy(&cmd, cmds + (size_t)b * 5, 20);
        if (cmd.instanceCount == 0 || cmd.indexCount == 0) continue;
        const GkMat4 mvp = MatMul(ubo->ViewProjection, s->nodes[cmd.firstInstance].worldTS);
        const uint32_t instanceIdx = cmd.firstInstance + 1;
        for (uint32_t t = 0; t < cmd.indexCount / 3; ++t) {
            f4 c[3];
            uint32_t triangleIdx = 0;
            for (int k = 0; k < 3; ++k) {
                const uint32_t vertexIndex = s->raster_indices[cmd.firstIndex + 3 * t + k] + (uint32_t)cmd.vertexOffset;
                if (k == 0) triangleIdx = vertexIndex - (uint32_t)cmd.vertexOffset;
                const uint32_t vi = s->reorders[vertexIndex];
                c[k] = mul(mvp, F4(h2f(s->simple[(size_t)vi * 4]), h2f(s->simple[(size_t)vi * 4 + 1]), h2f(s->simple[(size_t)vi * 4 + 2]), 1.0f));
            }
            const uint32_t id = ((instanceIdx & 0x7FFFu) << 17) | (triangleIdx & 0x1FFFFu);
            const bool in[3] = {c[0].z >= 0.0f, c[1].z >= 0.0f, c[2].z >= 0.0f};
            if (in[0] && in[1] && in[2]) { RasterClipped(c, w, h, id, keys); continue; }
            if (!in[0] && !in[1] && !in[2]) continue;
            f4 poly[4];
            int n = 0;
            for (int k = 0; k < 3; ++k) {  // Sutherland-Hodgman against z >= 0, in vertex order
                const f4 a = c[k], bb = c[(k + 1) % 3];
                if (in[k]) poly[n++] = a;
                if (in[k] != in[(k + 1) % 3]) {
                    const float tt = a.z / (a.z - bb.z);
                    poly[n++] = F4(a.x + (bb.x - a.x) * tt, a.y + (bb.y - a.y) * tt, 0.0f, a.w + (bb.w - a.w) * tt);
                }
            }
            const f4 t0[3] = {poly[0], poly[1], poly[2]};
            RasterClipped(t0, w, h, id, keys);
            if (n == 4) {
                const f4 t1[3] = {poly[0], poly[2], poly[3]};
                RasterClipped(t1, w, h, id, keys);
            }
        }
    }
    for (size_t i = 0; i < keys.size(); ++i) vis[i] = (uint32_t)(keys[i] & 0xFFFFFFFFull);
}

GkoFrame* gko_frame_create(uint32_t w, uint32_t h) {
    GkoFrame* f = new GkoFrame();
    f->w = w;
    f->h = h;
    size_t n = (size_t)w * h;
    for (auto* img : {&f->accum_diffuse, &f->accum_specular, &f->accum_albedo, &f->prev_diffuse, &f->prev_specular, &f->prev_albedo,
                      &f->single_diffuse, &f->single_specular, &f->albedo, &f->normal, &f->denoised})
        img->assign(n * 4, 0);
    f->objectid0.assign(n, 0);
    f->objectid1.assign(n, 0);
    f->primitiveid.assign(n, 0);
    f->motion.assign(n * 2, 0.f);
    f->prev_depth.assign(n, 0.f);
    f->moment.assign(n, 0);
    f->vis.assign(n, 0);
    f->radiance.assign(n * 12, 0.f);
    return f;
}
void gko_frame_destroy(GkoFrame* f) { delete f; }

int gko_reproject(GkoFrame* f, const GkcuFrameParams* params) {
    parallel_pixels(0, f->h, f->w, 0, [&](int x, int y) { ReProjectPixel(*f, *params, x, y); });
    return 0;
}
int gko_denoise(GkoFrame* f, const GkUniformBufferObject* ubo) {
    parallel_pixels(0, f->h, f->w, 0, [&](int x, int y) { DenoisePixel(*f, *ubo, x, y); });
    return 0;
}
void gko_end_frame_copies(GkoFrame* f) {  // PathTracingRenderer.cpp:160-182, VulkanBaseRenderer.cpp:1136-1152
    f->prev_diffuse = f->accum_diffuse;
    f->prev_specular = f->accum_specular;
    f->prev_albedo = f->accum_albedo;
    f->objectid1 = f->objectid0;
}

int gko_render_frame(GkoScene* s, GkoFrame* f, const GkUniformBufferObject* ubo, const GkcuFrameParams* params, int threads, uint32_t y0,
                     uint32_t y1, int run_post) {
    y1 = std::min(y1, f->h);
    f->primary_rays = 0;
    f->closest_rays = 0;
    f->any_rays = 0;
    f->use_vis = params->primary_visibility != 0;
    if (f->use_vis) {  // VulkanBaseRenderer::PreRender: "gpu cull" against last frame's depth, then the "visibility pass" (whole image)
        std::vector<uint32_t> cmds(std::max<size_t>(1, s->nodes.size()) * 5), stats(8);
        gko_cull(s, ubo, f->prev_depth.data(), f->w, f->h, cmds.data(), stats.data());
        gko_visibility_pass(s, ubo, cmds.data(), (uint32_t)s->nodes.size(), f->w, f->h, f->vis.data());
    }
    auto t0 = std::chrono::steady_clock::now();
    parallel_pixels(y0, y1, f->w, threads, [&](int x, int y) { RenderPixel(*s, *f, *ubo, *params, x, y); });
    f->rt_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (run_post) {
        parallel_pixels(0, f->h, f->w, threads, [&](int x, int y) { ReProjectPixel(*f, *params, x, y); });
        parallel_pixels(0, f->h, f->w, threads, [&](int x, int y) { DenoisePixel(*f, *ubo, x, y); });
        gko_end_frame_copies(f);
    }
    return 0;
}

void* gko_image(GkoFrame* f, uint32_t rt, size_t* bytes) {
    auto ret16 = [&](std::vector<uint16_t>& v) { if (bytes) *bytes = v.size() * 2; return (void*)v.data(); };
    auto ret32 = [&](std::vector<uint32_t>& v) { if (bytes) *bytes = v.size() * 4; return (void*)v.data(); };
    switch (rt) {
        case GK_RT_ACCUMLATE_DIFFUSE: return ret16(f->accum_diffuse);
        case GK_RT_ACCUMLATE_SPECULAR: return ret16(f->accum_specular);
        case GK_RT_ACCUMLATE_ALBEDO: return ret16(f->accum_albedo);
        case GK_RT_SINGLE_PREV_DIFFUSE: return ret16(f->prev_diffuse);
        case GK_RT_SINGLE_PREV_SPECULAR: return ret16(f->prev_specular);
        case GK_RT_SINGLE_PREV_ALBEDO: return ret16(f->prev_albedo);
        case GK_RT_SINGLE_DIFFUSE: return ret16(f->single_diffuse);
        case GK_RT_SINGLE_SPECULAR: return ret16(f->single_specular);
        case GK_RT_ALBEDO: return ret16(f->albedo);
        case GK_RT_NORMAL: return ret16(f->normal);
        case GK_RT_DENOISED: return ret16(f->denoised);
        case GK_RT_MOTIONMOMENT: return ret16(f->moment);
        case GK_RT_MINIGBUFFER: return ret32(f->vis);
        case GK_RT_OBJEDCTID_0: return ret32(f->objectid0);
        case GK_RT_OBJEDCTID_1: return ret32(f->objectid1);
        case GK_RT_EXT_PRIMITIVEID: return ret32(f->primitiveid);
        case GK_RT_MOTIONVECTOR: if (bytes) *bytes = f->motion.size() * 4; return f->motion.data();
        case GK_RT_PREV_DEPTHBUFFER: if (bytes) *bytes = f->prev_depth.size() * 4; return f->prev_depth.data();
    }
    if (bytes) *bytes = 0;
    return nullptr;
}
void gko_get_counters(GkoFrame* f, GkcuCounters* out) {
    memset(out, 0, sizeof(*out));
    out->primary_rays = f->primary_rays;
    out->closest_hit_rays = f->closest_rays;
    out->any_hit_rays = f->any_rays;
}
double gko_last_rt_seconds(GkoFrame* f) { return f->rt_seconds; }

void gko_trace_rays(GkoScene* s, const GkcuRay* rays, uint32_t count, GkcuHit* hits, uint32_t any_hit) {
    parallel_pixels(0, 1, count, 0, [&](int xi, int) {
        {
            const uint32_t i = (uint32_t)xi;
            const GkcuRay& r = rays[i];
            if (any_hit) {
                bool occ = s->occluded(F3(r.ox, r.oy, r.oz), F3(r.dx, r.dy, r.dz), r.tmin, r.tmax);
                hits[i] = {0, 0, 0, occ ? 0u : 0xFFFFFFFFu, 0};
            } else {
                Hit h = s->closest(F3(r.ox, r.oy, r.oz), F3(r.dx, r.dy, r.dz), r.tmin, r.tmax);
                hits[i] = {h.t, h.u, h.v, h.inst, h.prim};
            }
        }
    });
}

void gko_pcg4d(uint32_t v[4]) { u4 s{v[0], v[1], v[2], v[3]}; pcg4d(s); v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w; }
float gko_uint_to_float(uint32_t x) { return uint_to_float(x); }
uint16_t gko_float_to_half(float f) { return f2h(f); }
float gko_half_to_float(uint16_t h) { return h2f(h); }
void gko_gt_tonemap(const float in[3], float out[3]) { f3 r = GT_Tonemapping(F3(in[0], in[1], in[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void gko_ggx_sampling(uint32_t seed[4], float roughness, const float normal[3], float out[3]) {
    u4 s{seed[0], seed[1], seed[2], seed[3]};
    f3 r = ggxSampling(s, roughness, F3(normal[0], normal[1], normal[2]));
    seed[0] = s.x; seed[1] = s.y; seed[2] = s.z; seed[3] = s.w;
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void gko_random_in_hemisphere1(uint32_t seed[4], float out[3]) {
    u4 s{seed[0], seed[1], seed[2], seed[3]};
    f3 r = RandomInHemiSphere1(s);
    seed[0] = s.x; seed[1] = s.y; seed[2] = s.z; seed[3] = s.w;
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float gko_schlick(float cosine, float ior) { return Schlick(cosine, ior); }
void gko_st2084(const float in[3], float out[3]) { f3 r = LinearToST2084UE(F3(in[0], in[1], in[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void gko_align_with_normal(const float ray[3], const float n[3], float out[3]) {
    f3 r = AlignWithNormal(F3(ray[0], ray[1], ray[2]), F3(n[0], n[1], n[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float gko_random_float(uint32_t seed[4]) {
    u4 s{seed[0], seed[1], seed[2], seed[3]};
    float r = RandomFloat(s);
    seed[0] = s.x; seed[1] = s.y; seed[2] = s.z; seed[3] = s.w;
    return r;
}
void gko_random_float2(uint32_t seed[4], float out[2]) {
    u4 s{seed[0], seed[1], seed[2], seed[3]};
    f2 r = RandomFloat2(s);
    seed[0] = s.x; seed[1] = s.y; seed[2] = s.z; seed[3] = s.w;
    out[0] = r.x; out[1] = r.y;
}
void gko_random_in_unit_disk(uint32_t seed[4], float out[2]) {
    u4 s{seed[0], seed[1], seed[2], seed[3]};
    f2 r = RandomInUnitDisk(s);
    seed[0] = s.x; seed[1] = s.y; seed[2] = s.z; seed[3] = s.w;
    out[0] = r.x; out[1] = r.y;
}
void gko_random_in_cone(uint32_t seed[4], float cos_theta, float out[3]) {
    u4 s{seed[0], seed[1], seed[2], seed[3]};
    f3 r = RandomInCone(s, cos_theta);
    seed[0] = s.x; seed[1] = s.y; seed[2] = s.z; seed[3] = s.w;
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void gko_det_sincos(float x, float* s, float* c) { gk_det_sincosf(x, s, c); }
float gko_det_math(int fn, float a, float b) {
    switch (fn) {
        case 0: return gk_det_atan2f(a, b);
        case 1: return gk_det_acosf(a);
        case 2: return gk_det_expf(a);
        case 3: return gk_det_logf(a);
        default: return gk_det_powf(a, b);
    }
}
float* gko_radiance(GkoFrame* f, size_t* floats) { if (floats) *floats = f->radiance.size(); return f->radiance.data(); }
void gko_resolve(GkoFrame* f, const GkUniformBufferObject* ubo) {
    const size_t n = (size_t)f->w * f->h;
    for (size_t o = 0; o < n; ++o) {
        int x = (int)(o % f->w), y = (int)(o / f->w);
        int sample = std::max(1, (int)ubo->NumberOfSamples * (f->moment[o] > 0 ? 4 : 1));
        const float* rc = &f->radiance[o * 4];
        const float* rr = &f->radiance[(n + o) * 4];
        const float* rd = &f->radiance[(2 * n + o) * 4];
        f4 fc = F4(rc[0], rc[1], rc[2], rc[3]) / float(sample), fr = F4(rr[0], rr[1], rr[2], rr[3]) / float(sample);
        fc = fc + F4(rd[0], rd[1], rd[2], rd[3]);
        f->store4(f->single_diffuse, x, y, fc);
        f->store4(f->single_specular, x, y, fr);
    }
}

}  // extern "C"
