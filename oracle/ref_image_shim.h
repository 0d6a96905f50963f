// Storage images, the camera block and the `Bindless` object for compiling whole shaders of the reference as C++ (oracle/ref_stencils.cpp,
// oracle/ref_pathtrace.cpp): images in the reference's formats (VulkanBaseRenderer.cpp:439-457: fp16x4 with round-to-nearest-even
// stores, R32_UINT, RG32F, R32F, R16_UINT; out-of-range loads return 0, out-of-range stores are dropped = robust image access), the
// reference's own BasicTypes.slang compiled as C++ for every struct, and the integer-vector vocabulary the shaders use.
// TEST INFRASTRUCTURE ONLY, used only when /root/reference is mounted.
#pragma once
#include "ref_hlsl_shim.h"

namespace ref {
using std::floor;

// ---- more vector vocabulary
struct int2;
struct int2xy { int x, y; };
struct int2 {
    union {
        struct { int x, y; };
        int2xy xy;
    };
    int2() : x(0), y(0) {}
    int2(int a, int b) : x(a), y(b) {}
    int2(int2xy v) : x(v.x), y(v.y) {}
    int2(uint2 v) : x((int)v.x), y((int)v.y) {}
    explicit int2(float2 v) : x((int)v.x), y((int)v.y) {}
};
inline int2 operator+(int2 a, int2 b) { return int2(a.x + b.x, a.y + b.y); }
inline int2 operator+(int2xy a, int2 b) { return int2(a.x + b.x, a.y + b.y); }
inline int2 operator+(int2 a, int b) { return int2(a.x + b, a.y + b); }
inline int2 operator*(int2 a, int b) { return int2(a.x * b, a.y * b); }
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int2 operator-(int2 a, int2 b) { return int2(a.x - b.x, a.y - b.y); }
inline int2 clamp(int2 v, int2 lo, int2 hi) { return int2(min(max(v.x, lo.x), hi.x), min(max(v.y, lo.y), hi.y)); }
struct bool2 { bool x, y; };
inline bool all(bool2 b) { return b.x && b.y; }
// int < uint follows the usual arithmetic conversions (HLSL / Slang as C): the int operand becomes unsigned, so a negative coordinate fails
inline bool2 operator<(int2 a, uint2 b) { return bool2{(uint)a.x < b.x, (uint)a.y < b.y}; }
inline bool2 operator>=(int2 a, int2 b) { return bool2{a.x >= b.x, a.y >= b.y}; }
struct uint3 {
    union {
        struct { uint x, y, z; };
        uint2 xy;
    };
    uint3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
};
inline uint2 yx(uint3 v) { return uint2{v.y, v.x}; }
inline float2 to_float2(int2 v) { return float2((float)v.x, (float)v.y); }
inline float2 floor(float2 v) { return float2(std::floor(v.x), std::floor(v.y)); }
inline float2 frac(float2 v) { return float2(v.x - std::floor(v.x), v.y - std::floor(v.y)); }
inline float length(float v) { return std::fabs(v); }
inline float3 lerp(float3 a, float3 b, float t) { return a + (b - a) * t; }
inline float3 clamp(float3 v, float lo, float hi) { return float3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float3& operator/=(float3& a, float s) { a = a / s; return a; }
inline float4& operator/=(float4& a, float s) { a = a / s; return a; }

// ---- storage images
inline uint16_t f2h(float f) { _Float16 h = (_Float16)f; uint16_t b; std::memcpy(&b, &h, 2); return b; }  // round to nearest even
inline float h2f(uint16_t b) { _Float16 h; std::memcpy(&h, &b, 2); return (float)h; }
enum { FMT_F16X4 = 0, FMT_RG32F = 1 };
struct ImgF4 {
    uint8_t* base; int w, h, fmt;
    float* shadow = nullptr;  // optional fp32 copy of what is stored (diagnostics: differences below fp16 resolution)
    bool inside(int x, int y) const { return x >= 0 && y >= 0 && x < w && y < h; }
    float4 load(int x, int y) const {
        if (!base || !inside(x, y)) return float4(0, 0, 0, 0);
        if (fmt == FMT_RG32F) { const float* p = (const float*)base + ((size_t)y * w + x) * 2; return float4(p[0], p[1], 0.0f, 1.0f); }
        const uint16_t* p = (const uint16_t*)base + ((size_t)y * w + x) * 4;
        return float4(h2f(p[0]), h2f(p[1]), h2f(p[2]), h2f(p[3]));
    }
    void store(int x, int y, float4 v) {
        if (!base || !inside(x, y)) return;
        if (shadow) { float* q = shadow + ((size_t)y * w + x) * 4; q[0] = v.x; q[1] = v.y; q[2] = v.z; q[3] = v.w; }
        if (fmt == FMT_RG32F) { float* q = (float*)base + ((size_t)y * w + x) * 2; q[0] = v.x; q[1] = v.y; return; }
        uint16_t* p = (uint16_t*)base + ((size_t)y * w + x) * 4;
        p[0] = f2h(v.x); p[1] = f2h(v.y); p[2] = f2h(v.z); p[3] = f2h(v.w);
    }
    struct RGB : float3 {  // value loaded when the texel is addressed; assignment is a read-modify-write that keeps alpha
        ImgF4* img; int px, py;
        RGB(ImgF4* i, int x, int y, float3 v) : float3(v), img(i), px(x), py(y) {}
        RGB& operator=(float3 v) { float4 o = img->load(px, py); img->store(px, py, float4(v, o.w)); return *this; }
    };
    struct RG : float2 { RG(float2 v) : float2(v) {} };
    struct Texel {
        ImgF4* img; int px, py; float4 v;
        RGB rgb; RG rg;
        Texel(ImgF4* i, int x, int y) : img(i), px(x), py(y), v(i->load(x, y)), rgb(i, x, y, float3(v.x, v.y, v.z)), rg(float2(v.x, v.y)) {}
        operator float4() const { return v; }
        Texel& operator=(float4 n) { img->store(px, py, n); return *this; }
    };
    Texel operator[](int2 p) { return Texel(this, p.x, p.y); }
    void Store(int2 p, float4 v) { store(p.x, p.y, v); }
};
struct ImgU {
    uint8_t* base; int w, h, bytes;  // bytes per texel: 4 (R32_UINT) or 2 (R16_UINT)
    uint load(int x, int y) const {
        if (!base || x < 0 || y < 0 || x >= w || y >= h) return 0u;
        const size_t i = (size_t)y * w + x;
        return bytes == 4 ? ((const uint32_t*)base)[i] : (uint)((const uint16_t*)base)[i];
    }
    void store(int x, int y, uint v) const {
        if (!base || x < 0 || y < 0 || x >= w || y >= h) return;
        const size_t i = (size_t)y * w + x;
        if (bytes == 4) ((uint32_t*)base)[i] = v; else ((uint16_t*)base)[i] = (uint16_t)v;
    }
    struct Texel {
        const ImgU* img; int px, py; uint r;
        operator uint() const { return r; }
        Texel& operator=(uint v) { img->store(px, py, v); r = v; return *this; }
    };
    Texel operator[](int2 p) const { return Texel{this, p.x, p.y, load(p.x, p.y)}; }
    void Store(int2 p, uint v) const { store(p.x, p.y, v); }
};
struct ImgF1 {  // R32_SFLOAT
    float* base; int w, h;
    struct Texel {
        const ImgF1* img; int px, py;
        operator float() const { return img->inside(px, py) ? img->base[(size_t)py * img->w + px] : 0.0f; }
        Texel& operator=(float v) { if (img->inside(px, py)) img->base[(size_t)py * img->w + px] = v; return *this; }
    };
    bool inside(int x, int y) const { return base && x >= 0 && y >= 0 && x < w && y < h; }
    Texel operator[](int2 p) const { return Texel{this, p.x, p.y}; }
    void Store(int2 p, float v) const { (*this)[p] = v; }
};
template <class T> struct RWTexture2DSel;
template <> struct RWTexture2DSel<float4> { typedef ImgF4 type; };
template <> struct RWTexture2DSel<uint> { typedef ImgU type; };
template <> struct RWTexture2DSel<float> { typedef ImgF1 type; };
template <class T> using RWTexture2D = typename RWTexture2DSel<T>::type;

// ---- the camera block: the reference's own BasicTypes.slang, as src/Assets/UniformBuffer.hpp:30-48 includes it into C++
#define float4x4 float4x4
#define half4 half4
#define public
#define bool uint32_t
#include "BasicTypes.slang"
#undef public
#undef bool

#ifdef REF_WITH_SCENE
// Material, Vertex, GPUVertex: BasicTypes.slang:218-254 (its shader-only branch), extracted by oracle/Makefile
#define ALIGN_16 /* empty, as in the shader branch of BasicTypes.slang:5 */
#include "basic_types_shader_gen.inc"
#undef ALIGN_16
struct RaytracingAccelerationStructure { void* scene; };
// the reference's GPUScene push-constant block as the shaders see it (BasicTypes.slang:334-420: device pointers behind properties)
struct GpuSceneShim {
    UniformBufferObject* Camera;
    NodeProxy* Nodes;
    GPUVertex* Vertices;
    uint* Indices;
    Material* Materials;
    Material GetMaterial(int idx) { return Materials[min(MAX_MATERIALS - 1, max(0, idx))]; }  // BasicTypes.slang:376-379
    ModelData* Offsets;
    AmbientCube* Cubes;
    VoxelData* Voxels;
    LightObject* Lights;
    SphericalHarmonics* HDRSHs;
    uint custom_data_0;
    void* TLAS;
    RaytracingAccelerationStructure GetTLAS() { return RaytracingAccelerationStructure{TLAS}; }
};
struct Sampler2D {  // combined image sampler: bilinear, repeat, mip 0 only (the hardware sampler is the oracle's, see gk_oracle.h)
    int id;
    float4 SampleLevel(float2 uv, float lod) const;
};
#else
struct GpuSceneShim { UniformBufferObject* Camera; };
#endif
struct BindlessShim {
    // storage-image indices, assets/shaders/common/BindlessTexture.slang:8-30
    static const int RT_ACCUMLATE_DIFFUSE = 0, RT_SINGLE_DIFFUSE = 1, RT_MINIGBUFFER = 2, RT_PREV_DEPTHBUFFER = 10, RT_OBJEDCTID_0 = 3, RT_OBJEDCTID_1 = 4, RT_MOTIONVECTOR = 5, RT_ALBEDO = 6, RT_NORMAL = 7,
                     RT_DENOISED = 9, RT_ACCUMLATE_SPECULAR = 11, RT_SINGLE_SPECULAR = 12, RT_ACCUMLATE_ALBEDO = 13, RT_MOTIONMOMENT = 14,
                     RT_SINGLE_PREV_DIFFUSE = 20, RT_SINGLE_PREV_SPECULAR = 21, RT_SINGLE_PREV_ALBEDO = 22;
    ImgF4 f4[32];
    ImgU u[32];
    ImgF1 f1[32];
    GpuSceneShim scene;
    template <class T> RWTexture2D<T> GetStorageTexture(int index);
    template <class T> uint2 GetStorageTextureDimensions(int index) { return uint2{(uint)f4[index].w, (uint)f4[index].h}; }
    GpuSceneShim& GetGpuscene() { return scene; }
#ifdef REF_WITH_SCENE
    Sampler2D GetSampleTexture(int index) { return Sampler2D{index}; }
#endif
};
template <> inline ImgF4 BindlessShim::GetStorageTexture<float4>(int index) { return f4[index]; }
template <> inline ImgU BindlessShim::GetStorageTexture<uint>(int index) { return u[index]; }
template <> inline ImgF1 BindlessShim::GetStorageTexture<float>(int index) { return f1[index]; }
static BindlessShim Bindless;

}  // namespace ref
