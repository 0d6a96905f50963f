// Compiles the reference's two stencil compute shaders as C++ and runs them over caller-provided storage images, so that the oracle's
// restatements of SURVEY 8a rows a23 / a24 can be pinned against the reference's own text:
//   assets/shaders/Process.ReProject.comp.slang   (whole file: calculateWeight, FilterHistoryColor, main)
//   assets/shaders/Process.DenoiseJBF.comp.slang  (whole file: the groupshared tile, JBF, main)
// oracle/Makefile copies each file through sed into oracle/_ref/*_gen.inc where it only rewrites what C++ cannot parse: the
// [shader]/[numthreads] attributes, `import`, the `: SV_*` semantics, `let` -> `auto`, `groupshared` -> `static`,
// `inout T x` -> `T& x`, the push-constant declaration, `main` -> a unique name, and the one reversed swizzle `.yx`. The bodies are the
// reference's. This file supplies the vocabulary: HLSL vectors (ref_hlsl_shim.h), storage images in the reference's formats
// (VulkanBaseRenderer.cpp:439-457: fp16x4 with round-to-nearest-even stores, R32_UINT, RG32F, R16_UINT; out-of-range loads return 0
// and out-of-range stores are dropped, i.e. robust image access), the camera block (the reference's own BasicTypes.slang compiled as
// C++), and a thread group: DenoiseJBF's 8 x 8 group runs on 64 threads that meet at GroupMemoryBarrierWithGroupSync().
// The tonemappers come from the reference's Const_Func.slang extract (const_func_gen.inc). Built only when /root/reference is mounted.
// TEST INFRASTRUCTURE ONLY.
#include <barrier>
#include <thread>
#include <vector>

#include "ref_hlsl_shim.h"

namespace ref {
using std::exp;
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
    explicit int2(uint2 v) : x((int)v.x), y((int)v.y) {}
    explicit int2(float2 v) : x((int)v.x), y((int)v.y) {}
};
inline int2 operator+(int2 a, int2 b) { return int2(a.x + b.x, a.y + b.y); }
inline int2 operator+(int2xy a, int2 b) { return int2(a.x + b.x, a.y + b.y); }
inline int2 operator+(int2 a, int b) { return int2(a.x + b, a.y + b); }
inline int2 operator*(int2 a, int b) { return int2(a.x * b, a.y * b); }
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
inline int max(int a, int b) { return a > b ? a : b; }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float3& operator/=(float3& a, float s) { a = a / s; return a; }
inline float4& operator/=(float4& a, float s) { a = a / s; return a; }

// ---- storage images
inline uint16_t f2h(float f) { _Float16 h = (_Float16)f; uint16_t b; std::memcpy(&b, &h, 2); return b; }  // round to nearest even
inline float h2f(uint16_t b) { _Float16 h; std::memcpy(&h, &b, 2); return (float)h; }
enum { FMT_F16X4 = 0, FMT_RG32F = 1 };
struct ImgF4 {
    uint8_t* base; int w, h, fmt;
    bool inside(int x, int y) const { return x >= 0 && y >= 0 && x < w && y < h; }
    float4 load(int x, int y) const {
        if (!base || !inside(x, y)) return float4(0, 0, 0, 0);
        if (fmt == FMT_RG32F) { const float* p = (const float*)base + ((size_t)y * w + x) * 2; return float4(p[0], p[1], 0.0f, 1.0f); }
        const uint16_t* p = (const uint16_t*)base + ((size_t)y * w + x) * 4;
        return float4(h2f(p[0]), h2f(p[1]), h2f(p[2]), h2f(p[3]));
    }
    void store(int x, int y, float4 v) {
        if (!base || !inside(x, y) || fmt != FMT_F16X4) return;
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
};
struct ImgU {
    const uint8_t* base; int w, h, bytes;  // bytes per texel: 4 (R32_UINT) or 2 (R16_UINT)
    struct Texel { uint r; operator uint() const { return r; } };
    Texel operator[](int2 p) const {
        if (!base || p.x < 0 || p.y < 0 || p.x >= w || p.y >= h) return Texel{0u};
        const size_t i = (size_t)p.y * w + p.x;
        return Texel{bytes == 4 ? ((const uint32_t*)base)[i] : (uint)((const uint16_t*)base)[i]};
    }
};
template <class T> struct RWTexture2DSel;
template <> struct RWTexture2DSel<float4> { typedef ImgF4 type; };
template <> struct RWTexture2DSel<uint> { typedef ImgU type; };
template <class T> using RWTexture2D = typename RWTexture2DSel<T>::type;

// ---- the camera block: the reference's own BasicTypes.slang, as src/Assets/UniformBuffer.hpp:30-48 includes it into C++
#define float4x4 float4x4
#define half4 half4
#define public
#define bool uint32_t
#include "BasicTypes.slang"
#undef public
#undef bool

struct GpuSceneShim { UniformBufferObject* Camera; };
struct BindlessShim {
    // storage-image indices, assets/shaders/common/BindlessTexture.slang:8-30
    static const int RT_ACCUMLATE_DIFFUSE = 0, RT_SINGLE_DIFFUSE = 1, RT_OBJEDCTID_0 = 3, RT_OBJEDCTID_1 = 4, RT_MOTIONVECTOR = 5, RT_ALBEDO = 6, RT_NORMAL = 7,
                     RT_DENOISED = 9, RT_ACCUMLATE_SPECULAR = 11, RT_SINGLE_SPECULAR = 12, RT_ACCUMLATE_ALBEDO = 13, RT_MOTIONMOMENT = 14,
                     RT_SINGLE_PREV_DIFFUSE = 20, RT_SINGLE_PREV_SPECULAR = 21, RT_SINGLE_PREV_ALBEDO = 22;
    ImgF4 f4[32];
    ImgU u[32];
    GpuSceneShim scene;
    template <class T> RWTexture2D<T> GetStorageTexture(int index);
    template <class T> uint2 GetStorageTextureDimensions(int index) { return uint2{(uint)f4[index].w, (uint)f4[index].h}; }
    GpuSceneShim& GetGpuscene() { return scene; }
};
template <> inline ImgF4 BindlessShim::GetStorageTexture<float4>(int index) { return f4[index]; }
template <> inline ImgU BindlessShim::GetStorageTexture<uint>(int index) { return u[index]; }
static BindlessShim Bindless;

// ---- a thread group for GroupMemoryBarrierWithGroupSync()
static std::barrier<>* g_group_barrier = nullptr;
inline void GroupMemoryBarrierWithGroupSync() { g_group_barrier->arrive_and_wait(); }

#include "PreProcessor.slang"
#include "const_func_gen.inc"  // GT_Tonemapping, LinearToST2084UE (and the rest of the Const_Func.slang extract)
namespace reproject {
#include "reproject_gen.inc"
}
namespace denoise {
#include "denoise_gen.inc"
}
}  // namespace ref

using namespace ref;

// images[rt] = pointer to the storage image in its reference format (or NULL), indexed by the reference's RT_* numbers
static void bind(void* const* images, int w, int h) {
    const int f16[] = {0, 1, 6, 7, 9, 11, 12, 13, 20, 21, 22};
    for (int i : f16) Bindless.f4[i] = ImgF4{(uint8_t*)images[i], w, h, FMT_F16X4};
    Bindless.f4[5] = ImgF4{(uint8_t*)images[5], w, h, FMT_RG32F};
    Bindless.u[3] = ImgU{(const uint8_t*)images[3], w, h, 4};
    Bindless.u[4] = ImgU{(const uint8_t*)images[4], w, h, 4};
    Bindless.u[14] = ImgU{(const uint8_t*)images[14], w, h, 2};
}

extern "C" {
// PathTracingRenderer::Render's reproject dispatch (PathTracingRenderer.cpp:126-136): ceil(w/8) x ceil(h/8) groups of 8 x 8
void ref_reproject(void* const* images, int w, int h, int progressive, int temporal_frames, int fast_reproject) {
    bind(images, w, h);
    reproject::pushConsts.ProgressiveRender = progressive;
    reproject::pushConsts.TemporalFrames = temporal_frames;
    reproject::pushConsts.PrevDiffuseBindlessIdx = BindlessShim::RT_SINGLE_PREV_DIFFUSE;
    reproject::pushConsts.PrevSpecularBindlessIdx = BindlessShim::RT_SINGLE_PREV_SPECULAR;
    reproject::pushConsts.PrevAlbedoBindlessIdx = BindlessShim::RT_SINGLE_PREV_ALBEDO;
    reproject::pushConsts.FastReproject = fast_reproject;
    for (int y = 0; y < (h + 7) / 8 * 8; ++y)
        for (int x = 0; x < (w + 7) / 8 * 8; ++x) reproject::reproject_main(uint3((uint)x, (uint)y, 0u));
}

// the compose dispatch (PathTracingRenderer.cpp:138-157); ubo = the 784-byte UniformBufferObject
void ref_denoise(void* const* images, int w, int h, const void* ubo) {
    bind(images, w, h);
    UniformBufferObject cam;
    static_assert(sizeof(UniformBufferObject) == 784, "UniformBufferObject");
    std::memcpy(&cam, ubo, sizeof(cam));
    Bindless.scene.Camera = &cam;
    for (int gy = 0; gy < (h + 7) / 8; ++gy)
        for (int gx = 0; gx < (w + 7) / 8; ++gx) {
            std::barrier<> bar(64);
            g_group_barrier = &bar;
            std::vector<std::thread> pool;
            for (int t = 0; t < 64; ++t)
                pool.emplace_back([=]() {
                    const uint tx = t & 7, ty = t >> 3;
                    denoise::denoise_main(uint3(gx * 8 + tx, gy * 8 + ty, 0u), uint3(tx, ty, 0u), uint3((uint)gx, (uint)gy, 0u), (uint)t);
                });
            for (auto& th : pool) th.join();
        }
    Bindless.scene.Camera = nullptr;
}
}
