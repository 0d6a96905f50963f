// Compiles the reference's path-tracing compute shader as C++ and runs it pixel by pixel, so that the oracle's restatement of the whole
// RT pass can be pinned against the reference's own text:
//   assets/shaders/Core.PathTracing.comp.slang   whole file: main
//   assets/shaders/Util.BufferClear.comp.slang   whole file: main
//   assets/shaders/Bake.HwAmbientCube.comp.slang whole file: main (the probe bake), with FGpuProbeGenerator
//                                                (AmbientCube.slang:470-692), the sample grids (17-57), EvaluateSH + SampleIBLRough (Shading.slang:146-192)
//   assets/shaders/common/Shading.slang          the trace distances (15-18), CalculateMotionVectorV2 (50-72), FetchGBufferV2 (82-112),
//                                                SampleIBLV2 (194-199), FVisibilityBufferRayCasterV2 (336-495), FHardwareRayTracerV2
//                                                (735-827), FHardwareDirectIlluminatorV2 (886-908), FPathTracingRendererV2 (995-1261)
//   assets/shaders/common/AmbientCube.slang      grid constants (6-15), colour format and samplers (58-223, ambient_cube_gen.inc),
//                                                the sampler structs and FetchCubeV2 / FetchVoxelV2 (225-266), interpolateAmbientCubesV2 (290-382)
//   assets/shaders/common/Const_Func.slang, GeneralFunc.slang, BasicTypes.slang, PreProcessor.slang, ShaderClock.slang as in the other ref_* files
// oracle/Makefile copies the ranges through sed into oracle/_ref/*_gen.inc and rewrites only what C++ cannot parse: `public`, `[mutating]`,
// `let` -> `auto`, `in` / `out` / `inout` qualifiers -> references, interface-typed parameters -> references, `Common.` qualification,
// `T.Sample` -> `T::Sample` and the generic header of interpolateAmbientCubesV2 -> a template, a `;` after closing struct braces, and
// `float3 frac = frac(x)` -> `= ref::frac(x)` (C++ scoping of the initialiser). The bodies are the reference's.
//
// What is hardware, not shader text, comes from the oracle through its service exports (gk_oracle.h): RayQuery = the oracle's
// intersector (closest hit by (t, instance, primitive), any hit), the texture sampler, and the visibility buffer that the reference's
// raster pass writes (the caller fills RT_MINIGBUFFER by casting rays through the pixel centres, gko_cast_visibility). HLSL intrinsics (reflect, refract, distance, mul,
// ...) are defined here the way the oracle defines them, and sin / cos / atan2 / acos / exp / pow are the pinned kernels of include/gk_detmath.h. So this pins the integrator: the control flow, the order of RNG
// draws, every shading formula and every image store of the RT pass. Built only when /root/reference is mounted. TEST INFRASTRUCTURE ONLY.
#include <atomic>
#include <thread>
#include <vector>

#define REF_WITH_SCENE
#define REF_DET_MATH
#include "gk_oracle.h"
#include "ref_image_shim.h"

namespace ref {
static GkoScene* g_scene = nullptr;
static std::atomic<uint64_t> g_closest{0}, g_any{0};  // totals, folded in once per worker thread
static thread_local uint64_t t_closest = 0, t_any = 0;  // per-thread ray counters: no shared cache line on the ray path

inline float4 Sampler2D::SampleLevel(float2 uv, float lod) const {
    float4 r;
    gko_sample_texture(g_scene, (uint32_t)id, uv.x, uv.y, &r.x);
    return r;
}

// ---- more vocabulary
struct int3 {
    int x, y, z;
    int3(int a, int b, int c) : x(a), y(b), z(c) {}
    explicit int3(float3 v) : x((int)v.x), y((int)v.y), z((int)v.z) {}
    operator float3() const { return float3((float)x, (float)y, (float)z); }
};
inline int3 operator+(int3 a, int3 b) { return int3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 floor(float3 v) { return float3(std::floor(v.x), std::floor(v.y), std::floor(v.z)); }
inline float3 frac(float3 v) { return float3(v.x - std::floor(v.x), v.y - std::floor(v.y), v.z - std::floor(v.z)); }
inline float length(float3 a) { return sqrt(dot(a, a)); }
inline float distance(float3 a, float3 b) { return length(a - b); }
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
inline float3 reflect(float3 i, float3 n) { return i - n * (2.0f * dot(n, i)); }
inline float3 refract(float3 i, float3 n, float eta) {
    float ni = dot(n, i);
    float k = 1.0f - eta * eta * (1.0f - ni * ni);
    if (k < 0.0f) return float3(0, 0, 0);
    return i * eta - n * (eta * ni + sqrt(k));
}
inline float4 min(float4 a, float4 b) { return float4(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z), min(a.w, b.w)); }
inline float4 operator/(uint4 a, float s) { return float4((float)a.x / s, (float)a.y / s, (float)a.z / s, (float)a.w / s); }
inline float4& operator*=(float4& a, float4 b) { a = a * b; return a; }
inline float2 operator/(float2 a, float2 b) { return float2(a.x / b.x, a.y / b.y); }
inline float2 operator/(float2 a, float s) { return float2(a.x / s, a.y / s); }
inline float2 operator*(float2 a, uint2 b) { return float2(a.x * (float)b.x, a.y * (float)b.y); }
inline float4x4 mul(const float4x4& a, const float4x4& b) {
    float4x4 r;
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row) r.m[c][row] = a.m[0][row] * b.m[c][0] + a.m[1][row] * b.m[c][1] + a.m[2][row] * b.m[c][2] + a.m[3][row] * b.m[c][3];
    return r;
}

// ---- inline ray queries on the oracle's intersector
struct RayDesc { float3 Origin; float TMin; float3 Direction; float TMax; };
enum { RAY_FLAG_NONE = 0, RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH = 0x4, RAY_FLAG_SKIP_PROCEDURAL_PRIMITIVES = 0x200 };
enum { COMMITTED_NOTHING = 0, COMMITTED_TRIANGLE_HIT = 1 };
struct float4x3 { float3 row[4]; };  // CommittedRayWorldToObject(): the transpose of the 3 x 4 world-to-object matrix
inline float4 mul(const float4x3& M, float3 v) { return float4(dot(M.row[0], v), dot(M.row[1], v), dot(M.row[2], v), dot(M.row[3], v)); }
template <int FLAGS> struct RayQuery {
    GkcuRay r;
    GkcuHit h;
    void* scene;
    void TraceRayInline(RaytracingAccelerationStructure as, int flags, int mask, RayDesc ray) {
        scene = as.scene;
        r = GkcuRay{ray.Origin.x, ray.Origin.y, ray.Origin.z, ray.TMin, ray.Direction.x, ray.Direction.y, ray.Direction.z, ray.TMax};
    }
    bool Proceed() {
        const bool any = (FLAGS & RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH) != 0;
        gko_trace_one((GkoScene*)scene, &r, &h, any ? 1u : 0u);
        ++(any ? t_any : t_closest);
        return false;
    }
    uint CommittedStatus() const { return h.inst != 0xFFFFFFFFu ? COMMITTED_TRIANGLE_HIT : COMMITTED_NOTHING; }
    float CommittedRayT() const { return h.t; }
    float2 CommittedRayBarycentrics() const { return float2(h.u, h.v); }
    uint CommittedRayInstanceId() const { return h.inst; }  // the TLAS is built in node order (RayTraceBaseRenderer.cpp: one instance per NodeProxy)
    uint CommittedRayInstanceCustomIndex() const;           // = that node's instanceId
    uint CommittedPrimitiveIndex() const { return h.prim; }
    float4x3 CommittedRayWorldToObject() const {
        float m[12];
        gko_instance_w2o((GkoScene*)scene, h.inst, m);
        float4x3 M;
        for (int c = 0; c < 4; ++c) M.row[c] = float3(m[0 * 4 + c], m[1 * 4 + c], m[2 * 4 + c]);
        return M;
    }
};

// ---- interfaces (Shading.slang:256-276), declarations only
struct Vertex;
struct IPrimaryRayCaster {
    virtual bool TracePrimaryRay(int2 ipos, int2 isize, uint4& RandomSeed, Vertex& outVertex, NodeProxy& outNode, float3& outRayDir, float& outRayLength, float2& outPixelOffset) = 0;
};
struct IRayTracer {
    virtual bool TraceRay(float3 rayOrigin, float3 rayDir, float maxDistance, Vertex& outVertex, NodeProxy& outNode, float4& inoutRayColor) = 0;
    virtual bool TraceOcclusion(float3 rayOrigin, float3 rayDir) = 0;
    virtual bool TraceSegment(float3 rayOrigin, float3 rayTarget, float epsilon) = 0;
};
struct IDirectIlluminator {
    virtual void DirectIlluminate(uint4& RandomSeed, float3 position, float3 normal, float4& directIllumColor) = 0;
};

#include "PreProcessor.slang"
#include "const_func_gen.inc"
#include "general_func_gen.inc"
#include "ambient_cube_gen.inc"
#include "ambient_lookup_gen.inc"
template <int FLAGS> uint RayQuery<FLAGS>::CommittedRayInstanceCustomIndex() const { return Bindless.GetGpuscene().Nodes[h.inst].instanceId; }
#include "shading_gen.inc"
struct IProbeGenerator {  // AmbientCube.slang:464-468
    virtual void Render(IRayTracer& tracer, float3 origin, uint CubeIdx, uint4& RandomSeed) = 0;
    virtual void RenderAmbientCubeOnly(IRayTracer& tracer, float3 origin, uint CubeIdx, uint4& RandomSeed) = 0;
};
inline float lerp(float a, float b, float t) { return a + (b - a) * t; }
#include "probe_bake_gen.inc"
namespace pathtracing {
#include "pathtracing_gen.inc"
}
namespace bake {
#include "bake_gen.inc"
}
namespace bufferclear {
#include "bufferclear_gen.inc"
}
}  // namespace ref

using namespace ref;

extern "C" {
// images[rt] = the storage image in its reference format, indexed by the reference's RT_* numbers (NULL = unbound);
// images[2] = RT_MINIGBUFFER (R32_UINT), images[10] = RT_PREV_DEPTHBUFFER (R32F), images[14] = RT_MOTIONMOMENT (R16_UINT, read and written).
// counters[0] = closest-hit queries, counters[1] = any-hit queries.
// Rows [y0, y1) only (whole 8-row groups are not required: every thread of the two dispatches is independent), on `threads` host threads
// (0 = all cores): the timed CPU baseline of bench.py. Threads write disjoint pixels; the scene block and the images are read-only / per pixel.
void ref_pathtrace_rows(GkoScene* scene, const void* ubo, void* const* images, int w, int h, int y0, int y1, int threads, uint64_t* counters) {
    g_scene = scene;
    g_closest = 0;
    g_any = 0;
    const int f16[] = {1, 6, 7, 12};
    for (int i : f16) Bindless.f4[i] = ImgF4{(uint8_t*)images[i], w, h, FMT_F16X4, (float*)images[i + 32]};  // images[32 + rt]: optional fp32 shadow
    Bindless.f4[5] = ImgF4{(uint8_t*)images[5], w, h, FMT_RG32F};
    Bindless.u[2] = ImgU{(uint8_t*)images[2], w, h, 4};
    Bindless.u[3] = ImgU{(uint8_t*)images[3], w, h, 4};
    Bindless.u[14] = ImgU{(uint8_t*)images[14], w, h, 2};
    Bindless.f1[10] = ImgF1{(float*)images[10], w, h};
    UniformBufferObject cam;
    static_assert(sizeof(UniformBufferObject) == 784, "UniformBufferObject");
    static_assert(sizeof(Material) == 64 && sizeof(GPUVertex) == 24 && sizeof(NodeProxy) == 208 && sizeof(ModelData) == 64, "reference layouts");
    std::memcpy(&cam, ubo, sizeof(cam));
    GkoSceneArrays a;
    gko_scene_arrays(scene, &a);
    GpuSceneShim& S = Bindless.scene;
    S.Camera = &cam;
    S.Nodes = (NodeProxy*)a.nodes;
    S.Vertices = (GPUVertex*)a.vertices;
    S.Indices = (uint*)a.indices;
    S.Materials = (Material*)a.materials;
    S.Offsets = (ModelData*)a.offsets;
    S.Cubes = (AmbientCube*)a.cubes;
    S.Voxels = (VoxelData*)a.voxels;
    S.TLAS = scene;
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    // Work items are 8 x 8 pixel groups (the shader's numthreads) handed out by an atomic counter, so a band of a few rows still
    // keeps every host core busy; both dispatches of a group run back to back on one thread (its pixels are independent of the others').
    const int gx = (w + 7) / 8, gy0 = y0 / 8, gy1 = (y1 + 7) / 8;
    const int64_t total = (int64_t)gx * (gy1 - gy0);
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        t_closest = t_any = 0;
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= total) break;
            const int bx = (int)(i % gx) * 8, by = (gy0 + (int)(i / gx)) * 8;
            for (int pass = 0; pass < 2; ++pass)
                for (int y = std::max(by, y0); y < std::min(by + 8, y1); ++y)
                    for (int x = bx; x < bx + 8; ++x) {
                        if (pass == 0) bufferclear::bufferclear_main(uint3((uint)x, (uint)y, 0u));
                        else pathtracing::pathtracing_main(uint3((uint)x, (uint)y, 0u));
                    }
        }
        g_closest += t_closest;
        g_any += t_any;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    counters[0] = g_closest;
    counters[1] = g_any;
    S.Camera = nullptr;
}
// One frame: Util.BufferClear, then Core.PathTracing (PathTracingRenderer.cpp:96-124); both dispatches cover ceil(w/8) x ceil(h/8) groups of 8 x 8
void ref_pathtrace(GkoScene* scene, const void* ubo, void* const* images, int w, int h, uint64_t* counters) {
    ref_pathtrace_rows(scene, ubo, images, w, h, 0, (h + 7) / 8 * 8, 1, counters);
}

// Bake.HwAmbientCube.comp.slang over probes [first, first + count) of caller-provided probe arrays (updated in place), one thread after the
// other in index order (SceneList / Scene.cpp dispatches it with custom_data_0 = first)
void ref_probes_bake(GkoScene* scene, const void* ubo, void* cubes, void* voxels, uint32_t first, uint32_t count, uint64_t* counters) {
    g_scene = scene;
    t_closest = t_any = 0;  // single-threaded: the calling thread's counters are the totals
    UniformBufferObject cam;
    std::memcpy(&cam, ubo, sizeof(cam));
    GkoSceneArrays a;
    gko_scene_arrays(scene, &a);
    GpuSceneShim& S = Bindless.scene;
    S.Camera = &cam;
    S.Nodes = (NodeProxy*)a.nodes;
    S.Vertices = (GPUVertex*)a.vertices;
    S.Indices = (uint*)a.indices;
    S.Materials = (Material*)a.materials;
    S.Offsets = (ModelData*)a.offsets;
    S.Lights = (LightObject*)a.lights;
    S.HDRSHs = (SphericalHarmonics*)a.sky_sh;
    S.Cubes = (AmbientCube*)cubes;
    S.Voxels = (VoxelData*)voxels;
    S.TLAS = scene;
    S.custom_data_0 = first;
    static_assert(sizeof(LightObject) == 80 && sizeof(SphericalHarmonics) == 112 && sizeof(AmbientCube) == 56, "reference layouts");
    for (uint32_t i = 0; i < count; ++i) bake::bake_main(uint3(i, 0u, 0u));
    counters[0] = t_closest;
    counters[1] = t_any;
    S.Camera = nullptr;
}

// interpolateAmbientCubesV2<Full | DI> on caller-provided probe arrays; which: 0 DI, 2 Full
void ref_interpolate_ambient_cubes(int which, const void* cubes, const void* voxels, const float* pos, const float* n, float* out) {
    Bindless.scene.Cubes = (AmbientCube*)cubes;
    Bindless.scene.Voxels = (VoxelData*)voxels;
    float3 p(pos[0], pos[1], pos[2]), nn(n[0], n[1], n[2]);
    float4 r = which == 0 ? interpolateAmbientCubesV2<DIAmbientCubeSampler>(p, nn) : interpolateAmbientCubesV2<FullAmbientCubeSampler>(p, nn);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
}
