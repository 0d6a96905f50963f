// gk_raster.cu — the visibility front end of the reference frame: GPU-driven culling and the visibility-buffer pass.
//
//   "gpu cull"         assets/shaders/Task.GpuCull.comp.slang:199-227 (IsAABBInFrustum :127-197, IsOccludedVolume :56-115), dispatched at
//                      src/Rendering/VulkanBaseRenderer.cpp:603-647: per NodeProxy (= draw batch) a frustum test of the model's local box
//                      and an occlusion test of a 6 x 6 x 6 grid of box points against LAST frame's depth image, the result as one
//                      VkDrawIndexedIndirectCommand per batch plus the GPUDrivenStat counters the host reads back (Scene.cpp:412-419)
//   "visibility pass"  Rast.VisibilityPass.{vert,frag}.slang + vkCmdDrawIndexedIndirect (VulkanBaseRenderer.cpp:673-724): R32_UINT
//                      ((node + 1) & 0x7FFF) << 17 | (triangle & 0x1FFFF) of the nearest front-facing triangle under every pixel centre
//
// B200 has no rasteriser, so the second one is a software rasteriser following the Vulkan rasterisation rules the fixed-function unit
// implements: clip to z >= 0, perspective divide, viewport transform, 8 sub-pixel bits, back-face culling by the sign of the area
// (VK_CULL_MODE_BACK_BIT, VK_FRONT_FACE_COUNTER_CLOCKWISE, PipelineCommon/CommonComputePipeline.cpp:201-202), coverage at pixel centres
// with the top-left rule in exact integer arithmetic, depth = z/w interpolated in screen space, depth test LESS (:221). Fragments are
// resolved with one 64-bit atomicMin per covered pixel on (depth bits << 32 | id): the nearest wins, a depth tie goes to the smaller id, so
// the result does not depend on the order of the threads. The hardware's own snapping and tie rules are not published; the CPU checker
// restates THIS specification independently (tests only) and the two must agree bit for bit.
// The path tracer then reads the buffer exactly like FVisibilityBufferRayCasterV2 (Shading.slang:338-494) instead of casting a ray through
// the pixel centre (GkcuFrameParams.primary_visibility = 1). The format limits the mode to 32 767 proxies and 131 071 triangles per section.
#include <cfloat>
#include <cstdio>

#include "gk_context.h"
#include "gk_device.cuh"

using namespace gkd;

namespace gkcu {

namespace {

struct DrawCmd { uint32_t indexCount, instanceCount, firstIndex; int32_t vertexOffset; uint32_t firstInstance; };  // VkDrawIndexedIndirectCommand
static_assert(sizeof(DrawCmd) == 20, "VkDrawIndexedIndirectCommand");

// mul(A, B) of two column-major matrices, term order k = 0..3
GK_DEV GkMat4 matmul(const GkMat4& a, const GkMat4& b) {
    GkMat4 r;
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row)
            r.m[c][row] = a.m[0][row] * b.m[c][0] + a.m[1][row] * b.m[c][1] + a.m[2][row] * b.m[c][2] + a.m[3][row] * b.m[c][3];
    return r;
}

GK_DEV void box_corner(float3 mn, float3 mx, int i, float3& p) { p = f3(i & 1 ? mx.x : mn.x, i & 2 ? mx.y : mn.y, i & 4 ? mx.z : mn.z); }

// Task.GpuCull.comp.slang:127-197: the box is outside when all eight corners are outside one clip plane
GK_DEV bool IsAABBInFrustum(float3 mn, float3 mx, const GkMat4& wvp) {
    float4 clip[8];
    for (int i = 0; i < 8; ++i) {
        float3 p;
        box_corner(mn, mx, i, p);
        clip[i] = mul(wvp, f4(p.x, p.y, p.z, 1.0f));
    }
    bool out;
    out = true; for (int i = 0; i < 8; ++i) if (clip[i].x >= -clip[i].w) { out = false; break; } if (out) return false;
    out = true; for (int i = 0; i < 8; ++i) if (clip[i].x <= clip[i].w) { out = false; break; } if (out) return false;
    out = true; for (int i = 0; i < 8; ++i) if (clip[i].y >= -clip[i].w) { out = false; break; } if (out) return false;
    out = true; for (int i = 0; i < 8; ++i) if (clip[i].y <= clip[i].w) { out = false; break; } if (out) return false;
    out = true; for (int i = 0; i < 8; ++i) if (clip[i].z >= -clip[i].w) { out = false; break; } if (out) return false;
    out = true; for (int i = 0; i < 8; ++i) if (clip[i].z <= clip[i].w) { out = false; break; } if (out) return false;
    return true;
}

// :56-115: occluded only if every point of a 6 x 6 x 6 grid over the box projects into the image and lies behind last frame's depth
GK_DEV bool IsOccludedVolume(float3 mn, float3 mx, const GkMat4& wvp, const float* __restrict__ prev_depth, uint32_t w, uint32_t h) {
    const int GRID_SIZE = 6;
    const float GRID_SIZE_MINUS_ONE = GRID_SIZE - 1;
    for (int x = 0; x < GRID_SIZE; x++)
        for (int y = 0; y < GRID_SIZE; y++)
            for (int z = 0; z < GRID_SIZE; z++) {
                const float3 gridPoint = f3(mn.x + (mx.x - mn.x) * (float(x) / GRID_SIZE_MINUS_ONE), mn.y + (mx.y - mn.y) * (float(y) / GRID_SIZE_MINUS_ONE),
                                            mn.z + (mx.z - mn.z) * (float(z) / GRID_SIZE_MINUS_ONE));
                const float4 clipPos = mul(wvp, f4(gridPoint.x, gridPoint.y, gridPoint.z, 1.0f));
                const float3 ndc = f3(clipPos.x / clipPos.w, clipPos.y / clipPos.w, clipPos.z / clipPos.w);
                if (fabsf(ndc.x) > 1.0f || fabsf(ndc.y) > 1.0f || ndc.z < 0.0f || ndc.z > 1.0f) return false;
                const float2 screenUV = make_float2(ndc.x * 0.5f + 0.5f, ndc.y * 0.5f + 0.5f);
                const int px = (int)(screenUV.x * float(w)), py = (int)(screenUV.y * float(h));
                if (px < 0 || px >= (int)w || py < 0 || py >= (int)h) return false;
                const float previousDepth = prev_depth[(size_t)py * w + px];
                const float currentDepth = ndc.z - 0.0001f;
                if (previousDepth > currentDepth) return false;
            }
    return true;
}

// Task.GpuCull main, :199-227. One thread per NodeProxy; the reference rounds the dispatch up to 64 and lets the extra threads read stale
// proxies, here they do nothing.
__global__ void k_gpu_cull(const GkNodeProxy* __restrict__ nodes, uint32_t n, const GkModelData* __restrict__ offsets, const __grid_constant__ GkUniformBufferObject cam,
                           const float* __restrict__ prev_depth, uint32_t w, uint32_t h, DrawCmd* __restrict__ cmds, GkGPUDrivenStat* stats) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GkNodeProxy& node = nodes[i];
    const GkModelData model = offsets[node.modelId];
    const GkMat4 mvp = matmul(cam.ViewProjection, node.worldTS);
    const float3 mn = f3(model.localAabbMin.x, model.localAabbMin.y, model.localAabbMin.z), mx = f3(model.localAabbMax.x, model.localAabbMax.y, model.localAabbMax.z);
    bool shouldDraw = (node.visible > 0) && IsAABBInFrustum(mn, mx, mvp);
    if (shouldDraw) {
        atomicAdd(&stats->ProcessedCount, 1u);
        atomicAdd(&stats->TriangleCount, model.indexCount / 3);
        if (IsOccludedVolume(mn, mx, mvp, prev_depth, w, h)) {
            atomicAdd(&stats->CulledCount, 1u);
            atomicAdd(&stats->CulledTriangleCount, model.indexCount / 3);
            shouldDraw = false;
        }
    }
    DrawCmd c;
    c.instanceCount = shouldDraw ? 1 : 0;
    c.firstInstance = shouldDraw ? i : 0;
    c.indexCount = shouldDraw ? model.indexCount : 0;
    c.firstIndex = shouldDraw ? model.indexOffset : 0;
    c.vertexOffset = shouldDraw ? (int32_t)model.reorderOffset : 0;
    cmds[i] = c;
}

// ------------------------------------------------------------------------------------------------ rasteriser
#define GK_RASTER_SUBPIXEL 256
#define GK_RASTER_CLAMP (1ll << 30)
#define GK_RASTER_SMALL 256  // a triangle whose pixel box is larger goes to the second kernel (a block per triangle)

struct ScreenTri {
    int32_t x[3], y[3];  // fixed point, 8 sub-pixel bits, vertex order such that the doubled area is positive
    float z[3];
    uint32_t id;
};

GK_DEV long long snap(float v) {
    const float s = floorf(v * (float)GK_RASTER_SUBPIXEL + 0.5f);
    if (!(s > -(float)GK_RASTER_CLAMP)) return -GK_RASTER_CLAMP;  // also NaN
    if (!(s < (float)GK_RASTER_CLAMP)) return GK_RASTER_CLAMP;
    return (long long)s;
}

// perspective divide + viewport transform + snapping + back-face culling; false if nothing can be covered
GK_DEV bool setup_triangle(const float4* c, uint32_t w, uint32_t h, uint32_t id, ScreenTri& T, int& x0, int& x1, int& y0, int& y1) {
    long long X[3], Y[3];
    float Z[3];
    for (int k = 0; k < 3; ++k) {
        const float nx = c[k].x / c[k].w, ny = c[k].y / c[k].w;
        Z[k] = c[k].z / c[k].w;
        X[k] = snap((nx + 1.0f) * ((float)w * 0.5f));
        Y[k] = snap((ny + 1.0f) * ((float)h * 0.5f));
    }
    const long long area2 = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0]);
    // Vulkan: a = -1/2 sum(x_i y_i+1 - x_i+1 y_i) = -area2 / 2 in framebuffer coordinates; counter-clockwise front faces have a > 0
    if (area2 >= 0) return false;  // back-facing or degenerate
    // swap two vertices so that the edge functions are positive inside
    T.x[0] = (int32_t)X[0]; T.y[0] = (int32_t)Y[0]; T.z[0] = Z[0];
    T.x[1] = (int32_t)X[2]; T.y[1] = (int32_t)Y[2]; T.z[1] = Z[2];
    T.x[2] = (int32_t)X[1]; T.y[2] = (int32_t)Y[1]; T.z[2] = Z[1];
    T.id = id;
    const long long mnx = min(X[0], min(X[1], X[2])), mxx = max(X[0], max(X[1], X[2])), mny = min(Y[0], min(Y[1], Y[2])), mxy = max(Y[0], max(Y[1], Y[2]));
    // pixel p has its centre at p * 256 + 128
    auto ceil_div = [](long long a) { return (a >= 0) ? (a + 255) / 256 : -((-a) / 256); };
    auto floor_div = [](long long a) { return (a >= 0) ? a / 256 : -((-a + 255) / 256); };
    const long long px0 = max(0ll, ceil_div(mnx - 128)), px1 = min((long long)w - 1, floor_div(mxx - 128));
    const long long py0 = max(0ll, ceil_div(mny - 128)), py1 = min((long long)h - 1, floor_div(mxy - 128));
    if (px0 > px1 || py0 > py1) return false;
    x0 = (int)px0; x1 = (int)px1; y0 = (int)py0; y1 = (int)py1;
    return true;
}

// coverage (top-left rule) + depth of one pixel centre; writes the fragment with a 64-bit atomicMin
GK_DEV void shade_pixel(const ScreenTri& T, int px, int py, unsigned long long* __restrict__ keys, uint32_t w) {
    const long long Px = (long long)px * 256 + 128, Py = (long long)py * 256 + 128;
    long long E[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // edge i runs from vertex (i + 1) % 3 to (i + 2) % 3 and is the weight of vertex i
        const int a = (i + 1) % 3, b = (i + 2) % 3;
        const long long dx = (long long)T.x[b] - T.x[a], dy = (long long)T.y[b] - T.y[a];
        E[i] = dx * (Py - T.y[a]) - dy * (Px - T.x[a]);
        const bool topleft = dy < 0 || (dy == 0 && dx > 0);  // y grows downwards, interior on the positive side
        if (E[i] < 0 || (E[i] == 0 && !topleft)) return;
    }
    const double area = (double)(E[0] + E[1] + E[2]);
    const double zs = __dadd_rn(__dadd_rn(__dmul_rn((double)E[0], (double)T.z[0]), __dmul_rn((double)E[1], (double)T.z[1])), __dmul_rn((double)E[2], (double)T.z[2]));
    const float z = (float)__ddiv_rn(zs, area);
    if (!(z >= 0.0f && z < 1.0f)) return;  // depth clip; the far value 1.0 never passes LESS against the cleared buffer
    atomicMin(keys + (size_t)py * w + px, ((unsigned long long)__float_as_uint(z) << 32) | T.id);
}

struct LargeList { ScreenTri* tris; uint32_t* count; uint32_t capacity; };

GK_DEV void raster_clip_triangle(const float4* c, uint32_t w, uint32_t h, uint32_t id, unsigned long long* __restrict__ keys, const LargeList& L) {
    ScreenTri T;
    int x0, x1, y0, y1;
    if (!setup_triangle(c, w, h, id, T, x0, x1, y0, y1)) return;
    uint32_t slot = 0xFFFFFFFFu;
    if ((long long)(x1 - x0 + 1) * (y1 - y0 + 1) > GK_RASTER_SMALL) slot = atomicAdd(L.count, 1u);
    if (slot < L.capacity) {
        L.tris[slot] = T;
    } else {  // small, or the list of large triangles is full: this thread covers the box itself
        for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) shade_pixel(T, x, y, keys, w);
    }
}

// one triangle in clip space: clip against the near plane z >= 0 (Sutherland-Hodgman in vertex order), then rasterise the fan
GK_DEV void raster_triangle(const float4* v, uint32_t w, uint32_t h, uint32_t id, unsigned long long* __restrict__ keys, const LargeList& L) {
    const bool in0 = v[0].z >= 0.0f, in1 = v[1].z >= 0.0f, in2 = v[2].z >= 0.0f;
    if (in0 && in1 && in2) { raster_clip_triangle(v, w, h, id, keys, L); return; }
    if (!in0 && !in1 && !in2) return;
    float4 poly[4];
    int n = 0;
    const bool in[3] = {in0, in1, in2};
    for (int k = 0; k < 3; ++k) {
        const float4 a = v[k], b = v[(k + 1) % 3];
        if (in[k]) poly[n++] = a;
        if (in[k] != in[(k + 1) % 3]) {
            const float t = a.z / (a.z - b.z);
            poly[n++] = f4(a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, 0.0f, a.w + (b.w - a.w) * t);
        }
    }
    const float4 t0[3] = {poly[0], poly[1], poly[2]};
    raster_clip_triangle(t0, w, h, id, keys, L);
    if (n == 4) {
        const float4 t1[3] = {poly[0], poly[2], poly[3]};
        raster_clip_triangle(t1, w, h, id, keys, L);
    }
}

__global__ void k_raster_clear(unsigned long long* keys, size_t n, uint32_t* large_count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = 0x3F800000ull << 32;  // depth 1.0, colour 0 (VulkanBaseRenderer.cpp:676-678)
    if (i < 2) large_count[i] = 0;
}

// vkCmdDrawIndexedIndirect: one block per draw command (= batch = NodeProxy), the threads stride over its triangles.
// Rast.VisibilityPass.vert.slang:9-23 per vertex.
__global__ void __launch_bounds__(128) k_raster_batches(const DrawCmd* __restrict__ cmds, const GkNodeProxy* __restrict__ nodes, const __grid_constant__ GkUniformBufferObject cam,
                                                        const uint32_t* __restrict__ raster_indices, const uint32_t* __restrict__ reorders, const uint16_t* __restrict__ simple,
                                                        unsigned long long* __restrict__ keys, uint32_t w, uint32_t h, LargeList L) {
    const DrawCmd cmd = cmds[blockIdx.x];
    if (cmd.instanceCount == 0 || cmd.indexCount == 0) return;
    __shared__ GkMat4 s_mvp;
    if (threadIdx.x == 0) s_mvp = matmul(cam.ViewProjection, nodes[cmd.firstInstance].worldTS);
    __syncthreads();
    const uint32_t instanceIdx = cmd.firstInstance + 1;
    for (uint32_t t = threadIdx.x; t < cmd.indexCount / 3; t += blockDim.x) {
        float4 c[3];
        uint32_t triangleIdx = 0;
        for (int k = 0; k < 3; ++k) {
            const uint32_t index = __ldg(raster_indices + cmd.firstIndex + 3 * t + k);
            const uint32_t vertexIndex = index + (uint32_t)cmd.vertexOffset;
            if (k == 0) triangleIdx = vertexIndex - (uint32_t)cmd.vertexOffset;  // the provoking (first) vertex carries the flat attribute
            const uint32_t vi = __ldg(reorders + vertexIndex);
            const float3 p = f3(h2f(simple[(size_t)vi * 4]), h2f(simple[(size_t)vi * 4 + 1]), h2f(simple[(size_t)vi * 4 + 2]));
            c[k] = mul(s_mvp, f4(p.x, p.y, p.z, 1.0f));
        }
        const uint32_t id = ((instanceIdx & 0x7FFFu) << 17) | (triangleIdx & 0x1FFFFu);
        raster_triangle(c, w, h, id, keys, L);
    }
}

// the few triangles that cover many pixels (a floor, a wall): a block per triangle, the threads stride over its pixel box
__global__ void __launch_bounds__(256) k_raster_large(LargeList L, unsigned long long* __restrict__ keys, uint32_t w, uint32_t h) {
    const uint32_t n = min(*L.count, L.capacity);
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const ScreenTri T = L.tris[i];
        long long mnx = min(T.x[0], min(T.x[1], T.x[2])), mxx = max(T.x[0], max(T.x[1], T.x[2])), mny = min(T.y[0], min(T.y[1], T.y[2])), mxy = max(T.y[0], max(T.y[1], T.y[2]));
        auto ceil_div = [](long long a) { return (a >= 0) ? (a + 255) / 256 : -((-a) / 256); };
        auto floor_div = [](long long a) { return (a >= 0) ? a / 256 : -((-a + 255) / 256); };
        const int x0 = (int)max(0ll, ceil_div(mnx - 128)), x1 = (int)min((long long)w - 1, floor_div(mxx - 128));
        const int y0 = (int)max(0ll, ceil_div(mny - 128)), y1 = (int)min((long long)h - 1, floor_div(mxy - 128));
        const int bw = x1 - x0 + 1;
        const long long total = (long long)bw * (y1 - y0 + 1);
        for (long long p = threadIdx.x; p < total; p += blockDim.x) shade_pixel(T, x0 + (int)(p % bw), y0 + (int)(p / bw), keys, w);
    }
}

__global__ void k_raster_resolve(const unsigned long long* __restrict__ keys, uint32_t* __restrict__ vis, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vis[i] = (uint32_t)(keys[i] & 0xFFFFFFFFull);
}

}  // namespace

int run_gpu_cull(Context& c, const GkUniformBufferObject& ubo) {
    const uint32_t n = c.node_count;
    if (!c.draw_cmds.reserve(std::max<size_t>(1, n) * sizeof(DrawCmd)) || !c.gpu_stats.reserve(sizeof(GkGPUDrivenStat))) return 1;
    GK_CUDA(cudaMemsetAsync(c.gpu_stats.p, 0, sizeof(GkGPUDrivenStat), c.stream));  // Scene::UpdateNodes zeroes the counters after reading them (Scene.cpp:417)
    if (n)
        k_gpu_cull<<<(n + 63) / 64, 64, 0, c.stream>>>(c.nodes.as<GkNodeProxy>(), n, c.offsets.as<GkModelData>(), ubo, c.rt.prev_depth, c.rt.w, c.rt.h,
                                                      c.draw_cmds.as<DrawCmd>(), c.gpu_stats.as<GkGPUDrivenStat>());
    c.launches++;
    c.draw_count = n;
    GK_CUDA(cudaGetLastError());
    return 0;
}

int run_visibility_pass(Context& c, const GkUniformBufferObject& ubo) {
    if (!c.raster_indices.p || !c.reorders.p) { set_error("visibility pass: call gkcu_scene_upload_raster first"); return 1; }
    if (c.draw_count != c.node_count) { set_error("visibility pass: run the gpu cull first (gkcu_cull)"); return 1; }
    if (c.node_count > 32767) { set_error("visibility pass: the 15-bit instance field of the visibility buffer holds at most 32 767 proxies; use primary_visibility = 0"); return 1; }
    for (uint32_t i = 0; i < c.offset_count; ++i)
        if (c.host_offsets[i].indexCount / 3 > 131071) { set_error("visibility pass: a section has more than 131 071 triangles"); return 1; }
    const size_t n = (size_t)c.rt.w * c.rt.h;
    const uint32_t large_cap = 1u << 18;
    if (!c.vis_keys.reserve(n * 8) || !c.raster_large.reserve((size_t)large_cap * sizeof(ScreenTri)) || !c.raster_counters.reserve(16)) return 1;
    LargeList L{c.raster_large.as<ScreenTri>(), c.raster_counters.as<uint32_t>(), large_cap};
    k_raster_clear<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(c.vis_keys.as<unsigned long long>(), n, L.count);
    if (c.node_count)
        k_raster_batches<<<c.node_count, 128, 0, c.stream>>>(c.draw_cmds.as<DrawCmd>(), c.nodes.as<GkNodeProxy>(), ubo, c.raster_indices.as<uint32_t>(), c.reorders.as<uint32_t>(),
                                                             c.simple.as<uint16_t>(), c.vis_keys.as<unsigned long long>(), c.rt.w, c.rt.h, L);
    k_raster_large<<<c.sm_count * 4, 256, 0, c.stream>>>(L, c.vis_keys.as<unsigned long long>(), c.rt.w, c.rt.h);
    k_raster_resolve<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(c.vis_keys.as<unsigned long long>(), c.rt.minigbuffer, n);
    c.launches += 4;
    GK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gkcu
