// gk_api.cu — the C ABI of libgknext_cuda.so (include/gkcu.h). Thin: owns device memory, orders the passes of a frame
// on one CUDA stream, converts failures into return codes + gkcu_last_error(). There is no CPU fallback anywhere: every
// entry point either runs the sm_100a kernels or fails.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "gk_context.h"

struct GkcuContext {
    gkcu::Context c;
};

namespace gkcu {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
bool check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    g_error = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}

bool DevBuf::reserve(size_t bytes) {
    if (bytes <= cap && p) return true;
    if (bytes == 0) bytes = 16;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    if (!check(cudaMalloc(&p, want), "cudaMalloc")) { p = nullptr; return false; }
    cap = want;
    return true;
}
void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

static int alloc_zero(void** p, size_t bytes, cudaStream_t st) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    GK_CUDA(cudaMalloc(p, bytes));
    GK_CUDA(cudaMemsetAsync(*p, 0, bytes, st));
    return 0;
}

namespace {
__global__ void k_publish(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, uint32_t words) {
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
}
}  // namespace
int publish_words(Context& c, const void* src, uint32_t dst_word, uint32_t words, cudaStream_t st) {
    k_publish<<<1, 64, 0, st>>>(static_cast<const uint32_t*>(src), c.pub_dev + dst_word, words);
    c.launches++;
    GK_CUDA(cudaGetLastError());
    return 0;
}

static void free_targets(RenderTargets& t) {
    void** ptrs[] = {(void**)&t.accum_diffuse[0], (void**)&t.accum_diffuse[1], (void**)&t.accum_specular[0], (void**)&t.accum_specular[1],
                     (void**)&t.accum_albedo[0], (void**)&t.accum_albedo[1], (void**)&t.single_diffuse, (void**)&t.single_specular,
                     (void**)&t.albedo, (void**)&t.normal, (void**)&t.denoised, (void**)&t.objectid[0], (void**)&t.objectid[1],
                     (void**)&t.primitiveid, (void**)&t.motion, (void**)&t.prev_depth, (void**)&t.moment, (void**)&t.minigbuffer};
    for (void** p : ptrs) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    t.w = t.h = 0;
}

static int upload_tex_table(Context& c) {
    struct TexDescHost { const float4* texels; uint32_t w, h, pad; };
    std::vector<TexDescHost> tab(c.textures.size());
    for (size_t i = 0; i < c.textures.size(); ++i) tab[i] = {c.textures[i].texels, c.textures[i].w, c.textures[i].h, 0};
    if (!c.tex_table.reserve(std::max<size_t>(1, tab.size()) * sizeof(TexDescHost))) return 1;
    if (!tab.empty()) GK_CUDA(cudaMemcpyAsync(c.tex_table.p, tab.data(), tab.size() * sizeof(TexDescHost), cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    c.tex_count = (uint32_t)tab.size();
    return 0;
}

}  // namespace gkcu

using namespace gkcu;

// A failed call of the previous entry point (or of another library on this thread) can leave a non-sticky CUDA error behind; cub and the
// GK_CUDA checks would report it as their own. Every entry point starts clean; GKCU_DEBUG=1 names the call that left it.
static void clear_stale_error(Context& c, const char* api) {
    const cudaError_t stale = cudaGetLastError();
    if (stale != cudaSuccess) {
        static const bool dbg = getenv("GKCU_DEBUG") != nullptr;
        if (dbg) fprintf(stderr, "gkcu: %s entered with the stale CUDA error '%s' left by %s\n", api, cudaGetErrorString(stale), c.last_api ? c.last_api : "(nothing)");
    }
    c.last_api = api;
}

#define CTX_OR_FAIL(ctx)                                   \
    if (!(ctx)) { set_error("null context"); return 1; }   \
    Context& c = (ctx)->c;                                 \
    if (!check(cudaSetDevice(c.device), "cudaSetDevice")) return 1; \
    clear_stale_error(c, __func__);

extern "C" {

const char* gkcu_last_error(void) { return g_error.c_str(); }

int gkcu_create(int device_ordinal, GkcuContext** out_ctx) {
    if (!out_ctx) { set_error("gkcu_create: out_ctx is null"); return 1; }
    *out_ctx = nullptr;
    int count = 0;
    if (!check(cudaGetDeviceCount(&count), "cudaGetDeviceCount") || count == 0) {
        if (g_error.empty()) set_error("no CUDA device visible: libgknext_cuda has no CPU fallback");
        return 1;
    }
    if (device_ordinal < 0 || device_ordinal >= count) { set_error("gkcu_create: bad device ordinal"); return 1; }
    GK_CUDA(cudaSetDevice(device_ordinal));
    cudaDeviceProp prop;
    GK_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
    if (prop.major < 10) {
        char msg[160];
        snprintf(msg, sizeof(msg), "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device_ordinal, prop.major, prop.minor);
        set_error(msg);
        return 1;
    }
    GkcuContext* ctx = new GkcuContext();
    Context& c = ctx->c;
    c.device = device_ordinal;
    c.sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("GKCU_TRACE_VARIANT")) c.trace_variant = atoi(e);
    if (const char* e = getenv("GKCU_TMA")) c.use_tma = atoi(e);
    if (const char* e = getenv("GKCU_FAST_STENCIL")) c.fast_stencil = atoi(e) != 0;
    if (const char* e = getenv("GKCU_SHADE_MINB")) c.shade_minb = atoi(e);
    if (const char* e = getenv("GKCU_TLAS_REFIT")) c.tlas_refit = atoi(e) != 0;
    if (const char* e = getenv("GKCU_TLAS_REBUILD_EVERY")) c.tlas_rebuild_every = std::max(0, atoi(e));
    if (const char* e = getenv("GKCU_TRACE_RPB")) c.trace_rpb = std::max(32, std::min(65536, atoi(e)));
    if (const char* e = getenv("GKCU_TRACE_GRID")) c.trace_grid_mult = std::max(1, std::min(64, atoi(e)));
    if (const char* e = getenv("GKCU_MATCH_CARVEOUT")) c.match_carveout = atoi(e) != 0;
    if (const char* e = getenv("GKCU_TRACE_CARVEOUT")) c.trace_carveout = std::max(-1, std::min(100, atoi(e)));
    if (const char* e = getenv("GKCU_SORT_RAYS")) c.sort_rays = std::max(0, std::min(19, atoi(e)));
    if (const char* e = getenv("GKCU_SORT_EXTENT")) c.sort_extent = (float)atof(e);
    if (const char* e = getenv("GKCU_REFILL")) c.refill_below = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("GKCU_NODE_FORMAT")) c.node_format = atoi(e) == 4 ? 4 : 8;
    if (!check(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking), "cudaStreamCreate")) { delete ctx; return 1; }
    if (!check(cudaHostAlloc((void**)&c.pub_host, 64 * 5 * 4, cudaHostAllocMapped), "cudaHostAlloc") ||
        !check(cudaHostGetDevicePointer((void**)&c.pub_dev, c.pub_host, 0), "cudaHostGetDevicePointer")) { delete ctx; return 1; }
    memset(c.pub_host, 0, 64 * 5 * 4);
    for (auto& e : c.ev) cudaEventCreate(&e);
    for (auto& e : c.kev) cudaEventCreate(&e);
    for (auto& l : c.wev)
        for (auto& e : l) cudaEventCreate(&e);
    for (auto& l : c.lane_ev)
        for (auto& e : l) cudaEventCreate(&e);
    for (auto& l : c.pev)
        for (auto& e : l) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    if (const char* e = getenv("GKCU_KERNEL_TIMING")) c.kernel_timing = atoi(e) != 0;
    if (const char* e = getenv("GKCU_LANES")) c.lanes = std::max(1, std::min(4, atoi(e)));
    // lane streams only when the frame is split into lanes: every stream takes one of the device's hardware queues (8 by default), and two
    // streams that share a queue serialise - fatal for kernels that wait for a kernel of another context (the multi-GPU device barrier)
    for (int l = 0; l < 4 && c.lanes > 1; ++l) cudaStreamCreateWithFlags(&c.lane_stream[l], cudaStreamNonBlocking);
    if (const char* e = getenv("GKCU_TAIL_SEGMENTS")) {
        c.tail_segments = 0;
        for (const char* q = e; *q && c.tail_segments < 14;) {
            const int v = atoi(q);
            if (v > 0) c.tail_segment_rays[c.tail_segments++] = v;
            while (*q && *q != ',' && *q != ':') ++q;
            if (*q) ++q;
        }
    }
    if (const char* e = getenv("GKCU_TAIL_SORT")) c.tail_sort = atoi(e) != 0;
    if (const char* e = getenv("GKCU_TAIL_SOLO")) c.tail_solo = std::max(0, std::min(1 << 20, atoi(e)));
    if (const char* e = getenv("GKCU_TAIL_SPREAD")) c.tail_spread = std::max(1, std::min(32, atoi(e)));
    if (const char* e = getenv("GKCU_TAIL_FUSED")) c.tail_fused = atoi(e) != 0;
    if (const char* e = getenv("GKCU_TAIL_CAP")) c.tail_cap = std::max(128, std::min(1 << 20, atoi(e)));
    if (const char* e = getenv("GKCU_TAIL_DIV")) c.tail_div = std::max(1, std::min(65536, atoi(e)));
    if (const char* e = getenv("GKCU_WAVE_CHUNK")) c.wave_chunk = std::max(1, std::min(7, atoi(e)));
    if (!c.ubo.reserve(sizeof(GkUniformBufferObject))) { delete ctx; return 1; }
    *out_ctx = ctx;
    return 0;
}

static void drop_peers(Context& c);

int gkcu_destroy(GkcuContext* ctx) {
    if (!ctx) return 0;
    Context& c = ctx->c;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    drop_peers(c);
    gkcu_unbind_external_image(ctx, GK_RT_DENOISED);
    gkcu_bind_external_semaphore(ctx, GKCU_EXT_OPAQUE_FD, -1);
    path_release(c);
    free_targets(c.rt);
    for (auto& t : c.textures)
        if (t.texels) cudaFree(t.texels);
    DevBuf* bufs[] = {&c.vertices, &c.simple, &c.indices, &c.offsets, &c.materials, &c.nodes, &c.blas_nodes, &c.blas_node_lo, &c.blas_node_hi,
                      &c.blas_tris, &c.blas_tri_prim, &c.blas_table, &c.shade_tris, &c.shade_inst, &c.tlas_nodes, &c.tlas_node_lo, &c.tlas_node_hi, &c.blas_leaf, &c.tlas_leaf, &c.raster_indices, &c.reorders, &c.draw_cmds, &c.gpu_stats, &c.vis_keys, &c.raster_large, &c.raster_counters, &c.instances, &c.inst_lo,
                      &c.inst_hi, &c.tlas_active_ids, &c.inst_flags, &c.inst_flags_prev, &c.select_tmp, &c.select_count, &c.node_stage, &c.node_stage_idx, &c.tex_table, &c.probes, &c.probe_meta, &c.pages, &c.cube_snapshot, &c.lights, &c.sky_sh, &c.ubo, &c.trace_rays, &c.trace_hits,
                      &c.peer_flags};
    for (DevBuf* b : bufs) b->release();
    BuildScratch& s = c.scratch;
    DevBuf* sb[] = {&s.prim_lo, &s.prim_hi, &s.keys, &s.keys_sorted, &s.order, &s.order_tmp, &s.sort_tmp, &s.n2_lo, &s.n2_hi, &s.n2_left, &s.n2_right,
                    &s.n2_parent, &s.n2_first, &s.n2_count, &s.n2_flag, &s.queue_a, &s.queue_b, &s.counters, &s.bounds, &s.bins, &s.pos_node,
                    &s.pos_node_tmp, &s.active, &s.cursors};
    for (DevBuf* b : sb) b->release();
    for (auto& e : c.ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : c.kev)
        if (e) cudaEventDestroy(e);
    for (auto& l : c.wev)
        for (auto& e : l)
            if (e) cudaEventDestroy(e);
    for (auto& l : c.pev)
        for (auto& e : l)
            if (e) cudaEventDestroy(e);
    if (c.pub_host) cudaFreeHost(c.pub_host);
    for (auto& l : c.lane_ev)
        for (auto& e : l)
            if (e) cudaEventDestroy(e);
    for (auto& st : c.lane_stream)
        if (st) cudaStreamDestroy(st);
    if (c.copy_stream) { cudaStreamSynchronize(c.copy_stream); cudaStreamDestroy(c.copy_stream); cudaEventDestroy(c.copy_ready); cudaEventDestroy(c.copy_done); }
    cudaStreamDestroy(c.stream);
    delete ctx;
    return 0;
}

int gkcu_resize(GkcuContext* ctx, uint32_t width, uint32_t height) {
    CTX_OR_FAIL(ctx);
    if (width == 0 || height == 0 || (uint64_t)width * height >= (1ull << 31)) { set_error("gkcu_resize: bad extent"); return 1; }
    GK_CUDA(cudaStreamSynchronize(c.stream));
    if (c.copy_stream) GK_CUDA(cudaStreamSynchronize(c.copy_stream));
    c.copy_pending = false;
    drop_peers(c);
    gkcu_unbind_external_image(ctx, GK_RT_DENOISED);  // the bound image had the old extent
    RenderTargets& t = c.rt;
    free_targets(t);
    size_t n = (size_t)width * height;
    cudaStream_t st = c.stream;
    for (int k = 0; k < 2; ++k) {
        if (alloc_zero((void**)&t.accum_diffuse[k], n * 8, st) || alloc_zero((void**)&t.accum_specular[k], n * 8, st) ||
            alloc_zero((void**)&t.accum_albedo[k], n * 8, st) || alloc_zero((void**)&t.objectid[k], n * 4, st))
            return 1;
    }
    if (alloc_zero((void**)&t.single_diffuse, n * 8, st) || alloc_zero((void**)&t.single_specular, n * 8, st) || alloc_zero((void**)&t.albedo, n * 8, st) ||
        alloc_zero((void**)&t.normal, n * 8, st) || alloc_zero((void**)&t.denoised, n * 8, st) || alloc_zero((void**)&t.primitiveid, n * 4, st) ||
        alloc_zero((void**)&t.motion, n * 8, st) || alloc_zero((void**)&t.prev_depth, n * 4, st) || alloc_zero((void**)&t.moment, n * 2, st) ||
        alloc_zero((void**)&t.minigbuffer, n * 4, st))
        return 1;
    t.w = width;
    t.h = height;
    t.cur = 0;
    if (path_resize(c)) return 1;
    GK_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int gkcu_scene_upload(GkcuContext* ctx, const GkcuSceneDesc* d) {
    CTX_OR_FAIL(ctx);
    if (!d || !d->vertices || !d->simple_vertices || !d->indices || !d->offsets) { set_error("gkcu_scene_upload: null array"); return 1; }
    if (d->offset_count != d->model_count * 10) { set_error("gkcu_scene_upload: offsets must hold 10 sections per model"); return 1; }
    // validate everything before the context is touched: a rejected upload leaves the previous scene (and its BLAS) intact
    for (uint32_t m = 0; m < d->model_count; ++m) {
        const GkModelData& md = d->offsets[m * 10];
        uint64_t idx = 0;
        for (int s = 0; s < 10; ++s) idx += d->offsets[m * 10 + s].indexCount;
        if ((uint64_t)md.indexOffset + idx > d->index_count || (uint64_t)md.vertexOffset + md.vertexCount > d->vertex_count) {
            set_error("gkcu_scene_upload: ModelData ranges exceed the buffers");
            return 1;
        }
        for (uint64_t i = 0; i < idx; ++i)
            if (d->indices[md.indexOffset + i] >= md.vertexCount) { set_error("gkcu_scene_upload: index beyond the model's vertex count"); return 1; }
    }
    cudaStream_t st = c.stream;
    c.raster_index_count = c.reorder_count = 0;  // the raster-side buffers belong to the previous scene
    c.blas_built = false;  // from here on the old acceleration structures no longer describe the buffers
    c.tlas_topo_valid = false;
    c.tlas_dirty = true;
    c.tlas.node_count = 0;
    if (!c.vertices.reserve((size_t)d->vertex_count * sizeof(GkGPUVertex)) || !c.simple.reserve((size_t)d->vertex_count * 8) ||
        !c.indices.reserve((size_t)d->index_count * 4) || !c.offsets.reserve((size_t)d->offset_count * sizeof(GkModelData)))
        return 1;
    GK_CUDA(cudaMemcpyAsync(c.vertices.p, d->vertices, (size_t)d->vertex_count * sizeof(GkGPUVertex), cudaMemcpyHostToDevice, st));
    GK_CUDA(cudaMemcpyAsync(c.simple.p, d->simple_vertices, (size_t)d->vertex_count * 8, cudaMemcpyHostToDevice, st));
    GK_CUDA(cudaMemcpyAsync(c.indices.p, d->indices, (size_t)d->index_count * 4, cudaMemcpyHostToDevice, st));
    GK_CUDA(cudaMemcpyAsync(c.offsets.p, d->offsets, (size_t)d->offset_count * sizeof(GkModelData), cudaMemcpyHostToDevice, st));
    GK_CUDA(cudaStreamSynchronize(st));
    c.vertex_count = d->vertex_count;
    c.index_count = d->index_count;
    c.offset_count = d->offset_count;
    c.model_count = d->model_count;
    c.host_offsets.assign(d->offsets, d->offsets + d->offset_count);
    return 0;
}

int gkcu_blas_build(GkcuContext* ctx) {
    CTX_OR_FAIL(ctx);
    if (c.model_count == 0) { set_error("gkcu_blas_build: no scene uploaded"); return 1; }
    return build_all_blas(c);
}

int gkcu_blas_refit(GkcuContext* ctx, uint32_t model_index, const uint16_t* simple_vertices, uint32_t vertex_count) {
    CTX_OR_FAIL(ctx);
    if (!c.blas_built || model_index >= c.model_count) { set_error("gkcu_blas_refit: BLAS not built / bad model index"); return 1; }
    const GkModelData& md = c.host_offsets[model_index * 10];
    if (!simple_vertices || vertex_count != md.vertexCount) { set_error("gkcu_blas_refit: vertex_count must equal the model's vertex count"); return 1; }
    GK_CUDA(cudaMemcpyAsync(c.simple.as<uint16_t>() + (size_t)md.vertexOffset * 4, simple_vertices, (size_t)vertex_count * 8, cudaMemcpyHostToDevice, c.stream));
    return refit_blas(c, model_index);
}

int gkcu_update_nodes(GkcuContext* ctx, const GkNodeProxy* nodes, uint32_t count) {
    CTX_OR_FAIL(ctx);
    if (count && !nodes) { set_error("gkcu_update_nodes: null array"); return 1; }
    if (c.skip_unchanged_nodes) {  // gkcu_set_skip_unchanged_nodes: identical proxies keep the TLAS (and skip the upload)
        const size_t bytes = (size_t)count * sizeof(GkNodeProxy);
        if (count == c.node_count && !c.tlas_dirty && c.nodes_shadow.size() == bytes && (bytes == 0 || memcmp(c.nodes_shadow.data(), nodes, bytes) == 0)) return 0;
        c.nodes_shadow.assign(reinterpret_cast<const uint8_t*>(nodes), reinterpret_cast<const uint8_t*>(nodes) + bytes);
    }
    if (!c.nodes.reserve(std::max<size_t>(1, count) * sizeof(GkNodeProxy))) return 1;
    if (count) GK_CUDA(cudaMemcpyAsync(c.nodes.p, nodes, (size_t)count * sizeof(GkNodeProxy), cudaMemcpyHostToDevice, c.stream));
    c.node_count = count;
    // which instances make it into the TLAS (mask != 0, RayTraceBaseRenderer.cpp:178) is decided on the device by build_tlas
    c.tlas_dirty = true;
    GK_CUDA(cudaStreamSynchronize(c.stream));  // the caller may reuse its array
    return 0;
}

int gkcu_update_nodes_sparse(GkcuContext* ctx, const uint32_t* indices, const GkNodeProxy* proxies, uint32_t count) {
    CTX_OR_FAIL(ctx);
    if (count == 0) return 0;
    if (!indices || !proxies) { set_error("gkcu_update_nodes_sparse: null array"); return 1; }
    if (c.node_count == 0) { set_error("gkcu_update_nodes_sparse: call gkcu_update_nodes first"); return 1; }
    for (uint32_t k = 0; k < count; ++k)
        if (indices[k] >= c.node_count) { set_error("gkcu_update_nodes_sparse: index beyond the node count"); return 1; }
    if (!c.node_stage.reserve((size_t)count * sizeof(GkNodeProxy)) || !c.node_stage_idx.reserve((size_t)count * 4)) return 1;
    GK_CUDA(cudaMemcpyAsync(c.node_stage.p, proxies, (size_t)count * sizeof(GkNodeProxy), cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaMemcpyAsync(c.node_stage_idx.p, indices, (size_t)count * 4, cudaMemcpyHostToDevice, c.stream));
    if (scatter_nodes(c, count)) return 1;
    c.nodes_shadow.clear();  // the shadow copy no longer matches the device array
    c.tlas_dirty = true;
    GK_CUDA(cudaStreamSynchronize(c.stream));  // the caller may reuse its arrays
    return 0;
}

int gkcu_update_materials(GkcuContext* ctx, const GkMaterial* materials, uint32_t count) {
    CTX_OR_FAIL(ctx);
    if (count && !materials) { set_error("gkcu_update_materials: null array"); return 1; }
    if (count > GK_MAX_MATERIALS) { set_error("gkcu_update_materials: more than MAX_MATERIALS"); return 1; }
    if (!c.materials.reserve(std::max<size_t>(1, count) * sizeof(GkMaterial))) return 1;
    if (count) GK_CUDA(cudaMemcpyAsync(c.materials.p, materials, (size_t)count * sizeof(GkMaterial), cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    c.material_count = count;
    return 0;
}

int gkcu_tlas_build(GkcuContext* ctx) {
    CTX_OR_FAIL(ctx);
    if (!c.blas_built) { set_error("gkcu_tlas_build: call gkcu_blas_build first"); return 1; }
    if (c.skip_unchanged_nodes && !c.tlas_dirty) { c.timings.tlas_update_ms = 0; c.frame_built_tlas = false; return 0; }
    cudaEventRecord(c.kev[7], c.stream);
    int r = build_tlas(c);
    cudaEventRecord(c.ev[7], c.stream);
    if (r) return r;
    GK_CUDA(cudaStreamSynchronize(c.stream));
    cudaEventElapsedTime(&c.timings.tlas_update_ms, c.kev[7], c.ev[7]);
    c.frame_built_tlas = false;
    return 0;
}

int gkcu_set_sky(GkcuContext* ctx, uint32_t id, const float* rgba, uint32_t width, uint32_t height) {
    CTX_OR_FAIL(ctx);
    if (!rgba || width == 0 || height == 0 || id > 4096) { set_error("gkcu_set_sky: bad arguments"); return 1; }
    if (c.textures.size() <= id) c.textures.resize(id + 1);
    Context::Tex& t = c.textures[id];
    GK_CUDA(cudaStreamSynchronize(c.stream));
    if (t.texels) cudaFree(t.texels);
    t.texels = nullptr;
    GK_CUDA(cudaMalloc(&t.texels, (size_t)width * height * 16));
    GK_CUDA(cudaMemcpyAsync(t.texels, rgba, (size_t)width * height * 16, cudaMemcpyHostToDevice, c.stream));
    t.w = width;
    t.h = height;
    return upload_tex_table(c);
}

// host-side view of the packed device layout (gk_shading.cuh ProbeRec / ProbeMeta)
struct HostProbeRec { uint32_t cube[14], dist0, dist1; };
struct HostProbeMeta { uint32_t matId, age; };

int gkcu_set_probes(GkcuContext* ctx, const GkAmbientCube* cubes, const GkVoxelData* voxels) {
    CTX_OR_FAIL(ctx);
    if (!cubes || !voxels) { c.has_probes = false; return 0; }
    size_t n = (size_t)GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_Z;
    if (!c.probes.reserve(n * sizeof(HostProbeRec)) || !c.probe_meta.reserve(n * sizeof(HostProbeMeta))) return 1;
    std::vector<HostProbeRec> rec(n);
    std::vector<HostProbeMeta> meta(n);
    for (size_t i = 0; i < n; ++i) {
        memcpy(rec[i].cube, &cubes[i], sizeof(GkAmbientCube));
        rec[i].dist0 = voxels[i].distanceToSolid_gg_z01;
        rec[i].dist1 = voxels[i].distanceToSolid_x01_y01;
        meta[i].matId = voxels[i].matId;
        meta[i].age = voxels[i].age;
    }
    GK_CUDA(cudaMemcpyAsync(c.probes.p, rec.data(), n * sizeof(HostProbeRec), cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaMemcpyAsync(c.probe_meta.p, meta.data(), n * sizeof(HostProbeMeta), cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    c.has_probes = true;
    return 0;
}

int gkcu_set_lights(GkcuContext* ctx, const GkLightObject* lights, uint32_t count) {
    CTX_OR_FAIL(ctx);
    if (count && !lights) { set_error("gkcu_set_lights: null lights"); return 1; }
    if (count) {
        if (!c.lights.reserve((size_t)count * sizeof(GkLightObject))) return 1;
        GK_CUDA(cudaMemcpyAsync(c.lights.p, lights, (size_t)count * sizeof(GkLightObject), cudaMemcpyHostToDevice, c.stream));
        GK_CUDA(cudaStreamSynchronize(c.stream));
    }
    c.light_count = count;
    return 0;
}

int gkcu_set_sky_sh(GkcuContext* ctx, uint32_t id, const GkSphericalHarmonics* sh) {
    CTX_OR_FAIL(ctx);
    if (!sh || id >= 100) { set_error("gkcu_set_sky_sh: null coefficients or id >= 100 (Scene.cpp:38)"); return 1; }
    if (!c.sky_sh.p) {
        if (!c.sky_sh.reserve(100 * sizeof(GkSphericalHarmonics))) return 1;
        GK_CUDA(cudaMemsetAsync(c.sky_sh.p, 0, 100 * sizeof(GkSphericalHarmonics), c.stream));
    }
    GK_CUDA(cudaMemcpyAsync(c.sky_sh.as<GkSphericalHarmonics>() + id, sh, sizeof(GkSphericalHarmonics), cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    c.sky_sh_count = std::max(c.sky_sh_count, id + 1);
    return 0;
}

static int ensure_tlas(Context& c) {
    if (!c.blas_built) { set_error("call gkcu_scene_upload + gkcu_blas_build first"); return 1; }
    if (c.tlas_dirty) return build_tlas(c);
    return 0;
}

int gkcu_probes_voxelize(GkcuContext* ctx) {
    CTX_OR_FAIL(ctx);
    if (ensure_tlas(c)) return 1;
    return probes_voxelize(c);
}

int gkcu_probes_bake(GkcuContext* ctx, const GkUniformBufferObject* ubo, uint32_t first_cube, uint32_t cube_count) {
    CTX_OR_FAIL(ctx);
    if (!ubo) { set_error("gkcu_probes_bake: null ubo"); return 1; }
    if (ensure_tlas(c)) return 1;
    return probes_bake(c, *ubo, first_cube, cube_count);
}

int gkcu_probes_read(GkcuContext* ctx, GkAmbientCube* cubes, GkVoxelData* voxels, GkPageIndex* pages) {
    CTX_OR_FAIL(ctx);
    if (!c.has_probes) { set_error("gkcu_probes_read: no probe state"); return 1; }
    const size_t n = (size_t)GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_Z;
    GK_CUDA(cudaStreamSynchronize(c.stream));
    if (cubes || voxels) {
        std::vector<HostProbeRec> rec(n);
        std::vector<HostProbeMeta> meta(n);
        GK_CUDA(cudaMemcpy(rec.data(), c.probes.p, n * sizeof(HostProbeRec), cudaMemcpyDeviceToHost));
        GK_CUDA(cudaMemcpy(meta.data(), c.probe_meta.p, n * sizeof(HostProbeMeta), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) {
            if (cubes) memcpy(&cubes[i], rec[i].cube, sizeof(GkAmbientCube));
            if (voxels) {
                voxels[i].matId = meta[i].matId;
                voxels[i].age = meta[i].age;
                voxels[i].distanceToSolid_gg_z01 = rec[i].dist0;
                voxels[i].distanceToSolid_x01_y01 = rec[i].dist1;
            }
        }
    }
    if (pages) {
        if (!c.pages.p) { set_error("gkcu_probes_read: no page index (probes were set with gkcu_set_probes)"); return 1; }
        GK_CUDA(cudaMemcpy(pages, c.pages.p, (size_t)GK_PAGE_COUNT * GK_PAGE_COUNT * sizeof(GkPageIndex), cudaMemcpyDeviceToHost));
    }
    return 0;
}

static int frame_prologue(Context& c, const GkUniformBufferObject* ubo, const GkcuFrameParams* params) {
    if (!ubo || !params) { set_error("null ubo/params"); return 1; }
    if (c.rt.w == 0) { set_error("call gkcu_resize first"); return 1; }
    if (!c.blas_built) { set_error("call gkcu_scene_upload + gkcu_blas_build first"); return 1; }
    if (params->spp_rank >= std::max(1u, params->spp_world)) { set_error("spp_rank >= max(1, spp_world)"); return 1; }
    if (params->spp_world > GKCU_MAX_PEERS) { set_error("spp_world > 8"); return 1; }
    GK_CUDA(cudaMemcpyAsync(c.ubo.p, ubo, sizeof(GkUniformBufferObject), cudaMemcpyHostToDevice, c.stream));
    return 0;
}

// zero-copy present: tell the consumer of the bound external image that RT_DENOISED is complete
static int signal_present(Context& c) {
    if (!c.ext_sem) return 0;
    cudaExternalSemaphoreSignalParams sp{};
    sp.params.fence.value = ++c.ext_sem_value;  // ignored by binary semaphores (VK_SEMAPHORE_TYPE_BINARY)
    GK_CUDA(cudaSignalExternalSemaphoresAsync(&c.ext_sem, &sp, 1, c.stream));
    return 0;
}

static void end_frame_swaps(Context& c) {
    // copy pass (PathTracingRenderer.cpp:160-182) + object-id copy (VulkanBaseRenderer.cpp:1136-1152) as pointer swaps:
    // next frame's PREV_* / OBJECTID_1 are this frame's ACCUM_* / OBJECTID_0.
    c.rt.cur ^= 1;
}

int gkcu_render_frame(GkcuContext* ctx, const GkUniformBufferObject* ubo, const GkcuFrameParams* params) {
    CTX_OR_FAIL(ctx);
    const uint64_t launches0 = c.launches;
    if (frame_prologue(c, ubo, params)) return 1;
    cudaStream_t st = c.stream;
    cudaEventRecord(c.ev[0], st);
    c.frame_built_tlas = c.tlas_dirty;
    if (c.tlas_dirty && build_tlas(c)) return 1;  // RayTraceBaseRenderer::PreRender "TLAS Update"
    if (params->primary_visibility) {             // VulkanBaseRenderer::PreRender "gpu cull" + "visibility pass"
        if (run_gpu_cull(c, *ubo) || run_visibility_pass(c, *ubo)) return 1;
    }
    cudaEventRecord(c.ev[1], st);
    if (run_rt_pass(c, *ubo, *params)) return 1;
    cudaEventRecord(c.ev[2], st);
    if (!params->skip_post) {
        if (run_reproject(c, *params)) return 1;
        cudaEventRecord(c.ev[3], st);
        if (run_denoise(c, *ubo)) return 1;
        cudaEventRecord(c.ev[4], st);
        if (signal_present(c)) return 1;
        end_frame_swaps(c);
    } else {
        cudaEventRecord(c.ev[3], st);
        cudaEventRecord(c.ev[4], st);
    }
    cudaEventRecord(c.ev[5], st);
    c.counters.kernel_launches = c.launches - launches0;
    return 0;
}

int gkcu_radiance_buffer(GkcuContext* ctx, void** device_ptr, size_t* bytes) {
    CTX_OR_FAIL(ctx);
    if (!device_ptr || !bytes) { set_error("null out pointer"); return 1; }
    return radiance_buffer(c, device_ptr, bytes);
}

int gkcu_render_resolve(GkcuContext* ctx, const GkUniformBufferObject* ubo, const GkcuFrameParams* params) {
    CTX_OR_FAIL(ctx);
    const uint64_t launches0 = c.launches;
    if (frame_prologue(c, ubo, params)) return 1;
    cudaStream_t st = c.stream;
    cudaEventRecord(c.ev[2], st);
    if (run_resolve(c, *ubo, *params)) return 1;
    if (run_reproject(c, *params)) return 1;
    cudaEventRecord(c.ev[3], st);
    if (run_denoise(c, *ubo)) return 1;
    cudaEventRecord(c.ev[4], st);
    if (signal_present(c)) return 1;
    end_frame_swaps(c);
    cudaEventRecord(c.ev[5], st);
    c.counters.kernel_launches += c.launches - launches0;
    return 0;
}

static void drop_peers(Context& c) {
    for (void* p : c.ipc_opened) cudaIpcCloseMemHandle(p);
    c.ipc_opened.clear();
    memset(c.peer, 0, sizeof(c.peer));
    // The barrier epoch and the flag block survive: the peers keep counting, and a rank that restarted at 0 would see their old
    // flags as "arrived" and read sums that are not written yet. (A resize frees the buffers the peers imported, so every rank of
    // the group has to import again after any of them resized - see gkcu.h.)
}

int gkcu_peer_export(GkcuContext* ctx, uint32_t what, void* handle, void** device_ptr) {
    CTX_OR_FAIL(ctx);
    void* p = peer_local_buffer(c, what);
    if (!p) { set_error("gkcu_peer_export: unknown buffer or gkcu_resize not called"); return 1; }
    if (handle) {
        static_assert(sizeof(cudaIpcMemHandle_t) == GKCU_IPC_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t h;
        GK_CUDA(cudaIpcGetMemHandle(&h, p));
        memcpy(handle, &h, sizeof(h));
    }
    if (device_ptr) *device_ptr = p;
    return 0;
}

int gkcu_peer_import(GkcuContext* ctx, uint32_t rank, uint32_t what, const void* handle, void* device_ptr) {
    CTX_OR_FAIL(ctx);
    if (rank >= GKCU_MAX_PEERS || what >= GKCU_PEER_KINDS) { set_error("gkcu_peer_import: rank >= 8 or unknown buffer"); return 1; }
    if (!c.preloaded) {  // a context that imports peers is about to run multi-GPU frames
        path_preload();
        post_preload();
        c.preloaded = true;
    }
    void* p = device_ptr;
    if (handle) {
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, sizeof(h));
        GK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c.ipc_opened.push_back(p);
    }
    if (!p) { set_error("gkcu_peer_import: neither a handle nor a device pointer"); return 1; }
    c.peer[rank][what] = p;
    return 0;
}

int gkcu_render_resolve_strip(GkcuContext* ctx, const GkUniformBufferObject* ubo, const GkcuFrameParams* params) {
    CTX_OR_FAIL(ctx);
    if (frame_prologue(c, ubo, params)) return 1;
    const uint64_t launches0 = c.launches;
    if (run_resolve_strip(c, *ubo, *params)) return 1;
    c.counters.kernel_launches += c.launches - launches0;
    return 0;
}

int gkcu_render_resolve_post_strip(GkcuContext* ctx, const GkUniformBufferObject* ubo, const GkcuFrameParams* params) {
    CTX_OR_FAIL(ctx);
    if (frame_prologue(c, ubo, params)) return 1;
    const uint64_t launches0 = c.launches;
    cudaEventRecord(c.ev[2], c.stream);
    if (run_resolve_post_strip(c, *ubo, *params)) return 1;
    cudaEventRecord(c.ev[3], c.stream);
    cudaEventRecord(c.ev[4], c.stream);
    if (params->spp_rank == 0 && signal_present(c)) return 1;
    end_frame_swaps(c);
    cudaEventRecord(c.ev[5], c.stream);
    c.counters.kernel_launches += c.launches - launches0;
    return 0;
}

int gkcu_synchronize(GkcuContext* ctx) {
    CTX_OR_FAIL(ctx);
    GK_CUDA(cudaStreamSynchronize(c.stream));
    if (c.peer_flags.p && c.peer_epoch) {  // a device-side barrier that gave up means a peer never arrived
        unsigned long long timeouts = 0;
        GK_CUDA(cudaMemcpy(&timeouts, static_cast<unsigned long long*>(c.peer_flags.p) + 16, 8, cudaMemcpyDeviceToHost));
        if (timeouts) {
            cudaMemset(static_cast<unsigned long long*>(c.peer_flags.p) + 16, 0, 8);
            set_error("multi-GPU barrier timed out: a peer did not reach gkcu_render_resolve_strip");
            return 1;
        }
    }
    return 0;
}

int gkcu_readback(GkcuContext* ctx, uint32_t rt_index, void* dst, size_t dst_bytes) {
    CTX_OR_FAIL(ctx);
    RenderTargets& t = c.rt;
    if (t.w == 0) { set_error("gkcu_readback: no images"); return 1; }
    size_t n = (size_t)t.w * t.h;
    const void* src = nullptr;
    size_t bytes = 0;
    // after a frame the swap has happened: the newest ACCUM_* / OBJECTID_0 live in slot cur^1 (== PREV_* / OBJECTID_1,
    // exactly as in the reference after its copy pass)
    const int latest = t.cur ^ 1;
    switch (rt_index) {
        case GK_RT_ACCUMLATE_DIFFUSE: case GK_RT_SINGLE_PREV_DIFFUSE: src = t.accum_diffuse[latest]; bytes = n * 8; break;
        case GK_RT_ACCUMLATE_SPECULAR: case GK_RT_SINGLE_PREV_SPECULAR: src = t.accum_specular[latest]; bytes = n * 8; break;
        case GK_RT_ACCUMLATE_ALBEDO: case GK_RT_SINGLE_PREV_ALBEDO: src = t.accum_albedo[latest]; bytes = n * 8; break;
        case GK_RT_OBJEDCTID_0: case GK_RT_OBJEDCTID_1: src = t.objectid[latest]; bytes = n * 4; break;
        case GK_RT_SINGLE_DIFFUSE: src = t.single_diffuse; bytes = n * 8; break;
        case GK_RT_SINGLE_SPECULAR: src = t.single_specular; bytes = n * 8; break;
        case GK_RT_ALBEDO: src = t.albedo; bytes = n * 8; break;
        case GK_RT_NORMAL: src = t.normal; bytes = n * 8; break;
        case GK_RT_DENOISED: src = t.denoised_out(); bytes = n * 8; break;
        case GK_RT_MOTIONVECTOR: src = t.motion; bytes = n * 8; break;
        case GK_RT_PREV_DEPTHBUFFER: src = t.prev_depth; bytes = n * 4; break;
        case GK_RT_MOTIONMOMENT: src = t.moment; bytes = n * 2; break;
        case GK_RT_EXT_PRIMITIVEID: src = t.primitiveid; bytes = n * 4; break;
        case GK_RT_MINIGBUFFER: src = t.minigbuffer; bytes = n * 4; break;
        default: set_error("gkcu_readback: unknown storage image index"); return 1;
    }
    if (!dst || dst_bytes < bytes) { set_error("gkcu_readback: destination too small"); return 1; }
    GK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

int gkcu_readback_denoised_async(GkcuContext* ctx, void* dst, size_t dst_bytes) {
    CTX_OR_FAIL(ctx);
    RenderTargets& t = c.rt;
    const size_t bytes = (size_t)t.w * t.h * 8;
    if (t.w == 0) { set_error("gkcu_readback_denoised_async: no images"); return 1; }
    if (!dst || dst_bytes < bytes) { set_error("gkcu_readback_denoised_async: destination too small"); return 1; }
    if (!c.copy_stream) {
        GK_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
        GK_CUDA(cudaEventCreateWithFlags(&c.copy_ready, cudaEventDisableTiming));
        GK_CUDA(cudaEventCreateWithFlags(&c.copy_done, cudaEventDisableTiming));
    }
    GK_CUDA(cudaEventRecord(c.copy_ready, c.stream));           // the frame that produced the image
    GK_CUDA(cudaStreamWaitEvent(c.copy_stream, c.copy_ready, 0));
    GK_CUDA(cudaMemcpyAsync(dst, t.denoised_out(), bytes, cudaMemcpyDeviceToHost, c.copy_stream));
    GK_CUDA(cudaEventRecord(c.copy_done, c.copy_stream));       // run_denoise waits for this before it overwrites the image
    c.copy_pending = true;
    return 0;
}

int gkcu_readback_wait(GkcuContext* ctx) {
    CTX_OR_FAIL(ctx);
    if (c.copy_stream) GK_CUDA(cudaStreamSynchronize(c.copy_stream));
    return 0;
}

// Test/bring-up hook: overwrite a storage image from host memory (same formats as gkcu_readback); the "latest" slot.
int gkcu_upload_image(GkcuContext* ctx, uint32_t rt_index, const void* src, size_t src_bytes) {
    CTX_OR_FAIL(ctx);
    RenderTargets& t = c.rt;
    if (t.w == 0 || !src) { set_error("gkcu_upload_image: no images"); return 1; }
    size_t n = (size_t)t.w * t.h;
    void* dst = nullptr;
    size_t bytes = 0;
    const int prev = t.cur ^ 1, cur = t.cur;
    switch (rt_index) {
        case GK_RT_SINGLE_PREV_DIFFUSE: dst = t.accum_diffuse[prev]; bytes = n * 8; break;
        case GK_RT_SINGLE_PREV_SPECULAR: dst = t.accum_specular[prev]; bytes = n * 8; break;
        case GK_RT_SINGLE_PREV_ALBEDO: dst = t.accum_albedo[prev]; bytes = n * 8; break;
        case GK_RT_ACCUMLATE_DIFFUSE: dst = t.accum_diffuse[cur]; bytes = n * 8; break;
        case GK_RT_ACCUMLATE_SPECULAR: dst = t.accum_specular[cur]; bytes = n * 8; break;
        case GK_RT_ACCUMLATE_ALBEDO: dst = t.accum_albedo[cur]; bytes = n * 8; break;
        case GK_RT_OBJEDCTID_0: dst = t.objectid[cur]; bytes = n * 4; break;
        case GK_RT_OBJEDCTID_1: dst = t.objectid[prev]; bytes = n * 4; break;
        case GK_RT_SINGLE_DIFFUSE: dst = t.single_diffuse; bytes = n * 8; break;
        case GK_RT_SINGLE_SPECULAR: dst = t.single_specular; bytes = n * 8; break;
        case GK_RT_ALBEDO: dst = t.albedo; bytes = n * 8; break;
        case GK_RT_NORMAL: dst = t.normal; bytes = n * 8; break;
        case GK_RT_MOTIONVECTOR: dst = t.motion; bytes = n * 8; break;
        case GK_RT_MOTIONMOMENT: dst = t.moment; bytes = n * 2; break;
        case GK_RT_MINIGBUFFER: dst = t.minigbuffer; bytes = n * 4; break;
        case GK_RT_PREV_DEPTHBUFFER: dst = t.prev_depth; bytes = n * 4; break;
        default: set_error("gkcu_upload_image: unsupported storage image index"); return 1;
    }
    if (src_bytes != bytes) { set_error("gkcu_upload_image: size mismatch"); return 1; }
    GK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

// Test/bring-up hook: run only the post passes (reproject -> compose -> swaps) on the current images.
int gkcu_run_post(GkcuContext* ctx, const GkUniformBufferObject* ubo, const GkcuFrameParams* params) {
    CTX_OR_FAIL(ctx);
    if (!ubo || !params || c.rt.w == 0) { set_error("gkcu_run_post: bad arguments"); return 1; }
    GK_CUDA(cudaMemcpyAsync(c.ubo.p, ubo, sizeof(GkUniformBufferObject), cudaMemcpyHostToDevice, c.stream));
    cudaEventRecord(c.ev[2], c.stream);
    if (run_reproject(c, *params)) return 1;
    cudaEventRecord(c.ev[3], c.stream);
    if (run_denoise(c, *ubo)) return 1;
    cudaEventRecord(c.ev[4], c.stream);
    if (signal_present(c)) return 1;
    end_frame_swaps(c);
    cudaEventRecord(c.ev[5], c.stream);
    return 0;
}

int gkcu_get_timings(GkcuContext* ctx, GkcuTimings* out) {
    CTX_OR_FAIL(ctx);
    if (!out) { set_error("null out"); return 1; }
    GK_CUDA(cudaStreamSynchronize(c.stream));
    float ms = 0;
    if (c.frame_built_tlas && cudaEventElapsedTime(&ms, c.ev[0], c.ev[1]) == cudaSuccess) c.timings.tlas_update_ms = ms;
    if (cudaEventElapsedTime(&ms, c.ev[1], c.ev[2]) == cudaSuccess) c.timings.rt_pass_ms = ms;
    if (cudaEventElapsedTime(&ms, c.ev[2], c.ev[3]) == cudaSuccess) c.timings.reproject_pass_ms = ms;
    if (cudaEventElapsedTime(&ms, c.ev[3], c.ev[4]) == cudaSuccess) c.timings.compose_pass_ms = ms;
    if (cudaEventElapsedTime(&ms, c.ev[4], c.ev[5]) == cudaSuccess) c.timings.copy_pass_ms = ms;
    if (cudaEventElapsedTime(&ms, c.ev[0], c.ev[5]) == cudaSuccess) c.timings.frame_ms = ms;
    cudaGetLastError();
    *out = c.timings;
    return 0;
}

int gkcu_get_counters(GkcuContext* ctx, GkcuCounters* out) {
    CTX_OR_FAIL(ctx);
    if (!out) { set_error("null out"); return 1; }
    GK_CUDA(cudaStreamSynchronize(c.stream));
    fetch_ray_counters(c);
    *out = c.counters;
    return 0;
}

// Debug/validation hook (tests/test_gpu_bvh.py): raw device->host copies of the acceleration-structure pools.
// what: 0 TLAS nodes, 1 BLAS nodes, 2 BLAS triangles, 3 instances, 4/5 TLAS node boxes lo/hi, 6 BLAS table, 7 NodeProxy array,
// 8/9 TLAS / BLAS leaf tables (Node8), 10/11 BLAS node boxes lo/hi.
int gkcu_debug_read(GkcuContext* ctx, uint32_t what, size_t offset, void* dst, size_t bytes) {
    CTX_OR_FAIL(ctx);
    const DevBuf* b = nullptr;
    switch (what) {
        case 0: b = &c.tlas_nodes; break;
        case 1: b = &c.blas_nodes; break;
        case 2: b = &c.blas_tris; break;
        case 3: b = &c.instances; break;
        case 4: b = &c.tlas_node_lo; break;
        case 5: b = &c.tlas_node_hi; break;
        case 6: b = &c.blas_table; break;
        case 7: b = &c.nodes; break;
        case 8: b = &c.tlas_leaf; break;
        case 9: b = &c.blas_leaf; break;
        case 10: b = &c.blas_node_lo; break;
        case 11: b = &c.blas_node_hi; break;
        default: set_error("gkcu_debug_read: bad selector"); return 1;
    }
    if (!b->p || offset + bytes > b->cap) { set_error("gkcu_debug_read: out of range"); return 1; }
    GK_CUDA(cudaMemcpyAsync(dst, static_cast<const char*>(b->p) + offset, bytes, cudaMemcpyDeviceToHost, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}
// out[0] TLAS node count, [1] TLAS levels, [2] instance count, [3] model count, [4 + 4m ..] per model: node_base, node_count, tri_base, prim_count;
// behind the models: node format (4 | 8), TLAS leaf-table entries, then per model: leaf_base, leaf_count (Node8 leaf tables)
int gkcu_debug_info(GkcuContext* ctx, uint32_t* out, uint32_t capacity) {
    CTX_OR_FAIL(ctx);
    if (capacity < 6 + 6 * c.model_count) { set_error("gkcu_debug_info: capacity"); return 1; }
    if (c.tlas_dirty && c.blas_built && build_tlas(c)) return 1;
    out[0] = c.tlas.node_count;
    out[1] = c.tlas.level_start.empty() ? 0 : (uint32_t)c.tlas.level_start.size() - 1;
    out[2] = c.node_count;
    out[3] = c.model_count;
    for (uint32_t m = 0; m < c.model_count && m < c.blas.size(); ++m) {
        out[4 + 4 * m] = c.blas[m].node_base;
        out[5 + 4 * m] = c.blas[m].node_count;
        out[6 + 4 * m] = c.blas[m].tri_base;
        out[7 + 4 * m] = c.blas[m].prim_count;
    }
    uint32_t* ext = out + 4 + 4 * c.model_count;
    ext[0] = (uint32_t)c.node_format;
    ext[1] = c.tlas.leaf_count;
    for (uint32_t m = 0; m < c.model_count && m < c.blas.size(); ++m) {
        ext[2 + 2 * m] = c.blas[m].leaf_base;
        ext[3 + 2 * m] = c.blas[m].leaf_count;
    }
    return 0;
}

int gkcu_set_tlas_refit(GkcuContext* ctx, uint32_t enabled, uint32_t rebuild_every) {
    CTX_OR_FAIL(ctx);
    c.tlas_refit = enabled ? 1 : 0;
    c.tlas_rebuild_every = (int)std::min<uint32_t>(rebuild_every, 1u << 30);
    if (!enabled) c.tlas_topo_valid = false;
    return 0;
}
int gkcu_tlas_stats(GkcuContext* ctx, uint64_t* builds, uint64_t* refits) {
    CTX_OR_FAIL(ctx);
    if (builds) *builds = c.tlas_builds;
    if (refits) *refits = c.tlas_refits;
    return 0;
}
int gkcu_set_skip_unchanged_nodes(GkcuContext* ctx, uint32_t enabled) {
    CTX_OR_FAIL(ctx);
    c.skip_unchanged_nodes = enabled ? 1 : 0;
    c.nodes_shadow.clear();
    return 0;
}

int gkcu_set_fast_stencil(GkcuContext* ctx, uint32_t enabled) {
    CTX_OR_FAIL(ctx);
    c.fast_stencil = enabled ? 1 : 0;
    return 0;
}

int gkcu_set_kernel_timing(GkcuContext* ctx, uint32_t enabled) {
    CTX_OR_FAIL(ctx);
    c.kernel_timing = enabled ? 1 : 0;
    return 0;
}

int gkcu_get_kernel_stats(GkcuContext* ctx, GkcuKernelStats* out) {
    CTX_OR_FAIL(ctx);
    if (!out) { set_error("null out"); return 1; }
    GK_CUDA(cudaStreamSynchronize(c.stream));
    fetch_ray_counters(c);
    *out = c.kstats;
    return 0;
}

uint64_t gkcu_stream(GkcuContext* ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->c.stream : 0; }

int gkcu_trace_rays(GkcuContext* ctx, const GkcuRay* rays, uint32_t count, GkcuHit* hits, uint32_t any_hit) {
    CTX_OR_FAIL(ctx);
    if (!rays || !hits) { set_error("gkcu_trace_rays: null array"); return 1; }
    if (!c.blas_built) { set_error("gkcu_trace_rays: no acceleration structure"); return 1; }
    if (count == 0) return 0;
    if (c.tlas_dirty && build_tlas(c)) return 1;
    if (!c.trace_rays.reserve((size_t)count * sizeof(GkcuRay)) || !c.trace_hits.reserve((size_t)count * sizeof(GkcuHit))) return 1;
    GK_CUDA(cudaMemcpyAsync(c.trace_rays.p, rays, (size_t)count * sizeof(GkcuRay), cudaMemcpyHostToDevice, c.stream));
    if (run_trace_rays(c, count, any_hit != 0)) return 1;
    GK_CUDA(cudaMemcpyAsync(hits, c.trace_hits.p, (size_t)count * sizeof(GkcuHit), cudaMemcpyDeviceToHost, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------- zero-copy present (external memory / semaphore)
namespace {
// driver entry points of the virtual-memory API, fetched like cuTensorMapEncodeTiled (the library links cudart only)
template <class Fn>
Fn driver_fn(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    return reinterpret_cast<Fn>(p);
}
int unbind_external(Context& c) {
    cudaStreamSynchronize(c.stream);
    if (c.copy_stream) cudaStreamSynchronize(c.copy_stream);
    c.rt.denoised_ext = nullptr;
    if (c.ext_mem) { cudaDestroyExternalMemory(c.ext_mem); c.ext_mem = nullptr; }
    if (c.ext_vmm_va) {
        typedef int (*UnmapFn)(unsigned long long, size_t);
        typedef int (*FreeFn)(unsigned long long, size_t);
        typedef int (*ReleaseFn)(unsigned long long);
        auto unmap = driver_fn<UnmapFn>("cuMemUnmap");
        auto afree = driver_fn<FreeFn>("cuMemAddressFree");
        auto release = driver_fn<ReleaseFn>("cuMemRelease");
        if (unmap) unmap(c.ext_vmm_va, c.ext_vmm_size);
        if (afree) afree(c.ext_vmm_va, c.ext_vmm_size);
        if (release) release(c.ext_vmm_handle);
        c.ext_vmm_va = c.ext_vmm_handle = 0;
        c.ext_vmm_size = 0;
    }
    return 0;
}
}  // namespace

extern "C" int gkcu_unbind_external_image(GkcuContext* ctx, uint32_t rt_index) {
    CTX_OR_FAIL(ctx);
    if (rt_index != GK_RT_DENOISED) { set_error("gkcu_unbind_external_image: only RT_DENOISED can be bound"); return 1; }
    return unbind_external(c);
}

extern "C" int gkcu_bind_external_image(GkcuContext* ctx, uint32_t rt_index, uint32_t handle_type, int fd, size_t size_bytes, size_t offset_bytes,
                                        uint32_t row_pitch_bytes) {
    CTX_OR_FAIL(ctx);
    if (rt_index != GK_RT_DENOISED) { set_error("gkcu_bind_external_image: only RT_DENOISED (the image the engine presents) can be bound"); return 1; }
    if (c.rt.w == 0) { set_error("gkcu_bind_external_image: call gkcu_resize first"); return 1; }
    const size_t need = (size_t)c.rt.w * c.rt.h * 8;
    if (row_pitch_bytes != 0 && row_pitch_bytes != c.rt.w * 8) { set_error("gkcu_bind_external_image: only tightly packed linear images (row pitch = width * 8) are supported"); return 1; }
    if (fd < 0 || offset_bytes % 16 != 0 || offset_bytes + need > size_bytes) { set_error("gkcu_bind_external_image: bad descriptor, unaligned offset or allocation too small for the image"); return 1; }
    unbind_external(c);
    if (handle_type == GKCU_EXT_OPAQUE_FD) {
        cudaExternalMemoryHandleDesc hd{};
        hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
        hd.handle.fd = fd;  // ownership passes to CUDA on success (like vkGetMemoryFdKHR's contract on the importing side)
        hd.size = size_bytes;
        GK_CUDA(cudaImportExternalMemory(&c.ext_mem, &hd));
        cudaExternalMemoryBufferDesc bd{};
        bd.offset = offset_bytes;
        bd.size = need;
        void* p = nullptr;
        if (!check(cudaExternalMemoryGetMappedBuffer(&p, c.ext_mem, &bd), "cudaExternalMemoryGetMappedBuffer")) { unbind_external(c); return 1; }
        c.rt.denoised_ext = static_cast<uint2*>(p);
    } else if (handle_type == GKCU_EXT_CUDA_VMM_FD) {
        struct AccessDesc { int loc_type; int loc_id; int flags; };  // CUmemAccessDesc {CUmemLocation{type, id}, flags}
        typedef int (*ImportFn)(unsigned long long*, void*, int);
        typedef int (*ReserveFn)(unsigned long long*, size_t, size_t, unsigned long long, unsigned long long);
        typedef int (*MapFn)(unsigned long long, size_t, size_t, unsigned long long, unsigned long long);
        typedef int (*AccessFn)(unsigned long long, size_t, const AccessDesc*, size_t);
        auto import = driver_fn<ImportFn>("cuMemImportFromShareableHandle");
        auto reserve = driver_fn<ReserveFn>("cuMemAddressReserve");
        auto map = driver_fn<MapFn>("cuMemMap");
        auto access = driver_fn<AccessFn>("cuMemSetAccess");
        if (!import || !reserve || !map || !access) { set_error("gkcu_bind_external_image: the CUDA virtual-memory API is unavailable"); return 1; }
        unsigned long long h = 0, va = 0;
        if (import(&h, (void*)(uintptr_t)fd, 1 /* CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR */) != 0) { set_error("cuMemImportFromShareableHandle failed"); return 1; }
        c.ext_vmm_handle = h;
        c.ext_vmm_size = size_bytes;
        if (reserve(&va, size_bytes, 0, 0, 0) != 0) { set_error("cuMemAddressReserve failed"); unbind_external(c); return 1; }
        c.ext_vmm_va = va;
        AccessDesc ad{1 /* CU_MEM_LOCATION_TYPE_DEVICE */, c.device, 3 /* CU_MEM_ACCESS_FLAGS_PROT_READWRITE */};
        if (map(va, size_bytes, 0, h, 0) != 0 || access(va, size_bytes, &ad, 1) != 0) { set_error("cuMemMap / cuMemSetAccess failed"); unbind_external(c); return 1; }
        c.rt.denoised_ext = reinterpret_cast<uint2*>(va + offset_bytes);
    } else {
        set_error("gkcu_bind_external_image: unknown handle type");
        return 1;
    }
    return 0;
}

extern "C" int gkcu_bind_external_semaphore(GkcuContext* ctx, uint32_t handle_type, int fd) {
    CTX_OR_FAIL(ctx);
    GK_CUDA(cudaStreamSynchronize(c.stream));
    if (c.ext_sem) { cudaDestroyExternalSemaphore(c.ext_sem); c.ext_sem = nullptr; }
    if (fd < 0) return 0;  // unbind
    if (handle_type != GKCU_EXT_OPAQUE_FD) { set_error("gkcu_bind_external_semaphore: only OPAQUE_FD semaphores"); return 1; }
    cudaExternalSemaphoreHandleDesc sd{};
    sd.type = cudaExternalSemaphoreHandleTypeOpaqueFd;
    sd.handle.fd = fd;
    GK_CUDA(cudaImportExternalSemaphore(&c.ext_sem, &sd));
    c.ext_sem_value = 0;
    return 0;
}

// ---------------------------------------------------------------------- visibility front end (gk_raster.cu)
extern "C" int gkcu_scene_upload_raster(GkcuContext* ctx, const uint32_t* raster_indices, uint32_t index_count, const uint32_t* reorders, uint32_t reorder_count) {
    CTX_OR_FAIL(ctx);
    if (!raster_indices || !reorders) { set_error("gkcu_scene_upload_raster: null array"); return 1; }
    if (c.model_count == 0) { set_error("gkcu_scene_upload_raster: call gkcu_scene_upload first"); return 1; }
    if (index_count != c.index_count) { set_error("gkcu_scene_upload_raster: the raster index buffer must have as many indices as GPUScene.Indices"); return 1; }
    for (uint32_t i = 0; i < c.offset_count; ++i) {  // every index of a section must land inside Reorders, every Reorders entry inside the vertex buffer
        const GkModelData& md = c.host_offsets[i];
        for (uint32_t k = 0; k < md.indexCount; ++k)
            if ((uint64_t)raster_indices[md.indexOffset + k] + md.reorderOffset >= reorder_count) { set_error("gkcu_scene_upload_raster: index beyond Reorders"); return 1; }
    }
    for (uint32_t i = 0; i < reorder_count; ++i)
        if (reorders[i] >= c.vertex_count) { set_error("gkcu_scene_upload_raster: Reorders entry beyond the vertex buffer"); return 1; }
    if (!c.raster_indices.reserve((size_t)std::max(1u, index_count) * 4) || !c.reorders.reserve((size_t)std::max(1u, reorder_count) * 4)) return 1;
    GK_CUDA(cudaMemcpyAsync(c.raster_indices.p, raster_indices, (size_t)index_count * 4, cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaMemcpyAsync(c.reorders.p, reorders, (size_t)reorder_count * 4, cudaMemcpyHostToDevice, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    c.raster_index_count = index_count;
    c.reorder_count = reorder_count;
    return 0;
}

extern "C" int gkcu_cull(GkcuContext* ctx, const GkUniformBufferObject* ubo) {
    CTX_OR_FAIL(ctx);
    if (!ubo || c.rt.w == 0 || c.model_count == 0) { set_error("gkcu_cull: null ubo, or no images / scene yet"); return 1; }
    return run_gpu_cull(c, *ubo);
}

extern "C" int gkcu_visibility_pass(GkcuContext* ctx, const GkUniformBufferObject* ubo) {
    CTX_OR_FAIL(ctx);
    if (!ubo || c.rt.w == 0 || c.model_count == 0) { set_error("gkcu_visibility_pass: null ubo, or no images / scene yet"); return 1; }
    if (c.raster_index_count == 0) { set_error("gkcu_visibility_pass: call gkcu_scene_upload_raster first"); return 1; }
    return run_visibility_pass(c, *ubo);
}

extern "C" int gkcu_get_gpu_driven_stats(GkcuContext* ctx, GkGPUDrivenStat* out) {
    CTX_OR_FAIL(ctx);
    if (!out || !c.gpu_stats.p) { set_error("gkcu_get_gpu_driven_stats: null out, or gkcu_cull has not run"); return 1; }
    GK_CUDA(cudaMemcpyAsync(out, c.gpu_stats.p, sizeof(GkGPUDrivenStat), cudaMemcpyDeviceToHost, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

extern "C" int gkcu_read_draw_commands(GkcuContext* ctx, void* dst, uint32_t count) {
    CTX_OR_FAIL(ctx);
    if (!dst || !c.draw_cmds.p || count > c.draw_count || c.draw_count == 0xFFFFFFFFu) { set_error("gkcu_read_draw_commands: null dst, gkcu_cull has not run, or count beyond the batches"); return 1; }
    GK_CUDA(cudaMemcpyAsync(dst, c.draw_cmds.p, (size_t)count * 20, cudaMemcpyDeviceToHost, c.stream));
    GK_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}
