// gk_context.h — the renderer context behind the opaque GkcuContext handle, shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/gkcu.h"
#include "gk_bvh.cuh"

namespace gkcu {

void set_error(const std::string& msg);
bool check(cudaError_t e, const char* what);
#define GK_CUDA(expr)                                \
    do {                                             \
        if (!gkcu::check((expr), #expr)) return 1;   \
    } while (0)

// Growable device allocation (cudaMalloc is only hit when a buffer has to grow).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool reserve(size_t bytes);
    void release();
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

// One built wide BVH (BLAS of one model, or the TLAS).
struct WideBvhInfo {
    uint32_t node_base = 0;     // first node in the owning node pool
    uint32_t node_count = 0;
    uint32_t tri_base = 0;      // BLAS only: first leaf-ordered triangle in the triangle pool
    uint32_t prim_count = 0;
    uint32_t vertex_offset = 0, index_offset = 0;
    uint32_t leaf_base = 0, leaf_count = 0;  // Node8: this BVH's slice of the leaf table (the first GK_LEAF_SLACK entries are slack)
    std::vector<uint32_t> level_start;  // node index (relative) of each BFS level, size levels+1
    float root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
};

// Scratch + outputs of one BVH2 build, see gk_bvh_build.cu.
struct BuildScratch {
    DevBuf prim_lo, prim_hi, keys, keys_sorted, order, order_tmp, sort_tmp;
    DevBuf n2_lo, n2_hi, n2_left, n2_right, n2_parent, n2_first, n2_count, n2_flag;
    DevBuf queue_a, queue_b, counters, bounds;
    DevBuf bins, pos_node, pos_node_tmp, active, cursors;
};

struct RenderTargets {
    uint32_t w = 0, h = 0;
    // fp16x4 images as uint2 per pixel
    uint2 *accum_diffuse[2] = {nullptr, nullptr}, *accum_specular[2] = {nullptr, nullptr}, *accum_albedo[2] = {nullptr, nullptr};
    uint2 *single_diffuse = nullptr, *single_specular = nullptr, *albedo = nullptr, *normal = nullptr, *denoised = nullptr;
    uint32_t* objectid[2] = {nullptr, nullptr};
    uint32_t* primitiveid = nullptr;
    float2* motion = nullptr;
    float* prev_depth = nullptr;
    uint16_t* moment = nullptr;
    uint32_t* minigbuffer = nullptr;  // RT_MINIGBUFFER: the visibility buffer (R32_UINT), written by the visibility pass
    int cur = 0;  // accum_*[cur] is this frame's ACCUM, [cur^1] is PREV; objectid[cur] is OBJECTID_0
    uint2* denoised_ext = nullptr;  // RT_DENOISED bound to imported external memory (gkcu_bind_external_image): the compose pass writes here
    uint2* denoised_out() const { return denoised_ext ? denoised_ext : denoised; }
};

struct PathBuffers;  // gk_pathtrace.cu

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    // Small device -> host results (queue sizes, ray counters, TLAS level table) are STORED by kernels into mapped pinned memory instead of
    // copied with cudaMemcpyAsync: a 4-byte copy queues on the D2H copy engine behind the 16.6 MB asynchronous present of the previous
    // frame and stalls the stream for as long as that copy takes (8 ms on a box with a slow link). 64 words for the builders, then 64 per lane.
    uint32_t* pub_host = nullptr;
    uint32_t* pub_dev = nullptr;
    const char* last_api = nullptr;  // name of the previous C-ABI entry point (diagnostics)

    // scene (GPUScene buffers, BasicTypes.slang:261-291)
    DevBuf vertices, simple, indices, offsets, materials, nodes;
    uint32_t vertex_count = 0, index_count = 0, offset_count = 0, model_count = 0, material_count = 0, node_count = 0;
    std::vector<GkModelData> host_offsets;

    // acceleration structures
    DevBuf blas_nodes, blas_node_lo, blas_node_hi, blas_tris, blas_tri_prim;  // pools shared by all BLAS
    std::vector<WideBvhInfo> blas;
    DevBuf blas_table;  // per model: node_base, tri_base, prim_count, shade-triangle base + root box (2 x float4)
    DevBuf shade_tris;  // 32 B per primitive (primitive order): 3 normals + 3 uvs as fp16 + material slot, see gk_pathtrace.cu
    DevBuf shade_inst;  // 32 B per instance: shade-triangle base + the first 7 material ids of the node
    DevBuf tlas_nodes, tlas_node_lo, tlas_node_hi, instances, inst_lo, inst_hi, tlas_active_ids;
    DevBuf blas_leaf, tlas_leaf;  // Node8 leaf tables (triangle ranges / instance indices)
    // visibility front end (gk_raster.cu): GPUScene.Indices of the raster pass + GPUScene.Reorders, IndirectDrawCommands, GPUDrivenStats
    DevBuf raster_indices, reorders, draw_cmds, gpu_stats, vis_keys, raster_large, raster_counters;
    uint32_t raster_index_count = 0, reorder_count = 0, draw_count = 0xFFFFFFFFu;
    DevBuf inst_flags, select_tmp, select_count;  // device-side compaction of the instances that enter the TLAS
    DevBuf node_stage, node_stage_idx;            // gkcu_update_nodes_sparse staging
    WideBvhInfo tlas;
    bool tlas_dirty = true;
    // Per-frame TLAS update: refit the topology of the last full build while the same instances are active, rebuild when the set changes, when
    // the refitted tree has degraded (node areas + 30 %) or every tlas_rebuild_every frames (0 = never by count). gkcu_set_tlas_refit /
    // GKCU_TLAS_REFIT=0 rebuilds every frame like the reference's vkCmdBuildAccelerationStructures call (RayTraceBaseRenderer.cpp:205-218).
    int tlas_refit = 1;
    int tlas_rebuild_every = 0;
    bool tlas_topo_valid = false;
    uint32_t tlas_topo_n = 0, tlas_frames_since_build = 0;
    float tlas_built_cost = -1.0f;
    uint64_t tlas_builds = 0, tlas_refits = 0;
    DevBuf inst_flags_prev;
    int skip_unchanged_nodes = 0;        // gkcu_set_skip_unchanged_nodes: gkcu_update_nodes keeps the TLAS when the proxies are byte-identical
    std::vector<uint8_t> nodes_shadow;   // host copy of the last full NodeProxy[] upload (only kept while the option is on)
    bool frame_built_tlas = false;
    bool blas_built = false;
    BuildScratch scratch;
    int trace_variant = 5;  // GKCU_TRACE_VARIANT: 5 vote-selected step kind, 6 blocks/SM (default); 0 one step of any kind per iteration; 2 warp-granular batches, no refill
    int fast_stencil = 0;   // GKCU_FAST_STENCIL / gkcu_set_fast_stencil: approximate transcendentals + reciprocal in the compose pass (tolerance build)
    int use_tma = 1;        // GKCU_TMA=0 falls back to plain-load staging of the stencil tiles (2 / 3: only reproject / only denoise)
    int shade_minb = 7;     // GKCU_SHADE_MINB: 8 compiles k_shade for 8 blocks/SM (<= 64 registers)
    int trace_carveout = -1;  // k_trace shared-memory carve-out in percent (GKCU_TRACE_CARVEOUT), -1 = sized to the resident blocks
    int trace_rpb = 128;      // GKCU_TRACE_RPB: rays per block the k_trace grid is sized for
    int trace_grid_mult = 16; // GKCU_TRACE_GRID: k_trace grid cap in blocks per SM
    int match_carveout = 0;   // GKCU_MATCH_CARVEOUT=1: k_primary / k_shade / k_tail ask for k_trace's shared-memory split (see run_rt_pass)
    int carveout_applied = 0; // trace_carveout + 2 once the function attribute has been set on this context's device
    int node_format = 4;      // GKCU_NODE_FORMAT: 4 = fp32 4-wide Node4 (default: 4.5 % faster on the bench frame, profiles/r02_experiments.md), 8 = quantised 8-wide Node8 (gk_bvh.cuh)
    int sort_rays = 0;      // GKCU_SORT_RAYS = bits per axis of the origin grid the rays of a wave are sorted on before k_trace (0 = queue order)
    float sort_extent = 0;  // GKCU_SORT_EXTENT: side of the origin grid's cube around the world origin (0 = the TLAS root box)
    int refill_below = 16;  // k_trace: refill idle lanes when fewer than this many lanes are busy (GKCU_REFILL)
    int blas_builder = 1;  // 0 = LBVH, 1 = SAH-binned (GKCU_BLAS_BUILDER=lbvh|sah)

    // textures / probes
    struct Tex { float4* texels = nullptr; uint32_t w = 0, h = 0; };
    std::vector<Tex> textures;
    DevBuf tex_table;
    uint32_t tex_count = 0;
    DevBuf probes, probe_meta;  // packed probe grid: 64-byte records + (material, age), see gk_shading.cuh
    bool has_probes = false;
    DevBuf pages, cube_snapshot, lights, sky_sh;  // probe bake: page index, snapshot of the range being baked, GPUScene.Lights / HDRSHs
    uint32_t light_count = 0, sky_sh_count = 0;

    DevBuf ubo;
    RenderTargets rt;
    PathBuffers* path = nullptr;

    GkcuTimings timings{};
    GkcuCounters counters{};
    GkcuKernelStats kstats{};
    cudaEvent_t kev[8] = {};
    cudaEvent_t wev[4][24] = {};  // per lane: trace/shade brackets of the last 8 waves (ring)
    cudaEvent_t pev[4][8] = {};   // per lane: "queue size of wave k has reached the host" (ring, no timing)
    int wave_seen[4] = {}, wave_launched[4] = {};  // waves of the last frame whose events the host has read / that were launched
    int kernel_timing = 0;        // per-wave event brackets for GkcuKernelStats.trace_ms / shade_ms (~4 us each, 0.45 ms per bench frame): off by default, gkcu_set_kernel_timing(ctx, 1) / GKCU_KERNEL_TIMING=1 turn them on
    // The frame's pixels can be dealt to `lanes` independent wavefronts on their own streams (checkerboard of 8 x 16 tiles), so one
    // lane's shade kernel and the thin end of its trace kernel overlap the other lanes' kernels (GKCU_LANES, 1..4). Measured: 2 lanes
    // -0.7 %, 3 lanes +2 %, 4 lanes +5 % (the kernels already fill the GPU; overlap only shares it), so the default is one lane.
    int lanes = 1;
    cudaStream_t lane_stream[4] = {};
    cudaEvent_t lane_ev[4][4] = {};  // per lane: tail begin / end, done; [0][3]: primary done
    int tail_cap = 448;        // GKCU_TAIL_CAP: most paths per SM the wavefront hands to k_tail (k_tail: 127 registers, 512 resident threads per SM)
    int tail_sort = 1;         // GKCU_TAIL_SORT: k_tail takes its slots ordered by the samples their pixel still has to run, longest first and dealt round-robin to the warps (counting sort, three small kernels): tail 2.65 -> 2.3 ms; 0 = queue order
    int tail_solo = 0;         // GKCU_TAIL_SOLO: the K longest paths of the ordered tail get a warp each
    int tail_spread = 32;      // GKCU_TAIL_SPREAD: paths per warp of the ordered tail
    int tail_fused = 0;        // GKCU_TAIL_FUSED=1: the tail runs as persistent vote-scheduled warps (k_tail_fused) instead of one thread per path
    int tail_segments = 0;     // GKCU_TAIL_SEGMENTS=a,b,c: the tail runs as segments of at most a, b, c rays per path with a compaction between them, then to the end (k_tail_seg); 0 = one k_tail launch
    int tail_segment_rays[14] = {};
    int tail_div = 32;         // the wavefront hands over to k_tail once the wave is narrower than pixels / tail_div (GKCU_TAIL_DIV); 32 since the loop has no host waits: rt pass 28.2 -> 27.3 ms against 64, profiles/r02_experiments.md section 10
    int wave_chunk = 6;        // how many waves the launches may run ahead of the queue sizes the host has read, 1..7 (GKCU_WAVE_CHUNK)
    cudaEvent_t ev[8] = {};
    uint64_t launches = 0;
    DevBuf trace_rays, trace_hits;
    cudaStream_t copy_stream = nullptr;  // gkcu_readback_denoised_async
    cudaEvent_t copy_ready = nullptr, copy_done = nullptr;
    bool copy_pending = false;

    // multi-GPU peer resolve: peer[rank][what] device pointers (local, same-process, or opened from an IPC handle)
    void* peer[GKCU_MAX_PEERS][GKCU_PEER_KINDS] = {};
    bool preloaded = false;  // the kernels of the multi-GPU frame tail have been resolved (path_preload / post_preload)
    DevBuf peer_flags;          // this rank's flag block: ready[8], done[8], timeouts (uint64 each), written by the peers
    uint64_t peer_epoch = 0;
    std::vector<void*> ipc_opened;

    // zero-copy present: imported memory behind RT_DENOISED and the semaphore signalled after the compose pass
    cudaExternalMemory_t ext_mem = nullptr;      // GKCU_EXT_OPAQUE_FD
    unsigned long long ext_vmm_handle = 0;       // GKCU_EXT_CUDA_VMM_FD: CUmemGenericAllocationHandle
    unsigned long long ext_vmm_va = 0;           // ... and the address range it is mapped at
    size_t ext_vmm_size = 0;
    cudaExternalSemaphore_t ext_sem = nullptr;
    uint64_t ext_sem_value = 0;
};

gkd::TraceScene make_trace_scene(const Context& c);

// gk_bvh_build.cu
int build_all_blas(Context& c);
int refit_blas(Context& c, uint32_t model);
int build_tlas(Context& c);
int scatter_nodes(Context& c, uint32_t count);

// gk_pathtrace.cu
int path_resize(Context& c);
void path_release(Context& c);
int run_rt_pass(Context& c, const GkUniformBufferObject& ubo, const GkcuFrameParams& p);
int run_resolve(Context& c, const GkUniformBufferObject& ubo, const GkcuFrameParams& p);
int run_trace_rays(Context& c, uint32_t count, bool any_hit);
void fetch_ray_counters(Context& c);
int radiance_buffer(Context& c, void** ptr, size_t* bytes);
int run_resolve_strip(Context& c, const GkUniformBufferObject& ubo, const GkcuFrameParams& p);
int run_resolve_post_strip(Context& c, const GkUniformBufferObject& ubo, const GkcuFrameParams& p);
void path_preload();
void post_preload();
void* peer_local_buffer(Context& c, uint32_t what);

// gk_raster.cu
int run_gpu_cull(Context& c, const GkUniformBufferObject& ubo);
int run_visibility_pass(Context& c, const GkUniformBufferObject& ubo);

// gk_api.cu: dst_word..dst_word+words of the mapped block <- src (device), stream-ordered (a one-block kernel, no copy engine)
int publish_words(Context& c, const void* src, uint32_t dst_word, uint32_t words, cudaStream_t st);

// gk_probes.cu
int probes_voxelize(Context& c);
int probes_bake(Context& c, const GkUniformBufferObject& ubo, uint32_t first, uint32_t count);

// gk_post.cu
// Strip-parallel post of a multi-GPU frame (gkcu_render_resolve_post_strip): the reproject pass runs on rows [y_begin, y_end) and mirrors the
// rows [own_y0, own_y1) it owns into the other ranks' ACCUM_* images, the compose pass runs on the owned rows and writes `denoised`.
struct PostStrip {
    int y_begin = 0, y_end = 0, own_y0 = 0, own_y1 = 0, n_mirror = 0;
    uint2 *mirror_diffuse[GKCU_MAX_PEERS - 1] = {}, *mirror_specular[GKCU_MAX_PEERS - 1] = {}, *mirror_albedo[GKCU_MAX_PEERS - 1] = {};
    uint2* denoised = nullptr;  // where the compose pass stores (rank 0's RT_DENOISED over peer memory); nullptr = the local image
};
int run_reproject(Context& c, const GkcuFrameParams& p, const PostStrip* strip = nullptr);
int run_denoise(Context& c, const GkUniformBufferObject& ubo, const PostStrip* strip = nullptr);

}  // namespace gkcu
