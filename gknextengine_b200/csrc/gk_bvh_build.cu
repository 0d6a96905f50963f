// gk_bvh_build.cu — acceleration-structure builds on the GPU.
//
// Replaces the driver work behind vkCmdBuildAccelerationStructuresKHR in the reference:
//   BLAS per model, built at scene load   src/Vulkan/RayTracing/BottomLevelAccelerationStructure.cpp:42-59,
//                                         src/Rendering/RayTraceBaseRenderer.cpp:276-340
//   TLAS over the NodeProxy array, fully rebuilt when nodes change
//                                         src/Vulkan/RayTracing/TopLevelAccelerationStructure.cpp:50-81,
//                                         src/Rendering/RayTraceBaseRenderer.cpp:165-218
// plus a BLAS refit the reference does not have.
//
// Pipeline: primitive boxes -> binary BVH (Morton LBVH, Karras 2012; or level-synchronous SAH-binned build for BLAS)
//           -> collapse to 4-wide nodes with fp32 child boxes (Node4, gk_bvh.cuh) in BFS order
//           -> leaf-ordered triangle pool. The BFS level table makes the bottom-up refit a handful of launches.
#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <cooperative_groups.h>

#include "gk_context.h"

using namespace gkd;

namespace gkcu {

namespace {

static int kBlasLeafMax = 2;  // GKCU_LEAF_MAX (1..8); 2 measured best on the sphere/brick meshes (profiles/r01_experiments.md)
constexpr int kSahBins = 16;

__device__ __forceinline__ int f2o(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float o2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// ------------------------------------------------------------------------------------------- primitive setup
// Triangle boxes of one model from the BLAS inputs the reference hands to Vulkan: fp16x4 positions (stride 8 B) and
// uint32 model-local indices (src/Vulkan/RayTracing/BottomLevelGeometry.cpp:9-38).
__global__ void k_tri_boxes(const uint16_t* __restrict__ simple, const uint32_t* __restrict__ indices, uint32_t vertex_offset,
                            uint32_t index_offset, uint32_t n, float4* __restrict__ lo, float4* __restrict__ hi) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float3 mn = f3(FLT_MAX, FLT_MAX, FLT_MAX), mx = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int k = 0; k < 3; ++k) {
        uint32_t vi = vertex_offset + indices[index_offset + p * 3 + k];
        float3 v = f3(h2f(simple[vi * 4]), h2f(simple[vi * 4 + 1]), h2f(simple[vi * 4 + 2]));
        mn = f3(fminf(mn.x, v.x), fminf(mn.y, v.y), fminf(mn.z, v.z));
        mx = f3(fmaxf(mx.x, v.x), fmaxf(mx.y, v.y), fmaxf(mx.z, v.z));
    }
    lo[p] = f4(mn.x, mn.y, mn.z, 0);
    hi[p] = f4(mx.x, mx.y, mx.z, 0);
}

__global__ void k_init_bounds(int* b) {
    if (threadIdx.x < 3) b[threadIdx.x] = 0x7FFFFFFF;
    else if (threadIdx.x < 6) b[threadIdx.x] = (int)0x80000000;
}

__global__ void k_bounds(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, int* __restrict__ bounds) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 mn = f3(FLT_MAX, FLT_MAX, FLT_MAX), mx = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    if (i < n) {
        float4 a = lo[i], b = hi[i];
        if (a.x <= b.x) {  // inactive instances carry an inverted box
            mn = f3(a.x, a.y, a.z);
            mx = f3(b.x, b.y, b.z);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn.x = fminf(mn.x, __shfl_xor_sync(0xffffffffu, mn.x, o)); mn.y = fminf(mn.y, __shfl_xor_sync(0xffffffffu, mn.y, o));
        mn.z = fminf(mn.z, __shfl_xor_sync(0xffffffffu, mn.z, o)); mx.x = fmaxf(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, o));
        mx.y = fmaxf(mx.y, __shfl_xor_sync(0xffffffffu, mx.y, o)); mx.z = fmaxf(mx.z, __shfl_xor_sync(0xffffffffu, mx.z, o));
    }
    if ((threadIdx.x & 31) == 0 && mn.x <= mx.x) {
        atomicMin(&bounds[0], f2o(mn.x)); atomicMin(&bounds[1], f2o(mn.y)); atomicMin(&bounds[2], f2o(mn.z));
        atomicMax(&bounds[3], f2o(mx.x)); atomicMax(&bounds[4], f2o(mx.y)); atomicMax(&bounds[5], f2o(mx.z));
    }
}

__device__ __forceinline__ uint64_t expand21(uint32_t v) {
    uint64_t x = v & 0x1FFFFFu;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, const int* __restrict__ bounds,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ order) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 bmn = f3(o2f(bounds[0]), o2f(bounds[1]), o2f(bounds[2])), bmx = f3(o2f(bounds[3]), o2f(bounds[4]), o2f(bounds[5]));
    float4 a = lo[i], b = hi[i];
    uint64_t key;
    if (a.x > b.x) {
        key = 0x7FFFFFFFFFFFFFFFull;  // inactive: sorted to the end
    } else {
        float3 c = f3((a.x + b.x) * 0.5f, (a.y + b.y) * 0.5f, (a.z + b.z) * 0.5f);
        float3 e = f3(fmaxf(bmx.x - bmn.x, 1e-30f), fmaxf(bmx.y - bmn.y, 1e-30f), fmaxf(bmx.z - bmn.z, 1e-30f));
        const float s = 2097151.0f;
        uint32_t x = (uint32_t)fminf(fmaxf((c.x - bmn.x) / e.x * s, 0.0f), s);
        uint32_t y = (uint32_t)fminf(fmaxf((c.y - bmn.y) / e.y * s, 0.0f), s);
        uint32_t z = (uint32_t)fminf(fmaxf((c.z - bmn.z) / e.z * s, 0.0f), s);
        key = (expand21(x) << 2) | (expand21(y) << 1) | expand21(z);
    }
    keys[i] = key;
    order[i] = i;
}

// ------------------------------------------------------------------------------------------- Karras 2012
__device__ __forceinline__ int delta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// internal node i in [0, n-1); leaf j is node (n-1)+j
__global__ void k_karras(const uint64_t* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right, int* __restrict__ parent,
                         int* __restrict__ first, int* __restrict__ count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int l_idx = (lo == gamma) ? (n - 1 + gamma) : gamma;
    int r_idx = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    left[i] = l_idx;
    right[i] = r_idx;
    parent[l_idx] = i;
    parent[r_idx] = i;
    first[i] = lo;
    count[i] = hi - lo + 1;
    if (i == 0) parent[0] = -1;
}

__global__ void k_lbvh_leaves(const float4* __restrict__ plo, const float4* __restrict__ phi, const uint32_t* __restrict__ order, int n,
                              float4* __restrict__ nlo, float4* __restrict__ nhi, int* __restrict__ left, int* __restrict__ right,
                              int* __restrict__ first, int* __restrict__ count, int* __restrict__ flag) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int node = n - 1 + j;
    uint32_t p = order[j];
    nlo[node] = plo[p];
    nhi[node] = phi[p];
    left[node] = -1;
    right[node] = -1;
    first[node] = j;
    count[node] = 1;
    if (j < n - 1) flag[j] = 0;
}

__global__ void k_lbvh_fit(int n, const int* __restrict__ parent, const int* __restrict__ left, const int* __restrict__ right,
                           float4* __restrict__ nlo, float4* __restrict__ nhi, int* __restrict__ flag) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int node = parent[n - 1 + j];
    while (node >= 0) {
        __threadfence();
        if (atomicAdd(&flag[node], 1) == 0) return;  // first arrival: the sibling subtree is not done yet
        __threadfence();
        int l = left[node], r = right[node];
        float4 al = __ldcg(&nlo[l]), ah = __ldcg(&nhi[l]), bl = __ldcg(&nlo[r]), bh = __ldcg(&nhi[r]);
        // inverted (inactive) boxes are ignored by min/max only if both are inverted; handle explicitly
        bool av = al.x <= ah.x, bv = bl.x <= bh.x;
        float4 lo4, hi4;
        if (av && bv) {
            lo4 = f4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0);
            hi4 = f4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0);
        } else if (av) { lo4 = al; hi4 = ah; }
        else { lo4 = bl; hi4 = bh; }
        nlo[node] = lo4;
        nhi[node] = hi4;
        node = parent[node];
    }
}

// ------------------------------------------------------------------------------------------- wide-node writer
struct ChildBox { float3 lo, hi; uint32_t ref; };

__device__ void write_node4(Node4* out, int nchild, const ChildBox* ch) {
    Node4 node;
    const float inf = __int_as_float(0x7f800000);
    for (int c = 0; c < 4; ++c) {
        const bool v = c < nchild;
        node.lo_x[c] = v ? ch[c].lo.x : inf; node.hi_x[c] = v ? ch[c].hi.x : -inf;
        node.lo_y[c] = v ? ch[c].lo.y : inf; node.hi_y[c] = v ? ch[c].hi.y : -inf;
        node.lo_z[c] = v ? ch[c].lo.z : inf; node.hi_z[c] = v ? ch[c].hi.z : -inf;
        node.child[c] = v ? ch[c].ref : GK_CHILD_EMPTY;
    }
    node.n = (uint32_t)nchild;
    node.pad[0] = node.pad[1] = node.pad[2] = 0;
    *out = node;
}

__device__ __forceinline__ float box_area(float4 lo, float4 hi) {
    float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// Node8 (gk_bvh.cuh): quantise the child boxes onto the node's power-of-two grid, rounded outward. The checks use directed
// roundings of o + q * s, so a plane is moved one more cell out whenever the rounded sum could lie on the wrong side.
__device__ void write_node8(Node8* out, int nk, int ni, const ChildBox* ch, uint32_t child_base, uint32_t lb) {
    float3 plo = f3(FLT_MAX, FLT_MAX, FLT_MAX), phi = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int c = 0; c < nk; ++c) {
        plo = f3(fminf(plo.x, ch[c].lo.x), fminf(plo.y, ch[c].lo.y), fminf(plo.z, ch[c].lo.z));
        phi = f3(fmaxf(phi.x, ch[c].hi.x), fmaxf(phi.y, ch[c].hi.y), fmaxf(phi.z, ch[c].hi.z));
    }
    const float o[3] = {plo.x, plo.y, plo.z}, e[3] = {phi.x - plo.x, phi.y - plo.y, phi.z - plo.z};
    float sc[3];
    for (int a = 0; a < 3; ++a) {
        int ex = 0;
        const float need = e[a] / GK_Q_MAX * 1.000001f;  // the subtraction above was rounded
        sc[a] = 1.0f;
        if (need > 0.0f && need < 1e30f) { frexpf(need, &ex); sc[a] = ldexpf(1.0f, ex); }
        sc[a] = fmaxf(sc[a], 1e-18f);
    }
    Node8 node;
    node.ox = o[0]; node.oy = o[1]; node.oz = o[2];
    node.sx = sc[0]; node.sy = sc[1]; node.sz = sc[2];
    node.link = child_base | ((uint32_t)nk << 24) | ((uint32_t)ni << 28);
    node.lb = lb;
    for (int c = 0; c < 8; ++c) {
        if (c >= nk) {
            for (int a = 0; a < 3; ++a) { node.q[2 * a][c] = 0x7C00; node.q[2 * a + 1][c] = 0xFC00; }  // +inf / -inf: never hit
            continue;
        }
        const float lo[3] = {ch[c].lo.x, ch[c].lo.y, ch[c].lo.z}, hi[3] = {ch[c].hi.x, ch[c].hi.y, ch[c].hi.z};
        for (int a = 0; a < 3; ++a) {
            float ql = floorf((lo[a] - o[a]) / sc[a]), qh = ceilf((hi[a] - o[a]) / sc[a]);
            ql = fmaxf(ql, 0.0f);
            if (__fmaf_ru(ql, sc[a], o[a]) > lo[a]) ql = fmaxf(ql - 1.0f, 0.0f);  // o <= lo, so q = 0 is always on the right side
            if (__fmaf_rd(qh, sc[a], o[a]) < hi[a]) qh += 1.0f;
            if (__fmaf_rd(qh, sc[a], o[a]) < hi[a]) qh += 1.0f;
            node.q[2 * a][c] = __half_as_ushort(__float2half_rd(ql));   // integers <= 2048 are exact in fp16; the directed
            node.q[2 * a + 1][c] = __half_as_ushort(__float2half_ru(qh));  // conversions only matter if the grid were ever exceeded
        }
    }
    *out = node;
}

// One BFS level of the BVH2 -> Node8 collapse. Like collapse_item below, but up to eight children, the internal children in
// the first slots and in consecutive node slots, the leaf children behind them with consecutive leaf-table entries.
__device__ __forceinline__ void collapse_item8(int i, const int2* qin, int2* qout, int* nout, int* node_counter, int* leaf_counter,
                                  const float4* nlo, const float4* nhi, const int* left, const int* right, const int* first, const int* count,
                                  int leaf_max, const uint32_t* order, const uint32_t* remap, int tlas_mode, uint32_t tri_base_abs, Node8* out,
                                  uint32_t* leaf_out, float4* out_lo, float4* out_hi) {
    const int2 e = qin[i];
    int kids[8];
    int nk;
#define GK_IS_LEAF2(n2) (left[n2] < 0 || count[n2] <= leaf_max)
    if (GK_IS_LEAF2(e.x)) {  // degenerate: the whole BVH is one leaf
        kids[0] = e.x;
        nk = 1;
    } else {
        kids[0] = left[e.x];
        kids[1] = right[e.x];
        nk = 2;
        while (nk < 8) {
            int best = -1;
            float barea = -1.0f;
            for (int k = 0; k < nk; ++k) {
                if (GK_IS_LEAF2(kids[k])) continue;
                const float a = box_area(nlo[kids[k]], nhi[kids[k]]);
                if (a > barea) { barea = a; best = k; }
            }
            if (best < 0) break;
            const int n2 = kids[best];
            kids[best] = left[n2];
            kids[nk++] = right[n2];
        }
    }
    // internal children first (stable)
    int ord[8], ni = 0;
    for (int k = 0; k < nk; ++k)
        if (!GK_IS_LEAF2(kids[k])) ord[ni++] = kids[k];
    int nl = 0;
    for (int k = 0; k < nk; ++k)
        if (GK_IS_LEAF2(kids[k])) ord[ni + nl++] = kids[k];
    uint32_t cb = 0, l0 = 0;
    if (ni) {
        cb = (uint32_t)atomicAdd(node_counter, ni);
        const int q = atomicAdd(nout, ni);
        for (int r = 0; r < ni; ++r) qout[q + r] = make_int2(ord[r], (int)cb + r);
    }
    if (nl) l0 = (uint32_t)atomicAdd(leaf_counter, nl);
    ChildBox ch[8];
    for (int k = 0; k < nk; ++k) {
        const int n2 = ord[k];
        const float4 l = nlo[n2], h = nhi[n2];
        ch[k].lo = f3(l.x, l.y, l.z);
        ch[k].hi = f3(h.x, h.y, h.z);
        ch[k].ref = 0;
        if (k >= ni) {
            const uint32_t f = (uint32_t)first[n2];
            leaf_out[l0 + (k - ni)] = tlas_mode ? remap[order[f]] : ((tri_base_abs + f) | ((uint32_t)(count[n2] - 1) << 28));
        }
    }
    write_node8(out + e.y, nk, ni, ch, cb, nl ? l0 - (uint32_t)ni : 0u);
    out_lo[e.y] = nlo[e.x];
    out_hi[e.y] = nhi[e.x];
}

// One BFS level of the BVH2 -> Node4 collapse. Entry = (bvh2 node, node4 slot). A bvh2 node with <= leaf_max primitives
// becomes a leaf reference; otherwise the largest-area internal child is opened until there are 4 children.
__device__ __forceinline__ void collapse_item(int i, const int2* qin, const int* nin_ptr, int2* qout, int* nout, int* node_counter,
                                 const float4* nlo, const float4* nhi, const int* left,
                                 const int* right, const int* first, const int* count, int leaf_max,
                                 const uint32_t* order, const uint32_t* remap, int tlas_mode, Node4* out, float4* out_lo,
                                 float4* out_hi) {
    int2 e = qin[i];
    int kids[4];
    int nk;
#define GK_IS_LEAF2(n2) (left[n2] < 0 || count[n2] <= leaf_max)
    if (GK_IS_LEAF2(e.x)) {  // degenerate: the whole BVH is one leaf
        kids[0] = e.x;
        nk = 1;
    } else {
        kids[0] = left[e.x];
        kids[1] = right[e.x];
        nk = 2;
        while (nk < 4) {
            int best = -1;
            float barea = -1.0f;
            for (int k = 0; k < nk; ++k) {
                if (GK_IS_LEAF2(kids[k])) continue;
                float a = box_area(nlo[kids[k]], nhi[kids[k]]);
                if (a > barea) { barea = a; best = k; }
            }
            if (best < 0) break;
            int n2 = kids[best];
            kids[best] = left[n2];
            kids[nk++] = right[n2];
        }
    }
    ChildBox ch[4];
    for (int k = 0; k < nk; ++k) {
        int n2 = kids[k];
        float4 l = nlo[n2], h = nhi[n2];
        ch[k].lo = f3(l.x, l.y, l.z);
        ch[k].hi = f3(h.x, h.y, h.z);
        if (GK_IS_LEAF2(n2)) {
            uint32_t f = (uint32_t)first[n2];
            if (tlas_mode) f = remap[order[f]];  // TLAS leaves carry the instance index itself
            ch[k].ref = GK_LEAF_BIT | ((uint32_t)(count[n2] - 1) << 28) | f;
        } else {
            int slot = atomicAdd(node_counter, 1);
            int q = atomicAdd(nout, 1);
            qout[q] = make_int2(n2, slot);
            ch[k].ref = (uint32_t)slot;
        }
    }
    float4 l = nlo[e.x], h = nhi[e.x];
    write_node4(out + e.y, nk, ch);
    out_lo[e.y] = l;
    out_hi[e.y] = h;
}
struct Wide8Out { int fmt; int* leaf_counter; uint32_t* leaf_out; uint32_t tri_base_abs; };
__global__ void k_collapse_level(const int2* __restrict__ qin, const int* __restrict__ nin_ptr, int2* __restrict__ qout, int* __restrict__ nout, int* __restrict__ node_counter,
                                 const float4* __restrict__ nlo, const float4* __restrict__ nhi, const int* __restrict__ left,
                                 const int* __restrict__ right, const int* __restrict__ first, const int* __restrict__ count, int leaf_max,
                                 const uint32_t* __restrict__ order, const uint32_t* __restrict__ remap, int tlas_mode, Node4* __restrict__ out, float4* __restrict__ out_lo,
                                 float4* __restrict__ out_hi, Wide8Out w8) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *nin_ptr) return;
    if (w8.fmt == 8)
        collapse_item8(i, qin, qout, nout, node_counter, w8.leaf_counter, nlo, nhi, left, right, first, count, leaf_max, order, remap, tlas_mode, w8.tri_base_abs,
                       reinterpret_cast<Node8*>(out), w8.leaf_out, out_lo, out_hi);
    else
        collapse_item(i, qin, nin_ptr, qout, nout, node_counter, nlo, nhi, left, right, first, count, leaf_max, order, remap, tlas_mode, out, out_lo, out_hi);
}

__global__ void k_write_tris(const uint16_t* __restrict__ simple, const uint32_t* __restrict__ indices, uint32_t vertex_offset,
                             uint32_t index_offset, const uint32_t* __restrict__ order, uint32_t n, float4* __restrict__ tris,
                             uint32_t* __restrict__ tri_prim) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t p = order ? order[i] : tri_prim[i];
    float3 v[3];
    for (int k = 0; k < 3; ++k) {
        uint32_t vi = vertex_offset + indices[index_offset + p * 3 + k];
        v[k] = f3(h2f(simple[vi * 4]), h2f(simple[vi * 4 + 1]), h2f(simple[vi * 4 + 2]));
    }
    tris[(size_t)i * 3 + 0] = f4(v[0].x, v[0].y, v[0].z, __uint_as_float(p));
    tris[(size_t)i * 3 + 1] = f4(v[1].x, v[1].y, v[1].z, 0);
    tris[(size_t)i * 3 + 2] = f4(v[2].x, v[2].y, v[2].z, 0);
    if (order) tri_prim[i] = p;
}

// Compact shading record of one primitive (primitive order): what FHardwareRayTracerV2::TraceRay fetches after a hit
// (assets/shaders/common/Shading.slang:812-820: three indices + three 24-byte vertices) gathered into one 32-byte record:
// normals 3 x half3, texcoords 3 x half2, material slot of vertex 0. Values are the same fp16 bits the vertex buffer holds.
__global__ void k_shade_tris(const uint2* __restrict__ vertices, const uint32_t* __restrict__ indices, uint32_t vertex_offset, uint32_t index_offset,
                             uint32_t n, uint4* __restrict__ out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint16_t nrm[3][3], uv[3][2];
    uint32_t slot = 0;
    for (int k = 0; k < 3; ++k) {
        uint32_t vi = vertex_offset + indices[index_offset + p * 3 + k];
        const uint2* v = vertices + (size_t)vi * 3;
        uint2 a = v[0], b = v[1], c = v[2];
        nrm[k][0] = (uint16_t)(b.x & 0xFFFF); nrm[k][1] = (uint16_t)(b.x >> 16); nrm[k][2] = (uint16_t)(b.y & 0xFFFF);
        uv[k][0] = (uint16_t)(a.y >> 16); uv[k][1] = (uint16_t)(b.y >> 16);
        if (k == 0) slot = (c.y >> 16) & 0xFF;
    }
    uint4 r0, r1;
    r0.x = nrm[0][0] | ((uint32_t)nrm[0][1] << 16); r0.y = nrm[0][2] | ((uint32_t)nrm[1][0] << 16);
    r0.z = nrm[1][1] | ((uint32_t)nrm[1][2] << 16); r0.w = nrm[2][0] | ((uint32_t)nrm[2][1] << 16);
    r1.x = nrm[2][2] | ((uint32_t)uv[0][0] << 16); r1.y = uv[0][1] | ((uint32_t)uv[1][0] << 16);
    r1.z = uv[1][1] | ((uint32_t)uv[2][0] << 16); r1.w = uv[2][1] | (slot << 16);
    out[(size_t)p * 2] = r0;
    out[(size_t)p * 2 + 1] = r1;
}

// Bottom-up refit of one BFS level: child boxes from triangles (leaves) or from the already refitted deeper level.
__global__ void k_refit_level(Node4* __restrict__ nodes, float4* __restrict__ node_lo, float4* __restrict__ node_hi, uint32_t begin, uint32_t end,
                              const float4* __restrict__ tris) {
    uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const uint4 refs = *(reinterpret_cast<const uint4*>(nodes + i) + 6);
    const uint32_t child[4] = {refs.x, refs.y, refs.z, refs.w};
    ChildBox ch[4];
    int nk = (int)nodes[i].n;
    float3 nlo = f3(FLT_MAX, FLT_MAX, FLT_MAX), nhi = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int k = 0; k < nk; ++k) {
        uint32_t ref = child[k];
        float3 lo, hi;
        if (ref & GK_LEAF_BIT) {
            uint32_t first = ref & 0x0FFFFFFFu, cnt = ((ref >> 28) & 7u) + 1u;
            lo = f3(FLT_MAX, FLT_MAX, FLT_MAX);
            hi = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
            for (uint32_t t = 0; t < cnt * 3; ++t) {
                float4 v = tris[(size_t)first * 3 + t];
                lo = f3(fminf(lo.x, v.x), fminf(lo.y, v.y), fminf(lo.z, v.z));
                hi = f3(fmaxf(hi.x, v.x), fmaxf(hi.y, v.y), fmaxf(hi.z, v.z));
            }
        } else {
            float4 l = node_lo[ref], h = node_hi[ref];
            lo = f3(l.x, l.y, l.z);
            hi = f3(h.x, h.y, h.z);
        }
        ch[k].lo = lo; ch[k].hi = hi; ch[k].ref = ref;
        nlo = f3(fminf(nlo.x, lo.x), fminf(nlo.y, lo.y), fminf(nlo.z, lo.z));
        nhi = f3(fmaxf(nhi.x, hi.x), fmaxf(nhi.y, hi.y), fmaxf(nhi.z, hi.z));
    }
    write_node4(nodes + i, nk, ch);
    node_lo[i] = f4(nlo.x, nlo.y, nlo.z, 0);
    node_hi[i] = f4(nhi.x, nhi.y, nhi.z, 0);
}

// Node8 refit of one BFS level: the child boxes come from the triangles (leaf-table ranges, absolute in the triangle pool) or
// from the already refitted deeper level; the node is re-quantised on its new grid, links and slot order stay.
__global__ void k_refit_level8(Node8* __restrict__ nodes, float4* __restrict__ node_lo, float4* __restrict__ node_hi, uint32_t begin, uint32_t end,
                               const float4* __restrict__ tris_pool, const uint32_t* __restrict__ leaf) {
    uint32_t i = begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    const uint32_t link = nodes[i].link, lb = nodes[i].lb;
    const int nk = (int)((link >> 24) & 15u), ni = (int)(link >> 28);
    const uint32_t cb = link & 0x00FFFFFFu;
    ChildBox ch[8];
    float3 nlo = f3(FLT_MAX, FLT_MAX, FLT_MAX), nhi = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int k = 0; k < nk; ++k) {
        float3 lo, hi;
        if (k >= ni) {
            const uint32_t rec = leaf[lb + (uint32_t)k], first = rec & 0x0FFFFFFFu, cnt = (rec >> 28) + 1u;
            lo = f3(FLT_MAX, FLT_MAX, FLT_MAX);
            hi = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
            for (uint32_t t = 0; t < cnt * 3; ++t) {
                const float4 v = tris_pool[(size_t)first * 3 + t];
                lo = f3(fminf(lo.x, v.x), fminf(lo.y, v.y), fminf(lo.z, v.z));
                hi = f3(fmaxf(hi.x, v.x), fmaxf(hi.y, v.y), fmaxf(hi.z, v.z));
            }
        } else {
            const float4 l = node_lo[cb + (uint32_t)k], h = node_hi[cb + (uint32_t)k];
            lo = f3(l.x, l.y, l.z);
            hi = f3(h.x, h.y, h.z);
        }
        ch[k].lo = lo; ch[k].hi = hi; ch[k].ref = 0;
        nlo = f3(fminf(nlo.x, lo.x), fminf(nlo.y, lo.y), fminf(nlo.z, lo.z));
        nhi = f3(fmaxf(nhi.x, hi.x), fmaxf(nhi.y, hi.y), fmaxf(nhi.z, hi.z));
    }
    write_node8(nodes + i, nk, ni, ch, cb, lb);
    node_lo[i] = f4(nlo.x, nlo.y, nlo.z, 0);
    node_hi[i] = f4(nhi.x, nhi.y, nhi.z, 0);
}

// ------------------------------------------------------------------------------------------- SAH-binned BVH2 (BLAS)
// Level-synchronous top-down build: every level bins all primitives of the still-open nodes (16 bins x 3 axes, integer
// atomics on order-encoded floats so the tree is deterministic), picks the SAH-best plane per node, and partitions.
struct SahBin { int lo[3], hi[3], cnt, pad; };  // 32 B

__global__ void k_sah_init(int leaf_max, int n, const int* __restrict__ bounds, float4* nlo, float4* nhi, int* left, int* right, int* first, int* count,
                           uint32_t* order, int* pos_node, int* active, int* counters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { order[i] = i; pos_node[i] = 0; }
    if (i == 0) {
        nlo[0] = f4(o2f(bounds[0]), o2f(bounds[1]), o2f(bounds[2]), 0);
        nhi[0] = f4(o2f(bounds[3]), o2f(bounds[4]), o2f(bounds[5]), 0);
        left[0] = right[0] = -1;
        first[0] = 0;
        count[0] = n;
        active[0] = 0;      // active list: node ids; slot s <-> bins[s]
        counters[0] = 1;    // node count
        counters[1] = (n > leaf_max) ? 1 : 0;  // active count (this level)
        counters[2] = 0;    // active count (next level)
    }
}

__global__ void k_sah_clear_bins(SahBin* bins, const int* __restrict__ nslots_ptr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *nslots_ptr * 3 * kSahBins) return;
    SahBin b;
    b.lo[0] = b.lo[1] = b.lo[2] = 0x7FFFFFFF;
    b.hi[0] = b.hi[1] = b.hi[2] = (int)0x80000000;
    b.cnt = 0;
    b.pad = 0;
    bins[i] = b;
}

__device__ __forceinline__ int sah_bin_of(float c, float lo, float hi) {
    float e = hi - lo;
    if (!(e > 0.0f)) return 0;
    int b = (int)((c - lo) / e * (float)kSahBins);
    return max(0, min(kSahBins - 1, b));
}

// slot_of[node] = active slot or -1
// slot_of[] is either reset to -1 after every level (the multi-kernel build) or left as written (the fused build, which saves the
// pass and its barrier): there an entry is valid only if the level's active list names the node at that slot.
__device__ __forceinline__ int sah_slot_of(int node, const int* slot_of, const int* active, int nactive) {
    int slot = slot_of[node];
    if (active && (slot < 0 || slot >= nactive || active[slot] != node)) slot = -1;
    return slot;
}
__device__ __forceinline__ void sah_bin_item(int i, int n, const uint32_t* order, const int* pos_node, const int* slot_of,
                          const float4* plo, const float4* phi, const float4* nlo,
                          const float4* nhi, SahBin* bins, const int* active = nullptr, int nactive = 0) {
    int node = pos_node[i];
    int slot = sah_slot_of(node, slot_of, active, nactive);
    if (slot < 0) return;
    uint32_t p = order[i];
    float4 a = plo[p], b = phi[p], l = nlo[node], h = nhi[node];
    float c[3] = {(a.x + b.x) * 0.5f, (a.y + b.y) * 0.5f, (a.z + b.z) * 0.5f};
    float nl[3] = {l.x, l.y, l.z}, nh[3] = {h.x, h.y, h.z};
    int ia[3] = {f2o(a.x), f2o(a.y), f2o(a.z)}, ib[3] = {f2o(b.x), f2o(b.y), f2o(b.z)};
    for (int ax = 0; ax < 3; ++ax) {
        SahBin* bn = bins + ((size_t)slot * 3 + ax) * kSahBins + sah_bin_of(c[ax], nl[ax], nh[ax]);
        atomicMin(&bn->lo[0], ia[0]); atomicMin(&bn->lo[1], ia[1]); atomicMin(&bn->lo[2], ia[2]);
        atomicMax(&bn->hi[0], ib[0]); atomicMax(&bn->hi[1], ib[1]); atomicMax(&bn->hi[2], ib[2]);
        atomicAdd(&bn->cnt, 1);
    }
}
__global__ void k_sah_bin(int n, const uint32_t* __restrict__ order, const int* __restrict__ pos_node, const int* __restrict__ slot_of,
                          const float4* __restrict__ plo, const float4* __restrict__ phi, const float4* __restrict__ nlo,
                          const float4* __restrict__ nhi, SahBin* __restrict__ bins) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sah_bin_item(i, n, order, pos_node, slot_of, plo, phi, nlo, nhi, bins);
}

struct SahSplit { int axis, plane, nl, fallback; };

__device__ __forceinline__ void sah_split_item(int s, int leaf_max, const int* nactive_ptr, const int* active, const SahBin* bins, float4* nlo, float4* nhi, int* left,
                            int* right, int* first, int* count, int* counters, int* next_active, SahSplit* splits, int* cursors, int* next_count = nullptr) {
    if (!next_count) next_count = &counters[2];
    int node = active[s];
    int cnt = count[node];
    float bestCost = FLT_MAX;
    int bestAxis = -1, bestPlane = 0, bestNl = 0;
    for (int ax = 0; ax < 3; ++ax) {
        const SahBin* b = bins + ((size_t)s * 3 + ax) * kSahBins;
        // right-to-left suffix areas
        float rarea[kSahBins];
        int rcnt[kSahBins];
        int lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
        int c = 0;
        for (int k = kSahBins - 1; k >= 1; --k) {
            if (b[k].cnt) {
                for (int a = 0; a < 3; ++a) { lo[a] = min(lo[a], b[k].lo[a]); hi[a] = max(hi[a], b[k].hi[a]); }
                c += b[k].cnt;
            }
            rcnt[k] = c;
            rarea[k] = c ? box_area(f4(o2f(lo[0]), o2f(lo[1]), o2f(lo[2]), 0), f4(o2f(hi[0]), o2f(hi[1]), o2f(hi[2]), 0)) : 0.0f;
        }
        for (int a = 0; a < 3; ++a) { lo[a] = 0x7FFFFFFF; hi[a] = (int)0x80000000; }
        c = 0;
        for (int k = 0; k < kSahBins - 1; ++k) {
            if (b[k].cnt) {
                for (int a = 0; a < 3; ++a) { lo[a] = min(lo[a], b[k].lo[a]); hi[a] = max(hi[a], b[k].hi[a]); }
                c += b[k].cnt;
            }
            if (c == 0 || rcnt[k + 1] == 0) continue;
            float la = box_area(f4(o2f(lo[0]), o2f(lo[1]), o2f(lo[2]), 0), f4(o2f(hi[0]), o2f(hi[1]), o2f(hi[2]), 0));
            float cost = la * (float)c + rarea[k + 1] * (float)rcnt[k + 1];
            if (cost < bestCost) { bestCost = cost; bestAxis = ax; bestPlane = k; bestNl = c; }
        }
    }
    SahSplit sp;
    sp.axis = bestAxis;
    sp.plane = bestPlane;
    sp.nl = bestNl;
    sp.fallback = 0;
    float4 l = nlo[node], h = nhi[node];
    float4 llo = l, lhi = h, rlo = l, rhi = h;
    if (bestAxis < 0) {
        // every centroid fell into one bin on every axis: median split by position, children keep the parent box
        // a BLAS leaf may hold up to 8 triangles (leaf encoding); a TLAS leaf holds exactly one instance, so coincident
        // instances are always separated by the median split
        if (leaf_max > 1 && cnt <= 8) { splits[s].axis = -2; return; }
        sp.fallback = 1;
        sp.nl = cnt / 2;
    } else {
        const SahBin* b = bins + ((size_t)s * 3 + bestAxis) * kSahBins;
        int lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
        for (int k = 0; k <= bestPlane; ++k)
            if (b[k].cnt) for (int a = 0; a < 3; ++a) { lo[a] = min(lo[a], b[k].lo[a]); hi[a] = max(hi[a], b[k].hi[a]); }
        llo = f4(o2f(lo[0]), o2f(lo[1]), o2f(lo[2]), 0);
        lhi = f4(o2f(hi[0]), o2f(hi[1]), o2f(hi[2]), 0);
        for (int a = 0; a < 3; ++a) { lo[a] = 0x7FFFFFFF; hi[a] = (int)0x80000000; }
        for (int k = bestPlane + 1; k < kSahBins; ++k)
            if (b[k].cnt) for (int a = 0; a < 3; ++a) { lo[a] = min(lo[a], b[k].lo[a]); hi[a] = max(hi[a], b[k].hi[a]); }
        rlo = f4(o2f(lo[0]), o2f(lo[1]), o2f(lo[2]), 0);
        rhi = f4(o2f(hi[0]), o2f(hi[1]), o2f(hi[2]), 0);
    }
    int c0 = atomicAdd(&counters[0], 2);
    int f = first[node];
    left[node] = c0;
    right[node] = c0 + 1;
    nlo[c0] = llo; nhi[c0] = lhi; left[c0] = right[c0] = -1; first[c0] = f; count[c0] = sp.nl;
    nlo[c0 + 1] = rlo; nhi[c0 + 1] = rhi; left[c0 + 1] = right[c0 + 1] = -1; first[c0 + 1] = f + sp.nl; count[c0 + 1] = cnt - sp.nl;
    splits[s] = sp;
    cursors[2 * s] = 0;
    cursors[2 * s + 1] = 0;
    if (sp.nl > leaf_max) next_active[atomicAdd(next_count, 1)] = c0;
    if (cnt - sp.nl > leaf_max) next_active[atomicAdd(next_count, 1)] = c0 + 1;
}
__global__ void k_sah_split(int leaf_max, const int* __restrict__ nactive_ptr, const int* __restrict__ active, const SahBin* __restrict__ bins, float4* nlo, float4* nhi, int* left,
                            int* right, int* first, int* count, int* counters, int* next_active, SahSplit* splits, int* cursors) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= *nactive_ptr) return;
    sah_split_item(s, leaf_max, nactive_ptr, active, bins, nlo, nhi, left, right, first, count, counters, next_active, splits, cursors);
}

__device__ __forceinline__ void sah_partition_item(int i, int n, const uint32_t* order, const int* pos_node, const int* slot_of,
                                const float4* plo, const float4* phi, const float4* nlo,
                                const float4* nhi, const int* left, const int* first,
                                const SahSplit* splits, int* cursors, uint32_t* order_out,
                                int* pos_node_out, const int* active = nullptr, int nactive = 0) {
    int node = pos_node[i];
    int slot = sah_slot_of(node, slot_of, active, nactive);
    uint32_t p = order[i];
    if (slot < 0 || splits[slot].axis == -2) {
        order_out[i] = p;
        pos_node_out[i] = node;
        return;
    }
    SahSplit sp = splits[slot];
    int f = first[node];
    bool goLeft;
    if (sp.fallback) {
        goLeft = (i - f) < sp.nl;
    } else {
        float4 a = plo[p], b = phi[p], l = nlo[node], h = nhi[node];
        float c = sp.axis == 0 ? (a.x + b.x) * 0.5f : (sp.axis == 1 ? (a.y + b.y) * 0.5f : (a.z + b.z) * 0.5f);
        float nl = sp.axis == 0 ? l.x : (sp.axis == 1 ? l.y : l.z), nh = sp.axis == 0 ? h.x : (sp.axis == 1 ? h.y : h.z);
        goLeft = sah_bin_of(c, nl, nh) <= sp.plane;
    }
    int dst;
    if (sp.fallback) dst = i;
    else {
        // warp-aggregated cursor bump: lanes going to the same side of the same node take one atomic together (near the root
        // every box of the build would otherwise hit the same two counters)
        const int key = 2 * slot + (goLeft ? 0 : 1);
        const unsigned act = __activemask();
        const unsigned peers = __match_any_sync(act, key);
        const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&cursors[key], __popc(peers));
        base = __shfl_sync(peers, base, leader);
        dst = f + (goLeft ? 0 : sp.nl) + base + __popc(peers & ((1u << lane) - 1u));
    }
    order_out[dst] = p;
    pos_node_out[dst] = left[node] + (goLeft ? 0 : 1);
}
__global__ void k_sah_partition(int n, const uint32_t* __restrict__ order, const int* __restrict__ pos_node, const int* __restrict__ slot_of,
                                const float4* __restrict__ plo, const float4* __restrict__ phi, const float4* __restrict__ nlo,
                                const float4* __restrict__ nhi, const int* __restrict__ left, const int* __restrict__ first,
                                const SahSplit* __restrict__ splits, int* __restrict__ cursors, uint32_t* __restrict__ order_out,
                                int* __restrict__ pos_node_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sah_partition_item(i, n, order, pos_node, slot_of, plo, phi, nlo, nhi, left, first, splits, cursors, order_out, pos_node_out);
}

__global__ void k_fill_int(int* p, int n, int v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_mark_slots(const int* active, const int* __restrict__ nactive_ptr, int* slot_of, int unmark) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *nactive_ptr) slot_of[active[i]] = unmark ? -1 : i;
}
// level bookkeeping on the device, so the host only synchronises every few levels
__global__ void k_sah_next(int* counters) {
    counters[1] = counters[2];
    counters[2] = 0;
}
__global__ void k_collapse_next(int* counters, int level) {
    counters[16 + level + 2] = counters[8];  // nodes allocated so far = first node of level + 2
    counters[10] = counters[9];
    counters[9] = 0;
}


// ------------------------------------------------------------------------------------------- fused small build
// The whole SAH build + wide collapse of a small input (the per-frame TLAS of a few thousand instances) in ONE block: the
// level-synchronous phases above run back to back with __syncthreads() between them instead of ~130 tiny launches and three
// host round trips (0.93 ms -> see profiles/r01_experiments.md for 1.9 k instances). Same per-item code, same tree.
constexpr int kFusedSmemSlots = 24;  // open nodes whose bins fit the block-private copy (24 x 48 bins x 32 B = 36 KB)
struct FusedBuild {
    int n, leaf_max, max_slots;
    const float4 *plo, *phi;
    int* bounds;
    float4 *nlo, *nhi;
    int *left, *right, *first, *count, *slot_of;
    uint32_t *ord_a, *ord_b;
    int *pn_a, *pn_b, *act_a, *act_b, *cursors, *counters;
    SahSplit* splits;
    SahBin* bins;
    int2 *qa, *qb;
    const uint32_t* remap;
    Node4* out;
    float4 *out_lo, *out_hi;
    Wide8Out w8;
};
template <bool GRID>
__global__ void __launch_bounds__(GRID ? 256 : 1024) k_build_fused(FusedBuild B) {
    // GRID: a cooperative launch of a few blocks, phases separated by grid-wide barriers; otherwise one block and __syncthreads()
    const int T = blockDim.x * gridDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ SahBin s_bins[kFusedSmemSlots * 3 * kSahBins];
    auto sync = [&]() {
        if (GRID) cooperative_groups::this_grid().sync();
        else __syncthreads();
    };
    const int n = B.n;
    // bounds of the active boxes
    if (t < 3) B.bounds[t] = 0x7FFFFFFF;
    else if (t < 6) B.bounds[t] = (int)0x80000000;
    sync();
    {
        float3 mn = f3(FLT_MAX, FLT_MAX, FLT_MAX), mx = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
        for (int i = t; i < n; i += T) {
            const float4 a = B.plo[i], b = B.phi[i];
            if (a.x <= b.x) {
                mn = f3(fminf(mn.x, a.x), fminf(mn.y, a.y), fminf(mn.z, a.z));
                mx = f3(fmaxf(mx.x, b.x), fmaxf(mx.y, b.y), fmaxf(mx.z, b.z));
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            mn.x = fminf(mn.x, __shfl_xor_sync(0xffffffffu, mn.x, o)); mn.y = fminf(mn.y, __shfl_xor_sync(0xffffffffu, mn.y, o));
            mn.z = fminf(mn.z, __shfl_xor_sync(0xffffffffu, mn.z, o)); mx.x = fmaxf(mx.x, __shfl_xor_sync(0xffffffffu, mx.x, o));
            mx.y = fmaxf(mx.y, __shfl_xor_sync(0xffffffffu, mx.y, o)); mx.z = fmaxf(mx.z, __shfl_xor_sync(0xffffffffu, mx.z, o));
        }
        if ((t & 31) == 0 && mn.x <= mx.x) {
            atomicMin(&B.bounds[0], f2o(mn.x)); atomicMin(&B.bounds[1], f2o(mn.y)); atomicMin(&B.bounds[2], f2o(mn.z));
            atomicMax(&B.bounds[3], f2o(mx.x)); atomicMax(&B.bounds[4], f2o(mx.y)); atomicMax(&B.bounds[5], f2o(mx.z));
        }
    }
    sync();
    uint32_t *ord_a = B.ord_a, *ord_b = B.ord_b;
    int *pn_a = B.pn_a, *pn_b = B.pn_b, *act_a = B.act_a, *act_b = B.act_b;
    int* counters = B.counters;
    auto clear_bins = [&](int nslots) {
        for (int i = t; i < nslots * 3 * kSahBins; i += T) {
            SahBin b;
            b.lo[0] = b.lo[1] = b.lo[2] = 0x7FFFFFFF;
            b.hi[0] = b.hi[1] = b.hi[2] = (int)0x80000000;
            b.cnt = 0;
            b.pad = 0;
            B.bins[i] = b;
        }
    };
    // Three barriers per level (bin | split | partition) instead of six: the open-node counts rotate through three words, so the count
    // a level accumulates for its successor was zeroed a whole level earlier; slot_of[] is never reset (a stale entry fails the check
    // against the level's active list, sah_slot_of); the partition phase also clears the bins and marks the slots of the next level.
    // counters[1 + L % 3] = open nodes of level L.
    for (int i = t; i < n; i += T) { ord_a[i] = i; pn_a[i] = 0; }
    for (int i = t; i < 2 * n + 2; i += T) B.slot_of[i] = i == 0 ? 0 : -1;
    clear_bins(1);
    if (t == 0) {  // k_sah_init
        B.nlo[0] = f4(o2f(B.bounds[0]), o2f(B.bounds[1]), o2f(B.bounds[2]), 0);
        B.nhi[0] = f4(o2f(B.bounds[3]), o2f(B.bounds[4]), o2f(B.bounds[5]), 0);
        B.left[0] = B.right[0] = -1;
        B.first[0] = 0;
        B.count[0] = n;
        act_a[0] = 0;
        counters[0] = 1;
        counters[1] = (n > B.leaf_max) ? 1 : 0;
        counters[2] = 0;
        counters[3] = 0;
    }
    sync();
    for (int level = 0; level < 512; ++level) {
        const int nactive = counters[1 + level % 3];
        if (nactive == 0) break;
        int* next_count = counters + 1 + (level + 1) % 3;
        if (t == 0) counters[1 + (level + 2) % 3] = 0;  // the count of level - 1: every thread read it before that level's barriers
        if (nactive <= kFusedSmemSlots) {
            // few open nodes, every box hits the same handful of bins: bin into a block-private copy in shared memory first and
            // merge once per block (the top levels of a 100 k build otherwise serialise ~2 M same-address atomics each)
            for (int i = threadIdx.x; i < nactive * 3 * kSahBins; i += blockDim.x) {
                SahBin b;
                b.lo[0] = b.lo[1] = b.lo[2] = 0x7FFFFFFF;
                b.hi[0] = b.hi[1] = b.hi[2] = (int)0x80000000;
                b.cnt = 0;
                b.pad = 0;
                s_bins[i] = b;
            }
            __syncthreads();
            for (int i = t; i < n; i += T) sah_bin_item(i, n, ord_a, pn_a, B.slot_of, B.plo, B.phi, B.nlo, B.nhi, s_bins, act_a, nactive);
            __syncthreads();
            for (int i = threadIdx.x; i < nactive * 3 * kSahBins; i += blockDim.x) {
                const SahBin b = s_bins[i];
                if (b.cnt) {
                    SahBin* g = B.bins + i;
                    atomicMin(&g->lo[0], b.lo[0]); atomicMin(&g->lo[1], b.lo[1]); atomicMin(&g->lo[2], b.lo[2]);
                    atomicMax(&g->hi[0], b.hi[0]); atomicMax(&g->hi[1], b.hi[1]); atomicMax(&g->hi[2], b.hi[2]);
                    atomicAdd(&g->cnt, b.cnt);
                }
            }
        } else {
            for (int i = t; i < n; i += T) sah_bin_item(i, n, ord_a, pn_a, B.slot_of, B.plo, B.phi, B.nlo, B.nhi, B.bins, act_a, nactive);
        }
        sync();
        for (int sl = t; sl < nactive; sl += T)
            sah_split_item(sl, B.leaf_max, counters + 1, act_a, B.bins, B.nlo, B.nhi, B.left, B.right, B.first, B.count, counters, act_b, B.splits, B.cursors, next_count);
        sync();
        for (int i = t; i < n; i += T)
            sah_partition_item(i, n, ord_a, pn_a, B.slot_of, B.plo, B.phi, B.nlo, B.nhi, B.left, B.first, B.splits, B.cursors, ord_b, pn_b, act_a, nactive);
        {   // next level: its bins and slots (the children named in act_b are new nodes, no item of this phase looks them up)
            const int nnext = *next_count;
            clear_bins(nnext);
            for (int i = t; i < nnext; i += T) B.slot_of[act_b[i]] = i;
        }
        sync();
        uint32_t* to = ord_a; ord_a = ord_b; ord_b = to;
        int* tp = pn_a; pn_a = pn_b; pn_b = tp;
        int* ta = act_a; act_a = act_b; act_b = ta;
    }
    // collapse to Node4, BFS, one barrier per level; counters[8] node counter, counters[4 + L % 3] queue size of level L (rotating like the
    // open-node counts), [10] what is left in the queue at the end (0 unless the tree is too deep), [11] levels run, [16 + L] level table:
    // every queued item allocates exactly one node, so level L + 1 begins at the first node of level L plus the size of level L's queue.
    int2 *qa = B.qa, *qb = B.qb;
    if (t == 0) { qa[0] = make_int2(0, 0); counters[8] = 1; counters[4] = 1; counters[5] = 0; counters[6] = 0; counters[9] = 0; counters[10] = 0; counters[11] = 0; counters[12] = (int)GK_LEAF_SLACK; counters[13] = counters[14] = counters[15] = 0; }  // 13-15: unused, published with the rest
    for (int i = t; i < 48; i += T) counters[16 + i] = 0;
    sync();
    int run = 1;  // nodes allocated before this level's items run = first node of the next level
    int level = 0;
    for (; level < 46; ++level) {
        const int nin = counters[4 + level % 3];
        if (level > 0) {
            run += nin;
            if (t == 0) { counters[16 + level + 1] = run; counters[11] = level; }
        }
        if (nin == 0) break;
        int* nout = counters + 4 + (level + 1) % 3;
        if (t == 0) counters[4 + (level + 2) % 3] = 0;
        for (int i = t; i < nin; i += T) {
            if (B.w8.fmt == 8)
                collapse_item8(i, qa, qb, nout, counters + 8, counters + 12, B.nlo, B.nhi, B.left, B.right, B.first, B.count, B.leaf_max, ord_a, B.remap,
                               B.remap ? 1 : 0, B.w8.tri_base_abs, reinterpret_cast<Node8*>(B.out), B.w8.leaf_out, B.out_lo, B.out_hi);
            else
                collapse_item(i, qa, nullptr, qb, nout, counters + 8, B.nlo, B.nhi, B.left, B.right, B.first, B.count, B.leaf_max, ord_a, B.remap,
                              B.remap ? 1 : 0, B.out, B.out_lo, B.out_hi);
        }
        sync();
        int2* tq = qa; qa = qb; qb = tq;
    }
    if (t == 0 && level == 46) { counters[10] = counters[4 + level % 3]; counters[11] = 46; counters[16 + 47] = run + counters[4 + level % 3]; }
}

// ------------------------------------------------------------------------------------------- instances
// VkAccelerationStructureInstanceKHR equivalent (TopLevelAccelerationStructure.cpp:83-111): instance i <-> Nodes[i],
// BLAS = modelId / 10, mask = visible && !nort (RayTraceBaseRenderer.cpp:172-179). Also the world box for the TLAS.
__global__ void k_instances(const GkNodeProxy* __restrict__ nodes, uint32_t n, const uint4* __restrict__ blas_table, uint32_t model_count,
                            InstanceRec* __restrict__ inst, float4* __restrict__ lo, float4* __restrict__ hi, uint4* __restrict__ shade_inst,
                            uint8_t* __restrict__ flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const GkNodeProxy& nd = nodes[i];
    InstanceRec r;
    uint32_t blas = nd.modelId / 10;
    bool active = nd.visible && !nd.nort && blas < model_count;
    uint4 bt = make_uint4(0, 0, 0, 0);
    if (active) {
        bt = blas_table[blas * 3];
        if (bt.z == 0) active = false;
    }
    // rows of object->world
    float a[3][3], t[3];
    for (int r_ = 0; r_ < 3; ++r_) {
        for (int c = 0; c < 3; ++c) a[r_][c] = nd.worldTS.m[c][r_];
        t[r_] = nd.worldTS.m[3][r_];
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
    for (int r_ = 0; r_ < 3; ++r_) {
        for (int c = 0; c < 3; ++c) r.w2o[r_ * 4 + c] = inv[r_][c];
        r.w2o[r_ * 4 + 3] = -(inv[r_][0] * t[0] + inv[r_][1] * t[1] + inv[r_][2] * t[2]);
    }
    r.node_base = bt.x;
    r.tri_base = bt.y;
    r.active = active ? 1u : 0u;
    r.blas = blas;
    inst[i] = r;
    // shading record: where the instance's primitives start in the shade-triangle pool + the node's first material ids
    shade_inst[(size_t)i * 2] = make_uint4(bt.w, nd.matId[0], nd.matId[1], nd.matId[2]);
    shade_inst[(size_t)i * 2 + 1] = make_uint4(nd.matId[3], nd.matId[4], nd.matId[5], nd.matId[6]);
    flags[i] = active ? 1 : 0;
    if (active) {
        float4 bl = reinterpret_cast<const float4*>(blas_table)[blas * 3 + 1], bh = reinterpret_cast<const float4*>(blas_table)[blas * 3 + 2];
        float3 mn = f3(FLT_MAX, FLT_MAX, FLT_MAX), mx = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX), mag = f3(0, 0, 0);
        for (int c = 0; c < 8; ++c) {
            float3 p = f3(c & 1 ? bh.x : bl.x, c & 2 ? bh.y : bl.y, c & 4 ? bh.z : bl.z);
            float3 w = f3(a[0][0] * p.x + a[0][1] * p.y + a[0][2] * p.z + t[0], a[1][0] * p.x + a[1][1] * p.y + a[1][2] * p.z + t[1],
                          a[2][0] * p.x + a[2][1] * p.y + a[2][2] * p.z + t[2]);
            mn = f3(fminf(mn.x, w.x), fminf(mn.y, w.y), fminf(mn.z, w.z));
            mx = f3(fmaxf(mx.x, w.x), fmaxf(mx.y, w.y), fmaxf(mx.z, w.z));
            mag = f3(fmaxf(mag.x, fabsf(a[0][0] * p.x) + fabsf(a[0][1] * p.y) + fabsf(a[0][2] * p.z) + fabsf(t[0])),
                     fmaxf(mag.y, fabsf(a[1][0] * p.x) + fabsf(a[1][1] * p.y) + fabsf(a[1][2] * p.z) + fabsf(t[1])),
                     fmaxf(mag.z, fabsf(a[2][0] * p.x) + fabsf(a[2][1] * p.y) + fabsf(a[2][2] * p.z) + fabsf(t[2])));
        }
        // The ray is intersected in object space through the fp32 inverse; the world box has to contain the image of the
        // BLAS box under that map's exact inverse, which differs from worldTS * corner by rounding in the order of
        // 1e-7 x (the magnitude of the terms summed), so the pad scales with that magnitude and not just with the extent.
        float3 pad = (mx - mn) * 1e-6f + mag * 1e-6f + f3(1e-6f, 1e-6f, 1e-6f);
        lo[i] = f4(mn.x - pad.x, mn.y - pad.y, mn.z - pad.z, 0);
        hi[i] = f4(mx.x + pad.x, mx.y + pad.y, mx.z + pad.z, 0);
    } else {
        lo[i] = f4(1e30f, 1e30f, 1e30f, 0);
        hi[i] = f4(-1e30f, -1e30f, -1e30f, 0);
    }
}

__global__ void k_gather_boxes(const uint32_t* __restrict__ ids, uint32_t m, const float4* __restrict__ lo, const float4* __restrict__ hi,
                               float4* __restrict__ out_lo, float4* __restrict__ out_hi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    out_lo[i] = lo[ids[i]];
    out_hi[i] = hi[ids[i]];
}

constexpr float kRefitCostLimit = 1.3f;  // rebuild once the refitted tree's node areas have grown by 30 %
// ------------------------------------------------------------------------------------------- TLAS refit
// Keeps the topology of the last full build and recomputes every box from this frame's instance boxes, bottom-up over the BFS levels in one
// single-block launch (a few hundred to a few ten thousand nodes; __syncthreads() between the levels). The closest hit does not depend on the
// tree (gk_bvh.cuh), so a refitted TLAS gives the same bits as a rebuilt one; what it can lose is traversal speed when instances have moved far,
// which is why the sum of the node areas comes back with it: build_tlas_impl rebuilds when it has grown past kRefitCostLimit x the value of
// the fresh tree, when the set of active instances has changed, or every GKCU_TLAS_REBUILD_EVERY frames.
struct TlasRefitArgs {
    Node4* nodes;
    float4 *node_lo, *node_hi;
    const float4 *inst_lo, *inst_hi;
    float* cost;
    int levels;
    uint32_t level_start[49];
};
__global__ void __launch_bounds__(1024) k_tlas_refit(const __grid_constant__ TlasRefitArgs A) {
    float area = 0.0f;
    for (int L = A.levels - 1; L >= 0; --L) {
        for (uint32_t i = A.level_start[L] + threadIdx.x; i < A.level_start[L + 1]; i += blockDim.x) {
            const uint4 refs = *(reinterpret_cast<const uint4*>(A.nodes + i) + 6);
            const uint32_t child[4] = {refs.x, refs.y, refs.z, refs.w};
            const int nk = (int)A.nodes[i].n;
            ChildBox ch[4];
            float3 nlo = f3(FLT_MAX, FLT_MAX, FLT_MAX), nhi = f3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
            for (int k = 0; k < nk; ++k) {
                const uint32_t ref = child[k];
                float4 l, h;
                if (ref & GK_LEAF_BIT) { l = A.inst_lo[ref & 0x0FFFFFFFu]; h = A.inst_hi[ref & 0x0FFFFFFFu]; }  // a TLAS leaf is one instance
                else { l = A.node_lo[ref]; h = A.node_hi[ref]; }
                ch[k].lo = f3(l.x, l.y, l.z); ch[k].hi = f3(h.x, h.y, h.z); ch[k].ref = ref;
                nlo = f3(fminf(nlo.x, l.x), fminf(nlo.y, l.y), fminf(nlo.z, l.z));
                nhi = f3(fmaxf(nhi.x, h.x), fmaxf(nhi.y, h.y), fmaxf(nhi.z, h.z));
            }
            write_node4(A.nodes + i, nk, ch);
            A.node_lo[i] = f4(nlo.x, nlo.y, nlo.z, 0);
            A.node_hi[i] = f4(nhi.x, nhi.y, nhi.z, 0);
            area += box_area(f4(nlo.x, nlo.y, nlo.z, 0), f4(nhi.x, nhi.y, nhi.z, 0));
        }
        __syncthreads();
    }
    __shared__ float s_area[32];
    for (int o = 16; o > 0; o >>= 1) area += __shfl_xor_sync(0xffffffffu, area, o);
    if ((threadIdx.x & 31) == 0) s_area[threadIdx.x >> 5] = area;
    __syncthreads();
    if (threadIdx.x < 32) {
        float a = threadIdx.x < (blockDim.x >> 5) ? s_area[threadIdx.x] : 0.0f;
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (threadIdx.x == 0) *A.cost = a;
    }
}
// did the set of active instances change since the topology was built? (also keeps the new flags for the next comparison)
__global__ void k_flags_changed(const uint8_t* __restrict__ flags, uint8_t* __restrict__ prev, uint32_t n, uint32_t* changed) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t f = flags[i];
    if (f != prev[i]) { *changed = 1u; prev[i] = f; }
}

inline dim3 grid_for(size_t n, int block = 256) { return dim3((unsigned)((n + block - 1) / block)); }

struct Bvh2 {
    float4 *lo, *hi;
    int *left, *right, *first, *count;
    uint32_t* order;
    int root;
};

bool reserve_bvh2(BuildScratch& s, size_t n) {
    size_t nn = 2 * n + 2;
    return s.n2_lo.reserve(nn * 16) && s.n2_hi.reserve(nn * 16) && s.n2_left.reserve(nn * 4) && s.n2_right.reserve(nn * 4) &&
           s.n2_parent.reserve(nn * 4) && s.n2_first.reserve(nn * 4) && s.n2_count.reserve(nn * 4) && s.n2_flag.reserve(nn * 4) &&
           s.keys.reserve(n * 8) && s.keys_sorted.reserve(n * 8) && s.order.reserve(n * 4) && s.order_tmp.reserve(n * 4) &&
           s.queue_a.reserve((n + 4) * 8) && s.queue_b.reserve((n + 4) * 8) && s.counters.reserve(128 * 4) && s.bounds.reserve(64);
}

// Morton LBVH over n boxes (prim_lo/prim_hi in scratch). Inactive (inverted) boxes sort to the end and are excluded.
int build_lbvh(Context& c, uint32_t n_total, uint32_t n_active, Bvh2& out) {
    BuildScratch& s = c.scratch;
    cudaStream_t st = c.stream;
    uint32_t n = n_total;
    k_init_bounds<<<1, 32, 0, st>>>(s.bounds.as<int>());
    k_bounds<<<grid_for(n), 256, 0, st>>>(s.prim_lo.as<float4>(), s.prim_hi.as<float4>(), n, s.bounds.as<int>());
    k_morton<<<grid_for(n), 256, 0, st>>>(s.prim_lo.as<float4>(), s.prim_hi.as<float4>(), n, s.bounds.as<int>(), s.keys.as<uint64_t>(),
                                          s.order_tmp.as<uint32_t>());
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, s.keys.as<uint64_t>(), s.keys_sorted.as<uint64_t>(), s.order_tmp.as<uint32_t>(),
                                    s.order.as<uint32_t>(), (int)n, 0, 64, st);
    if (!s.sort_tmp.reserve(tmp_bytes)) return 1;
    GK_CUDA(cub::DeviceRadixSort::SortPairs(s.sort_tmp.p, tmp_bytes, s.keys.as<uint64_t>(), s.keys_sorted.as<uint64_t>(),
                                            s.order_tmp.as<uint32_t>(), s.order.as<uint32_t>(), (int)n, 0, 64, st));
    c.launches += 4;
    int m = (int)n_active;  // active primitives occupy sorted positions [0, m)
    out.lo = s.n2_lo.as<float4>(); out.hi = s.n2_hi.as<float4>();
    out.left = s.n2_left.as<int>(); out.right = s.n2_right.as<int>();
    out.first = s.n2_first.as<int>(); out.count = s.n2_count.as<int>();
    out.order = s.order.as<uint32_t>();
    out.root = 0;
    if (m == 0) return 0;
    k_lbvh_leaves<<<grid_for(m), 256, 0, st>>>(s.prim_lo.as<float4>(), s.prim_hi.as<float4>(), out.order, m, out.lo, out.hi, out.left, out.right,
                                               out.first, out.count, s.n2_flag.as<int>());
    if (m > 1) {
        k_karras<<<grid_for(m - 1), 256, 0, st>>>(s.keys_sorted.as<uint64_t>(), m, out.left, out.right, s.n2_parent.as<int>(), out.first, out.count);
        k_lbvh_fit<<<grid_for(m), 256, 0, st>>>(m, s.n2_parent.as<int>(), out.left, out.right, out.lo, out.hi, s.n2_flag.as<int>());
        c.launches += 2;
    }
    c.launches += 1;
    GK_CUDA(cudaGetLastError());
    return 0;
}

// Level-synchronous SAH-binned build over n boxes (all active).
int build_sah(Context& c, uint32_t n, int leaf_max, Bvh2& out) {
    BuildScratch& s = c.scratch;
    cudaStream_t st = c.stream;
    size_t max_slots = n / (leaf_max + 1) + 2;
    if (!s.bins.reserve(max_slots * 3 * kSahBins * sizeof(SahBin)) || !s.pos_node.reserve((size_t)n * 4) || !s.pos_node_tmp.reserve((size_t)n * 4) ||
        !s.active.reserve((max_slots * 2 + 4) * 4 * 2) || !s.cursors.reserve(max_slots * 2 * 4 + max_slots * sizeof(SahSplit) + 64))
        return 1;
    int* counters = s.counters.as<int>();
    float4 *nlo = s.n2_lo.as<float4>(), *nhi = s.n2_hi.as<float4>();
    int *left = s.n2_left.as<int>(), *right = s.n2_right.as<int>(), *first = s.n2_first.as<int>(), *count = s.n2_count.as<int>();
    int* slot_of = s.n2_flag.as<int>();
    uint32_t *ord_a = s.order.as<uint32_t>(), *ord_b = s.order_tmp.as<uint32_t>();
    int *pn_a = s.pos_node.as<int>(), *pn_b = s.pos_node_tmp.as<int>();
    int* act_a = s.active.as<int>();
    int* act_b = act_a + (max_slots * 2 + 4);
    int* cursors = s.cursors.as<int>();
    SahSplit* splits = reinterpret_cast<SahSplit*>(cursors + max_slots * 2 + 16);

    k_init_bounds<<<1, 32, 0, st>>>(s.bounds.as<int>());
    k_bounds<<<grid_for(n), 256, 0, st>>>(s.prim_lo.as<float4>(), s.prim_hi.as<float4>(), n, s.bounds.as<int>());
    k_sah_init<<<grid_for(n), 256, 0, st>>>(leaf_max, (int)n, s.bounds.as<int>(), nlo, nhi, left, right, first, count, ord_a, pn_a, act_a, counters);
    k_fill_int<<<grid_for(2 * (size_t)n + 2), 256, 0, st>>>(slot_of, (int)(2 * n + 2), -1);
    c.launches += 4;
    // One level = mark, clear, bin, split, partition, unmark, next. Kernels take the number of open nodes from device memory;
    // the host only bounds the grids (a level has at most min(2^level, max_slots) open nodes) and polls every 8 levels.
    if (n > (uint32_t)leaf_max) {
        for (int level = 0; level < 512; ++level) {
            const size_t ub = std::min<size_t>(max_slots, level < 30 ? (size_t)1 << level : max_slots);
            k_mark_slots<<<grid_for(ub), 256, 0, st>>>(act_a, counters + 1, slot_of, 0);
            k_sah_clear_bins<<<grid_for(ub * 3 * kSahBins), 256, 0, st>>>(s.bins.as<SahBin>(), counters + 1);
            k_sah_bin<<<grid_for(n), 256, 0, st>>>((int)n, ord_a, pn_a, slot_of, s.prim_lo.as<float4>(), s.prim_hi.as<float4>(), nlo, nhi, s.bins.as<SahBin>());
            k_sah_split<<<grid_for(ub, 128), 128, 0, st>>>(leaf_max, counters + 1, act_a, s.bins.as<SahBin>(), nlo, nhi, left, right, first, count, counters,
                                                         act_b, splits, cursors);
            k_sah_partition<<<grid_for(n), 256, 0, st>>>((int)n, ord_a, pn_a, slot_of, s.prim_lo.as<float4>(), s.prim_hi.as<float4>(), nlo, nhi, left, first,
                                                         splits, cursors, ord_b, pn_b);
            k_mark_slots<<<grid_for(ub), 256, 0, st>>>(act_a, counters + 1, slot_of, 1);
            k_sah_next<<<1, 1, 0, st>>>(counters);
            c.launches += 7;
            std::swap(ord_a, ord_b);
            std::swap(pn_a, pn_b);
            std::swap(act_a, act_b);
            if ((level & 7) == 7) {
                int h = 0;
                GK_CUDA(cudaMemcpyAsync(&h, counters + 1, 4, cudaMemcpyDeviceToHost, st));
                GK_CUDA(cudaStreamSynchronize(st));
                if (h == 0) break;
            }
        }
    }
    out.lo = nlo; out.hi = nhi; out.left = left; out.right = right; out.first = first; out.count = count;
    out.order = ord_a;
    out.root = 0;
    return 0;
}

// BVH2 -> Node4 (BFS), returns node count and the level table.
int collapse(Context& c, const Bvh2& b, uint32_t n_prims, int leaf_max, const uint32_t* remap, Node4* out, float4* out_lo, float4* out_hi,
             WideBvhInfo& info, Wide8Out w8) {
    BuildScratch& s = c.scratch;
    cudaStream_t st = c.stream;
    int* counters = s.counters.as<int>();  // [8] node counter, [9] out-queue count
    int2 *qa = s.queue_a.as<int2>(), *qb = s.queue_b.as<int2>();
    int2 root = make_int2(b.root, 0);
    int init[5] = {1, 0, 1, 0, (int)GK_LEAF_SLACK};  // [8] node counter, [9] next-level queue size, [10] this level's queue size, [12] leaf-table cursor
    w8.leaf_counter = counters + 12;
    GK_CUDA(cudaMemcpyAsync(qa, &root, sizeof(root), cudaMemcpyHostToDevice, st));
    GK_CUDA(cudaMemcpyAsync(counters + 8, init, 20, cudaMemcpyHostToDevice, st));
    GK_CUDA(cudaMemsetAsync(counters + 16, 0, 48 * 4, st));
    (void)n_prims;
    int levels_run = 0;
    for (int level = 0; level < 46; ++level) {
        const size_t ub = level < 10 ? std::min<size_t>((size_t)n_prims + 1, (size_t)1 << (3 * level)) : (size_t)n_prims + 1;
        k_collapse_level<<<grid_for(ub, 128), 128, 0, st>>>(qa, counters + 10, qb, counters + 9, counters + 8, b.lo, b.hi, b.left, b.right, b.first, b.count,
                                                            leaf_max, b.order, remap, remap ? 1 : 0, out, out_lo, out_hi, w8);
        k_collapse_next<<<1, 1, 0, st>>>(counters, level);
        c.launches += 2;
        std::swap(qa, qb);
        levels_run = level + 1;
        if ((level & 7) == 7) {
            int h = 0;
            GK_CUDA(cudaMemcpyAsync(&h, counters + 10, 4, cudaMemcpyDeviceToHost, st));
            GK_CUDA(cudaStreamSynchronize(st));
            if (h == 0) break;
        }
    }
    // level table: counters[16 + L] = first node of level L (level 0 starts at 0, level 1 at 1)
    int table[48];
    GK_CUDA(cudaMemcpyAsync(table, counters + 16, 48 * 4, cudaMemcpyDeviceToHost, st));
    int last[5];
    GK_CUDA(cudaMemcpyAsync(last, counters + 8, 20, cudaMemcpyDeviceToHost, st));
    GK_CUDA(cudaStreamSynchronize(st));
    if (last[2] != 0) { set_error("BVH collapse did not terminate (tree deeper than 46 wide levels)"); return 1; }
    info.leaf_count = (uint32_t)last[4];
    // level table: table[L] = first node of BFS level L; table[L + 2] was written after level L had been processed
    info.level_start.clear();
    info.level_start.push_back(0);
    table[0] = 0;
    table[1] = 1;
    for (int L = 1; L <= levels_run + 1 && L < 48; ++L) {
        if (L > 1 && (uint32_t)table[L] == info.level_start.back()) break;  // no new nodes: the tree ended one level earlier
        info.level_start.push_back((uint32_t)table[L]);
    }
    if (info.level_start.back() != (uint32_t)last[0]) { set_error("BVH collapse: inconsistent level table"); return 1; }
    info.node_count = (uint32_t)last[0];
    return 0;
}

// Small inputs: bounds + SAH build + collapse in one single-block launch, one host read at the end.
int build_fused(Context& c, uint32_t n, int leaf_max, const uint32_t* remap, Node4* out, float4* out_lo, float4* out_hi, WideBvhInfo& info, Wide8Out w8) {
    BuildScratch& s = c.scratch;
    cudaStream_t st = c.stream;
    size_t max_slots = n / (leaf_max + 1) + 2;
    if (!s.bins.reserve(max_slots * 3 * kSahBins * sizeof(SahBin)) || !s.pos_node.reserve((size_t)n * 4) || !s.pos_node_tmp.reserve((size_t)n * 4) ||
        !s.active.reserve((max_slots * 2 + 4) * 4 * 2) || !s.cursors.reserve(max_slots * 2 * 4 + max_slots * sizeof(SahSplit) + 64))
        return 1;
    FusedBuild B;
    B.n = (int)n; B.leaf_max = leaf_max; B.max_slots = (int)max_slots;
    B.plo = s.prim_lo.as<float4>(); B.phi = s.prim_hi.as<float4>();
    B.bounds = s.bounds.as<int>();
    B.nlo = s.n2_lo.as<float4>(); B.nhi = s.n2_hi.as<float4>();
    B.left = s.n2_left.as<int>(); B.right = s.n2_right.as<int>(); B.first = s.n2_first.as<int>(); B.count = s.n2_count.as<int>();
    B.slot_of = s.n2_flag.as<int>();
    B.ord_a = s.order.as<uint32_t>(); B.ord_b = s.order_tmp.as<uint32_t>();
    B.pn_a = s.pos_node.as<int>(); B.pn_b = s.pos_node_tmp.as<int>();
    B.act_a = s.active.as<int>(); B.act_b = B.act_a + (max_slots * 2 + 4);
    B.cursors = s.cursors.as<int>();
    B.splits = reinterpret_cast<SahSplit*>(B.cursors + max_slots * 2 + 16);
    B.counters = s.counters.as<int>();
    B.bins = s.bins.as<SahBin>();
    B.qa = s.queue_a.as<int2>(); B.qb = s.queue_b.as<int2>();
    B.remap = remap; B.out = out; B.out_lo = out_lo; B.out_hi = out_hi;
    B.w8 = w8;
    // one item per thread up to two blocks per SM (all co-resident: 64 registers x 256 threads); a single block for tiny inputs
    const int blocks = (int)std::min<size_t>((size_t)c.sm_count * 2, ((size_t)n + 255) / 256);
    // One block with __syncthreads() between the phases instead of the cooperative grid was measured for RTIO's 1 936 instances: 1.20 ms
    // against 0.73 ms (the phases are bound by the latency of their global atomics, which one SM cannot overlap) - so only tiny inputs
    // take the single block (GKCU_FUSED_SINGLE_MAX to re-measure, profiles/r02_experiments.md)
    static const uint32_t single_max = getenv("GKCU_FUSED_SINGLE_MAX") ? (uint32_t)atoi(getenv("GKCU_FUSED_SINGLE_MAX")) : 0u;
    if (blocks <= 1 || n <= single_max) {
        k_build_fused<false><<<1, n <= 256 ? 256 : 1024, 0, st>>>(B);
    } else {
        void* args[] = {&B};
        GK_CUDA(cudaLaunchCooperativeKernel((void*)k_build_fused<true>, dim3(blocks), dim3(256), args, 0, st));
    }
    c.launches += 1;
    int table[48];
    int last[5];
    if (publish_words(c, B.counters + 8, 8, 56, st)) return 1;  // counters[8..12] = totals, counters[16..63] = level table
    GK_CUDA(cudaStreamSynchronize(st));
    GK_CUDA(cudaGetLastError());
    memcpy(last, c.pub_host + 8, 20);
    memcpy(table, c.pub_host + 16, 48 * 4);
    if (last[2] != 0) { set_error("BVH collapse did not terminate (tree deeper than 46 wide levels)"); return 1; }
    const int levels_run = last[3];
    info.leaf_count = (uint32_t)last[4];
    info.level_start.clear();
    info.level_start.push_back(0);
    table[0] = 0;
    table[1] = 1;
    for (int L = 1; L <= levels_run + 1 && L < 48; ++L) {
        if (L > 1 && (uint32_t)table[L] == info.level_start.back()) break;
        info.level_start.push_back((uint32_t)table[L]);
    }
    if (info.level_start.back() != (uint32_t)last[0]) { set_error("BVH collapse: inconsistent level table"); return 1; }
    info.node_count = (uint32_t)last[0];
    return 0;
}

}  // namespace

namespace {
// gkcu_update_nodes_sparse: 13 lanes move the 13 x 16 bytes of one NodeProxy
__global__ void k_scatter_nodes(const uint4* __restrict__ stage, const uint32_t* __restrict__ idx, uint32_t count, uint4* __restrict__ nodes) {
    const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) / 16, part = threadIdx.x & 15;
    if (k >= count || part >= sizeof(GkNodeProxy) / 16) return;
    nodes[(size_t)idx[k] * (sizeof(GkNodeProxy) / 16) + part] = stage[(size_t)k * (sizeof(GkNodeProxy) / 16) + part];
}
}  // namespace

int scatter_nodes(Context& c, uint32_t count) {
    static_assert(sizeof(GkNodeProxy) % 16 == 0, "NodeProxy is moved in 16-byte pieces");
    k_scatter_nodes<<<(unsigned)(((size_t)count * 16 + 255) / 256), 256, 0, c.stream>>>(c.node_stage.as<uint4>(), c.node_stage_idx.as<uint32_t>(), count,
                                                                                   c.nodes.as<uint4>());
    c.launches++;
    GK_CUDA(cudaGetLastError());
    return 0;
}

gkd::TraceScene make_trace_scene(const Context& c) {
    gkd::TraceScene s;
    s.tlas = c.tlas_nodes.as<Node4>();
    s.blas = c.blas_nodes.as<Node4>();
    s.tris = c.blas_tris.as<float4>();
    s.inst = c.instances.as<InstanceRec>();
    s.tlas_leaf = c.tlas_leaf.as<uint32_t>();
    s.blas_leaf = c.blas_leaf.as<uint32_t>();
    s.tlas_nodes = c.tlas.node_count;
    s.fmt = (uint32_t)c.node_format;
    return s;
}

int build_all_blas(Context& c) {
    cudaStream_t st = c.stream;
    const char* env = getenv("GKCU_BLAS_BUILDER");
    if (env) c.blas_builder = (std::string(env) == "lbvh") ? 0 : 1;
    if (const char* e = getenv("GKCU_LEAF_MAX")) kBlasLeafMax = std::max(1, std::min(8, atoi(e)));
    c.blas.assign(c.model_count, WideBvhInfo());
    size_t total_prims = 0;
    for (uint32_t m = 0; m < c.model_count; ++m) {
        uint32_t idx = 0;
        for (int s = 0; s < 10; ++s) idx += c.host_offsets[m * 10 + s].indexCount;
        c.blas[m].prim_count = idx / 3;
        c.blas[m].vertex_offset = c.host_offsets[m * 10].vertexOffset;
        c.blas[m].index_offset = c.host_offsets[m * 10].indexOffset;
        total_prims += idx / 3;
    }
    size_t node_cap = total_prims + 2 * (size_t)c.model_count + 4;
    if (!c.blas_nodes.reserve(node_cap * sizeof(Node4)) || !c.blas_node_lo.reserve(node_cap * 16) || !c.blas_node_hi.reserve(node_cap * 16) ||
        !c.blas_tris.reserve((total_prims + 1) * 48) || !c.blas_tri_prim.reserve((total_prims + 1) * 4) ||
        !c.blas_table.reserve((size_t)(c.model_count + 1) * 48) || !c.shade_tris.reserve((total_prims + 1) * 32) ||
        !c.blas_leaf.reserve((total_prims + (size_t)(c.model_count + 1) * GK_LEAF_SLACK) * 4))
        return 1;
    cudaEventRecord(c.ev[6], st);
    uint32_t node_base = 0, tri_base = 0, leaf_base = 0;
    std::vector<float> table((size_t)c.model_count * 12, 0.0f);
    for (uint32_t m = 0; m < c.model_count; ++m) {
        WideBvhInfo& bi = c.blas[m];
        bi.node_base = node_base;
        bi.tri_base = tri_base;
        uint32_t n = bi.prim_count;
        bi.leaf_base = leaf_base;
        bi.leaf_count = 0;
        uint32_t* tb = reinterpret_cast<uint32_t*>(&table[(size_t)m * 12]);
        // Node8 instances carry the BLAS's leaf-table base where Node4 instances carry its triangle base (InstanceRec::tri_base)
        tb[0] = node_base; tb[1] = c.node_format == 8 ? leaf_base : tri_base; tb[2] = n; tb[3] = tri_base;  // shade triangles use the per-model triangle base, primitive order
        if (n == 0) continue;
        if (n >= (1u << 28)) { set_error("BLAS with >= 2^28 triangles"); return 1; }
        BuildScratch& s = c.scratch;
        if (!s.prim_lo.reserve((size_t)n * 16) || !s.prim_hi.reserve((size_t)n * 16) || !reserve_bvh2(s, n)) return 1;
        k_tri_boxes<<<grid_for(n), 256, 0, st>>>(c.simple.as<uint16_t>(), c.indices.as<uint32_t>(), bi.vertex_offset, bi.index_offset, n,
                                                 s.prim_lo.as<float4>(), s.prim_hi.as<float4>());
        c.launches += 1;
        Bvh2 b2;
        if (c.blas_builder == 1) { if (build_sah(c, n, kBlasLeafMax, b2)) return 1; }
        else { if (build_lbvh(c, n, n, b2)) return 1; }
        Node4* out = c.blas_nodes.as<Node4>() + node_base;
        Wide8Out w8{c.node_format, nullptr, c.blas_leaf.as<uint32_t>() + leaf_base, tri_base};
        if (collapse(c, b2, n, kBlasLeafMax, nullptr, out, c.blas_node_lo.as<float4>() + node_base, c.blas_node_hi.as<float4>() + node_base, bi, w8)) return 1;
        k_write_tris<<<grid_for(n), 256, 0, st>>>(c.simple.as<uint16_t>(), c.indices.as<uint32_t>(), bi.vertex_offset, bi.index_offset, b2.order, n,
                                                  c.blas_tris.as<float4>() + (size_t)tri_base * 3, c.blas_tri_prim.as<uint32_t>() + tri_base);
        k_shade_tris<<<grid_for(n), 256, 0, st>>>(c.vertices.as<uint2>(), c.indices.as<uint32_t>(), bi.vertex_offset, bi.index_offset, n,
                                                  c.shade_tris.as<uint4>() + (size_t)tri_base * 2);
        c.launches += 2;
        float4 rl, rh;
        GK_CUDA(cudaMemcpyAsync(&rl, c.blas_node_lo.as<float4>() + node_base, 16, cudaMemcpyDeviceToHost, st));
        GK_CUDA(cudaMemcpyAsync(&rh, c.blas_node_hi.as<float4>() + node_base, 16, cudaMemcpyDeviceToHost, st));
        GK_CUDA(cudaStreamSynchronize(st));
        bi.root_lo[0] = rl.x; bi.root_lo[1] = rl.y; bi.root_lo[2] = rl.z;
        bi.root_hi[0] = rh.x; bi.root_hi[1] = rh.y; bi.root_hi[2] = rh.z;
        float* tf = &table[(size_t)m * 12];
        tf[4] = rl.x; tf[5] = rl.y; tf[6] = rl.z; tf[7] = 0;
        tf[8] = rh.x; tf[9] = rh.y; tf[10] = rh.z; tf[11] = 0;
        node_base += bi.node_count;
        tri_base += n;
        leaf_base += c.node_format == 8 ? bi.leaf_count : 0u;
    }
    GK_CUDA(cudaMemcpyAsync(c.blas_table.p, table.data(), table.size() * 4, cudaMemcpyHostToDevice, st));
    cudaEventRecord(c.ev[7], st);
    GK_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&c.timings.blas_build_ms, c.ev[6], c.ev[7]);
    c.blas_built = true;
    c.tlas_dirty = true;
    return 0;
}

int refit_blas(Context& c, uint32_t model) {
    if (!c.blas_built || model >= c.model_count) { set_error("gkcu_blas_refit: no such BLAS"); return 1; }
    cudaStream_t st = c.stream;
    WideBvhInfo& bi = c.blas[model];
    if (bi.prim_count == 0) return 0;
    cudaEventRecord(c.ev[6], st);
    k_write_tris<<<grid_for(bi.prim_count), 256, 0, st>>>(c.simple.as<uint16_t>(), c.indices.as<uint32_t>(), bi.vertex_offset, bi.index_offset, nullptr,
                                                          bi.prim_count, c.blas_tris.as<float4>() + (size_t)bi.tri_base * 3,
                                                          c.blas_tri_prim.as<uint32_t>() + bi.tri_base);
    c.launches += 1;
    Node4* nodes = c.blas_nodes.as<Node4>() + bi.node_base;
    float4 *nlo = c.blas_node_lo.as<float4>() + bi.node_base, *nhi = c.blas_node_hi.as<float4>() + bi.node_base;
    for (int L = (int)bi.level_start.size() - 2; L >= 0; --L) {
        uint32_t b = bi.level_start[L], e = bi.level_start[L + 1];
        if (e <= b) continue;
        if (c.node_format == 8)
            k_refit_level8<<<grid_for(e - b, 128), 128, 0, st>>>(reinterpret_cast<Node8*>(nodes), nlo, nhi, b, e, c.blas_tris.as<float4>(),
                                                                 c.blas_leaf.as<uint32_t>() + bi.leaf_base);
        else
            k_refit_level<<<grid_for(e - b, 128), 128, 0, st>>>(nodes, nlo, nhi, b, e, c.blas_tris.as<float4>() + (size_t)bi.tri_base * 3);
        c.launches += 1;
    }
    // the root box feeds the instance boxes of the next TLAS build
    GK_CUDA(cudaMemcpyAsync(reinterpret_cast<float4*>(c.blas_table.p) + model * 3 + 1, nlo, 16, cudaMemcpyDeviceToDevice, st));
    GK_CUDA(cudaMemcpyAsync(reinterpret_cast<float4*>(c.blas_table.p) + model * 3 + 2, nhi, 16, cudaMemcpyDeviceToDevice, st));
    cudaEventRecord(c.ev[7], st);
    GK_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&c.timings.blas_refit_ms, c.ev[6], c.ev[7]);
    c.tlas_dirty = true;
    return 0;
}

static int build_tlas_impl(Context& c);
int build_tlas(Context& c) {
    const int r = build_tlas_impl(c);
    if (r) {  // a failed build leaves no TLAS behind and stays pending: the next frame fails again instead of rendering an empty scene
        c.tlas.node_count = 0;
        c.tlas_dirty = true;
    }
    return r;
}
static int build_tlas_impl(Context& c) {
    cudaStream_t st = c.stream;
    uint32_t n = c.node_count;
    const uint32_t kept_nodes = c.tlas.node_count;
    c.tlas.node_count = 0;
    c.tlas_dirty = false;
    if (n == 0 || !c.blas_built) return 0;
    if (n >= (1u << 28)) { set_error("too many instances"); return 1; }
    BuildScratch& s = c.scratch;
    if (!c.instances.reserve((size_t)n * sizeof(InstanceRec)) || !s.prim_lo.reserve((size_t)n * 16) || !s.prim_hi.reserve((size_t)n * 16) ||
        !reserve_bvh2(s, n) || !c.tlas_nodes.reserve(((size_t)n + 4) * sizeof(Node4)) || !c.tlas_node_lo.reserve(((size_t)n + 4) * 16) ||
        !c.tlas_node_hi.reserve(((size_t)n + 4) * 16))
        return 1;
    if (!c.inst_lo.reserve((size_t)n * 16) || !c.inst_hi.reserve((size_t)n * 16) || !c.shade_inst.reserve((size_t)n * 32)) return 1;
    if (!c.inst_flags.reserve(n) || !c.tlas_active_ids.reserve((size_t)n * 4) || !c.select_count.reserve(16) ||
        !c.tlas_leaf.reserve(((size_t)n + GK_LEAF_SLACK) * 4))
        return 1;
    const Wide8Out w8{c.node_format, nullptr, c.tlas_leaf.as<uint32_t>(), 0u};
    k_instances<<<grid_for(n, 128), 128, 0, st>>>(c.nodes.as<GkNodeProxy>(), n, c.blas_table.as<uint4>(), c.model_count, c.instances.as<InstanceRec>(),
                                             c.inst_lo.as<float4>(), c.inst_hi.as<float4>(), c.shade_inst.as<uint4>(), c.inst_flags.as<uint8_t>());
    c.launches += 1;
    const char* tb = getenv("GKCU_TLAS_BUILDER");
    const bool lbvh = tb && std::string(tb) == "lbvh";
    if (!c.inst_flags_prev.reserve(n)) return 1;
    if (c.tlas_refit && c.node_format == 4 && !lbvh && c.tlas_topo_valid && c.tlas_topo_n == n && kept_nodes > 0 && c.tlas.level_start.size() >= 2 &&
        c.tlas.level_start.size() <= 49) {
        // Refit first and ask afterwards: the refit and the question "is the topology still the right one" go to the device together and the
        // host waits once. When the answer is no, the refit has only touched a tree that is rebuilt below.
        uint32_t* words = c.select_count.as<uint32_t>();
        GK_CUDA(cudaMemsetAsync(words + 1, 0, 8, st));
        k_flags_changed<<<grid_for(n), 256, 0, st>>>(c.inst_flags.as<uint8_t>(), c.inst_flags_prev.as<uint8_t>(), n, words + 1);
        TlasRefitArgs A;
        A.nodes = c.tlas_nodes.as<Node4>(); A.node_lo = c.tlas_node_lo.as<float4>(); A.node_hi = c.tlas_node_hi.as<float4>();
        A.inst_lo = c.inst_lo.as<float4>(); A.inst_hi = c.inst_hi.as<float4>();
        A.cost = reinterpret_cast<float*>(words + 2);
        A.levels = (int)c.tlas.level_start.size() - 1;
        for (size_t k = 0; k < c.tlas.level_start.size(); ++k) A.level_start[k] = c.tlas.level_start[k];
        k_tlas_refit<<<1, 1024, 0, st>>>(A);
        c.launches += 2;
        if (publish_words(c, words + 1, 1, 2, st)) return 1;
        GK_CUDA(cudaStreamSynchronize(st));
        GK_CUDA(cudaGetLastError());
        const bool changed = c.pub_host[1] != 0;
        float cost = 0.0f;
        memcpy(&cost, c.pub_host + 2, 4);
        if (c.tlas_built_cost < 0.0f) c.tlas_built_cost = cost;  // first refit of a fresh tree: its boxes are the builder's
        const bool degraded = !(cost <= c.tlas_built_cost * kRefitCostLimit);
        const bool periodic = c.tlas_rebuild_every > 0 && c.tlas_frames_since_build + 1 >= (uint32_t)c.tlas_rebuild_every;
        if (!changed && !degraded && !periodic) {
            c.tlas.node_count = kept_nodes;
            c.tlas_frames_since_build++;
            c.tlas_refits++;
            return 0;
        }
    }
    c.tlas_topo_valid = false;
    {   // the instances with a non-zero mask, in index order (stable compaction), and how many there are
        thrust::counting_iterator<uint32_t> iota(0u);
        size_t tmp = 0;
        cub::DeviceSelect::Flagged(nullptr, tmp, iota, c.inst_flags.as<uint8_t>(), c.tlas_active_ids.as<uint32_t>(), c.select_count.as<uint32_t>(), (int)n, st);
        if (!c.select_tmp.reserve(tmp)) return 1;
        GK_CUDA(cub::DeviceSelect::Flagged(c.select_tmp.p, tmp, iota, c.inst_flags.as<uint8_t>(), c.tlas_active_ids.as<uint32_t>(), c.select_count.as<uint32_t>(),
                                           (int)n, st));
        if (publish_words(c, c.select_count.p, 0, 1, st)) return 1;
        GK_CUDA(cudaStreamSynchronize(st));
        const uint32_t m_host = c.pub_host[0];
        c.tlas.prim_count = m_host;
        c.launches += 1;
    }
    // only instances with a non-zero mask enter the TLAS (RayTraceBaseRenderer.cpp:178); their indices were listed by
    // gkcu_update_nodes. The TLAS is built with the SAH-binned builder by default: on RTIO.proc it halves the traversal time
    // compared with the Morton LBVH (the floor's 2 km box degrades the Morton order); GKCU_TLAS_BUILDER=lbvh selects the LBVH.
    const uint32_t m = c.tlas.prim_count;
    if (m == 0) return 0;
    k_gather_boxes<<<grid_for(m), 256, 0, st>>>(c.tlas_active_ids.as<uint32_t>(), m, c.inst_lo.as<float4>(), c.inst_hi.as<float4>(),
                                               s.prim_lo.as<float4>(), s.prim_hi.as<float4>());
    c.launches += 1;
    Bvh2 b2;
    const uint32_t fused_max = getenv("GKCU_TLAS_FUSED_MAX") ? (uint32_t)atoi(getenv("GKCU_TLAS_FUSED_MAX")) : 262144u;
    if (lbvh) {
        if (build_lbvh(c, m, m, b2)) return 1;
        if (collapse(c, b2, m, 1, c.tlas_active_ids.as<uint32_t>(), c.tlas_nodes.as<Node4>(), c.tlas_node_lo.as<float4>(), c.tlas_node_hi.as<float4>(), c.tlas, w8)) return 1;
    } else if (m <= fused_max) {
        if (build_fused(c, m, 1, c.tlas_active_ids.as<uint32_t>(), c.tlas_nodes.as<Node4>(), c.tlas_node_lo.as<float4>(), c.tlas_node_hi.as<float4>(), c.tlas, w8)) return 1;
    } else {
        if (build_sah(c, m, 1, b2)) return 1;
        if (collapse(c, b2, m, 1, c.tlas_active_ids.as<uint32_t>(), c.tlas_nodes.as<Node4>(), c.tlas_node_lo.as<float4>(), c.tlas_node_hi.as<float4>(), c.tlas, w8)) return 1;
    }
    // the traversal stack holds at most 3 (Node4: one per sibling) or 2 (Node8: one group of nodes, one of leaves) entries per
    // level (TLAS levels + the deepest BLAS)
    size_t blas_levels = 0;
    for (auto& b : c.blas) blas_levels = std::max(blas_levels, b.level_start.size() > 0 ? b.level_start.size() - 1 : 0);
    const size_t per_level = c.node_format == 8 ? 2 : 3;
    size_t need = per_level * (c.tlas.level_start.size() - 1) + per_level * blas_levels + 2;
    if (need > GK_TRAV_STACK) {
        char msg[160];
        snprintf(msg, sizeof(msg), "acceleration structure too deep for the traversal stack (%zu TLAS + %zu BLAS levels need %zu entries, have %d)",
                 c.tlas.level_start.size() - 1, blas_levels, need, GK_TRAV_STACK);
        set_error(msg);
        c.tlas.node_count = 0;
        return 1;
    }
    // this topology serves the following frames for as long as the same instances are active (refit above)
    GK_CUDA(cudaMemcpyAsync(c.inst_flags_prev.p, c.inst_flags.p, n, cudaMemcpyDeviceToDevice, st));
    c.tlas_topo_valid = true;
    c.tlas_topo_n = n;
    c.tlas_built_cost = -1.0f;
    c.tlas_frames_since_build = 0;
    c.tlas_builds++;
    return 0;
}

}  // namespace gkcu
