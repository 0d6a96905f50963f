// gk_bvh.cuh — acceleration-structure layouts and the stack traversal used by every tracing kernel.
// Replaces what VK_KHR_ray_query + RT cores do for the reference (RayQuery.TraceRayInline/Proceed at
// assets/shaders/common/Shading.slang:749-756, 794-801); B200 has no RT cores, so this is the hot loop.
//
// Layout (all 16-byte aligned, read with 128-bit loads). Two node formats, selected per context (GKCU_NODE_FORMAT):
//   Node8      128 B   8-wide QUANTISED node: child boxes as 11-bit integers on a per-node power-of-two grid,
//                      stored as fp16 bit patterns in plane-major order, so decoding a plane is one f16->f32 conversion on
//                      the FMA pipe and one FMA; children are addressed as base + slot, a traversal-stack entry is a
//                      (base, slot mask) group
//   Node4      128 B   4-wide node, fp32 child boxes in plane-major order (default: the kernel is bound by instruction issue and
//                      the 8-wide step costs twice the instructions for 23 % fewer steps, profiles/r02_experiments.md)
//   leaf tris   48 B   3 x float4 (v0|prim id, v1, v2), stored in leaf order
//   InstanceRec 64 B   world->object 3x4 + BLAS offsets; TLAS leaves reference one instance each
// Semantics (shared with the CPU checker): a hit needs tmin < t < tmax, no face culling; the committed closest hit
// is the minimum of (t, instance, primitive) over all such hits, which makes the result independent of traversal order.
#pragma once
#include "gk_device.cuh"

namespace gkd {

struct __align__(16) Node4 {
    float lo_x[4], hi_x[4];  // child boxes, one float4 per plane set; an empty slot has lo = +inf, hi = -inf
    float lo_y[4], hi_y[4];
    float lo_z[4], hi_z[4];
    uint32_t child[4];       // CHILD_EMPTY | leaf (bit31, count-1 in 30:28, first in 27:0) | node index
    uint32_t n, pad[3];      // child count
};
static_assert(sizeof(Node4) == 128, "Node4 must be 128 bytes (one cache line)");

// 8-wide quantised node. Child c's box on axis a is [o_a + qlo_a[c] * s_a, o_a + qhi_a[c] * s_a] with s_a a power of two and
// the q's integers in [0, 2048] (exact in fp16), rounded outward by the builder. Children are stored internal nodes first:
// slots [0, ni) are the nodes child_base + slot, slots [ni, n) are the leaf-table entries lb + slot (lb = first leaf - ni, kept
// >= 0 by eight slack entries in front of every leaf table). An empty slot has qlo = +inf, qhi = -inf and can never be hit.
// A leaf-table entry is a triangle range (first | (count - 1) << 28, absolute in the triangle pool) in a BLAS and an instance
// index in the TLAS.
struct __align__(16) Node8 {
    float ox, oy, oz;
    uint32_t link;       // child_base (bits 0..23) | n << 24 | ni << 28
    float sx, sy, sz;
    uint32_t lb;
    uint16_t q[6][8];    // qlo_x, qhi_x, qlo_y, qhi_y, qlo_z, qhi_z
};
static_assert(sizeof(Node8) == 128, "Node8 must be 128 bytes (one cache line)");
#ifndef GK_NODE_LD256
#define GK_NODE_LD256 0  // Node4 node step: 256-bit plane loads (experiment, profiles/r02_experiments.md)
#endif
#ifndef GK_N8_TMIN
#define GK_N8_TMIN 1  // Node8 stack entries carry the entry distance of their nearest member (24 bits) and are culled as a whole when popped
#endif
#ifndef GK_FOLD_PAD
#define GK_FOLD_PAD 0  // Node4 node step: the relative / absolute padding of the slab interval is folded into the per-ray constants (8 FMAs fewer per step)
#endif
#define GK_Q_MAX 2047.0f
#define GK_LEAF_SLACK 8u

struct __align__(16) InstanceRec {
    float w2o[12];       // world->object, 3 rows x 4
    uint32_t node_base;  // BLAS root index in the BLAS node pool
    uint32_t tri_base;   // Node4: first leaf-ordered triangle of the BLAS in the triangle pool; Node8: first entry of the BLAS's leaf table
    uint32_t active;     // visible && !nort && BLAS non-empty (VkAccelerationStructureInstanceKHR.mask)
    uint32_t blas;
};
static_assert(sizeof(InstanceRec) == 64, "InstanceRec must be 64 bytes");

#define GK_CHILD_EMPTY 0xFFFFFFFFu
#define GK_LEAF_BIT 0x80000000u
#define GK_TRAV_STACK 64
#define GK_NO_HIT 0xFFFFFFFFu
#ifndef GK_PUSH_ORDERED
#define GK_PUSH_ORDERED 1
#endif
#ifndef GK_NODE_BF
#define GK_NODE_BF 0  // Node4 node step: 1 = branch-free ordering and pushes, measured 5 % slower than the compare-and-push ladder (1.09 children are hit per node step on the bench frame, so the ladder mostly takes its short paths; profiles/r02_experiments.md)
#endif

struct TraceScene {
    const Node4* tlas;    // Node4 or Node8 array (both 128 B per node), see fmt
    const Node4* blas;
    const float4* tris;
    const InstanceRec* inst;
    const uint32_t* tlas_leaf;  // Node8: leaf tables
    const uint32_t* blas_leaf;
    uint32_t tlas_nodes;  // 0 -> empty scene
    uint32_t fmt;         // 4 or 8
};

struct Hit {
    float t, u, v;
    uint32_t inst, prim;
};

GK_DEV float safe_rcp(float x) {
    if (fabsf(x) < 1e-18f) return x < 0.0f ? -1e18f : 1e18f;
    return __frcp_rn(x);  // correctly rounded, == 1.0f / x
}
GK_DEV float sel3(int k, float3 v) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
// Node load: read-only path; GK_NODE_EVICT_LAST = 1 asks L1 to keep node lines over the streamed data (experiment)
#ifndef GK_NODE_EVICT_LAST
#define GK_NODE_EVICT_LAST 0
#endif
GK_DEV uint4 ld_node(const uint4* p) {
#if GK_NODE_EVICT_LAST
    uint4 v;
    asm("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
// 256-bit read-only global load (sm_100: LDG.E.ENL2.256), p must be 32-byte aligned
GK_DEV void ld256(const void* p, float (&v)[8]) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}

struct RayShear {
    int kx, ky, kz;
    float Sx, Sy, Sz;
};
GK_DEV RayShear make_shear(float3 d) {
    RayShear r;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    r.kz = (ax > ay) ? ((ax > az) ? 0 : 2) : ((ay > az) ? 1 : 2);
    r.kx = r.kz + 1; if (r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if (r.ky == 3) r.ky = 0;
    float dz = sel3(r.kz, d);
    if (dz < 0.0f) { int t = r.kx; r.kx = r.ky; r.ky = t; }
    r.Sx = sel3(r.kx, d) / dz;
    r.Sy = sel3(r.ky, d) / dz;
    r.Sz = 1.0f / dz;
    return r;
}

// Watertight ray/triangle test (Woop, Benthin, Wald 2013), no culling; returns t and the weights of v1, v2.
GK_DEV bool tri_test(float3 v0, float3 v1, float3 v2, float3 o, const RayShear& s, float& t, float& b1, float& b2) {
    const float3 A = v0 - o, B = v1 - o, C = v2 - o;
    const float Akz = sel3(s.kz, A), Bkz = sel3(s.kz, B), Ckz = sel3(s.kz, C);
    const float Ax = sel3(s.kx, A) - s.Sx * Akz, Ay = sel3(s.ky, A) - s.Sy * Akz;
    const float Bx = sel3(s.kx, B) - s.Sx * Bkz, By = sel3(s.ky, B) - s.Sy * Bkz;
    const float Cx = sel3(s.kx, C) - s.Sx * Ckz, Cy = sel3(s.ky, C) - s.Sy * Ckz;
    float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)__dsub_rn(__dmul_rn((double)Cx, (double)By), __dmul_rn((double)Cy, (double)Bx));
        V = (float)__dsub_rn(__dmul_rn((double)Ax, (double)Cy), __dmul_rn((double)Ay, (double)Cx));
        W = (float)__dsub_rn(__dmul_rn((double)Bx, (double)Ay), __dmul_rn((double)By, (double)Ax));
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    const float det = U + V + W;
    if (det == 0.0f) return false;
    const float Az = s.Sz * Akz, Bz = s.Sz * Bkz, Cz = s.Sz * Ckz;
    const float T = U * Az + V * Bz + W * Cz;
    const float inv = 1.0f / det;
    t = T * inv;
    b1 = V * inv;
    b2 = W * inv;
    return true;
}

GK_DEV float3 xform_point(const float* m, float3 p) {
    return f3(m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3], m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
              m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]);
}
GK_DEV float3 xform_vec(const float* m, float3 p) {
    return f3(m[0] * p.x + m[1] * p.y + m[2] * p.z, m[4] * p.x + m[5] * p.y + m[6] * p.z, m[8] * p.x + m[9] * p.y + m[10] * p.z);
}

// Where a traversal keeps its world-space ray while it is inside an instance: registers for the simple per-ray loop,
// shared memory for k_trace (nine registers fewer per lane; the values are only touched when an instance is entered or left).
struct WorldRegs {
    static constexpr bool kKeepsTlasSpace = false;
    float3 o_, d_, widir_;
    GK_DEV void set(float3 o, float3 d, float3 widir) { o_ = o; d_ = d; widir_ = widir; }
    GK_DEV float3 o() const { return o_; }
    GK_DEV float3 d() const { return d_; }
    GK_DEV float3 widir() const { return widir_; }
};
template <int STRIDE>
struct WorldSmem {
    static constexpr bool kKeepsTlasSpace = false;
    float* p;  // p[k * STRIDE], k = 0..8, one column per thread
    GK_DEV void set(float3 o, float3 d, float3 widir) {
        p[0] = o.x; p[STRIDE] = o.y; p[2 * STRIDE] = o.z;
        p[3 * STRIDE] = d.x; p[4 * STRIDE] = d.y; p[5 * STRIDE] = d.z;
        p[6 * STRIDE] = widir.x; p[7 * STRIDE] = widir.y; p[8 * STRIDE] = widir.z;
    }
    GK_DEV float3 o() const { return f3(p[0], p[STRIDE], p[2 * STRIDE]); }
    GK_DEV float3 d() const { return f3(p[3 * STRIDE], p[4 * STRIDE], p[5 * STRIDE]); }
    GK_DEV float3 widir() const { return f3(p[6 * STRIDE], p[7 * STRIDE], p[8 * STRIDE]); }
};
// k_trace<5, 4>: sixteen words per thread - the world-space ray and, next to it, the Node4 slab constants of the TLAS space
// (padded offsets and the sign mask), so that leaving an instance is ten shared-memory loads instead of recomputing set_space
// with three of 32 lanes active (6 % of the kernel's warp instructions, profiles/r02_base_k_trace.txt).
#ifndef GK_EXIT_SMEM
#define GK_EXIT_SMEM 1
#endif
template <int STRIDE>
struct WorldSmemX {
    static constexpr bool kKeepsTlasSpace = GK_EXIT_SMEM != 0;
    float* p;  // p[k * STRIDE], k = 0..15
    GK_DEV void set(float3 o, float3 d, float3 widir) {
        p[0] = o.x; p[STRIDE] = o.y; p[2 * STRIDE] = o.z;
        p[3 * STRIDE] = d.x; p[4 * STRIDE] = d.y; p[5 * STRIDE] = d.z;
        p[6 * STRIDE] = widir.x; p[7 * STRIDE] = widir.y; p[8 * STRIDE] = widir.z;
    }
    GK_DEV float3 o() const { return f3(p[0], p[STRIDE], p[2 * STRIDE]); }
    GK_DEV float3 d() const { return f3(p[3 * STRIDE], p[4 * STRIDE], p[5 * STRIDE]); }
    GK_DEV float3 widir() const { return f3(p[6 * STRIDE], p[7 * STRIDE], p[8 * STRIDE]); }
    GK_DEV void keep(float3 adjn, float3 adjf, uint32_t sgn) {
        p[9 * STRIDE] = adjn.x; p[10 * STRIDE] = adjn.y; p[11 * STRIDE] = adjn.z;
        p[12 * STRIDE] = adjf.x; p[13 * STRIDE] = adjf.y; p[14 * STRIDE] = adjf.z;
        p[15 * STRIDE] = __uint_as_float(sgn);
    }
    GK_DEV float3 adjn() const { return f3(p[9 * STRIDE], p[10 * STRIDE], p[11 * STRIDE]); }
    GK_DEV float3 adjf() const { return f3(p[12 * STRIDE], p[13 * STRIDE], p[14 * STRIDE]); }
    GK_DEV uint32_t sgn() const { return __float_as_uint(p[15 * STRIDE]); }
};


// Traversal stack: a per-thread array in local memory, entries are (node reference, entry distance). Keeping the top levels
// in shared memory instead was measured and is slower (k_trace is sensitive to L1 capacity and the carve-out costs more than
// the faster pops gain, profiles/r01_experiments.md).
struct LocalStack {
    uint2 e[GK_TRAV_STACK];
    GK_DEV void push(int& sp, uint2 v) { e[sp++] = v; }
    GK_DEV uint2 pop(int& sp) { return e[--sp]; }
    GK_DEV void put(int at, uint2 v) { e[at] = v; }
};

// The same stack with its top entry in registers: a pop hands out the register pair and requests the entry below it, whose load
// latency is then hidden behind the node step that follows instead of sitting between the pop and its cull test
// (8.6 % of k_trace's stall samples, profiles/r02_base_k_trace.txt). Element sp - 1 lives in `top`, elements 0 .. sp - 2 in e[].
#ifndef GK_TOS_REG
#define GK_TOS_REG 0
#endif
struct TosStack {
    uint2 e[GK_TRAV_STACK];
    uint2 top;
    GK_DEV void push(int& sp, uint2 v) { if (sp > 0) e[sp - 1] = top; top = v; ++sp; }
    GK_DEV uint2 pop(int& sp) { const uint2 r = top; --sp; if (sp > 0) top = e[sp - 1]; return r; }
};

// ---------------------------------------------------------------------------------------------------------------------
// Traversal state machine. One ray per lane; the same three steps serve the simple per-ray loop (trace_ray) and the
// persistent-thread kernel (k_trace), which lets a warp run node steps together and leaf steps together and refills
// finished lanes with new rays. FMT selects the node format (4: Node4, 8: Node8).
struct Trav {
    float3 ro;         // current-space origin (world or object)
    float3 idir;       // current-space reciprocal direction
    float3 adjn, adjf; // Node4: slab offsets -ro * idir, lowered / raised by their own rounding error (see set_space)
    uint32_t sgn;      // bit a: idir[a] < 0, i.e. the near plane of axis a is the child's upper plane
#if GK_FOLD_PAD
    float3 idn, idf;   // Node4: idir scaled by the relative padding of the near / far side (see set_space)
    float tminp;       // tmin with the near-side padding applied
#endif
    float tmin, tmax;
    Hit best;
    RayShear sh;
    const uint4* nodes;
    const float4* tris;
    const uint32_t* leaf;  // Node8: leaf table of the current space
    uint32_t cur;      // current item: node index, or GK_LEAF_BIT | leaf reference
    uint32_t cur_inst; // GK_NO_HIT while in the TLAS
    int sp;
    int sp_inst;       // stack height when the current instance was entered: popping down to it leaves the instance
    bool any;          // first-hit query
    bool done;
#ifdef GK_TRAV_STATS
    uint32_t st[8];    // diagnostics build: TLAS node steps, BLAS node steps, children hit, pushes, culled pops, instance entries, leaf steps, triangle tests
#endif
};
#ifdef GK_TRAV_STATS
#define GK_STAT(T, k, v) ((T).st[k] += (v))
#else
#define GK_STAT(T, k, v) ((void)0)
#endif

// Node4: slab distances are evaluated as t = plane * idir + adj with adj = -(ro * idir): one FMA per plane. adj carries a rounding
// error of half an ulp (|adj| * 2^-24) which the cancellation in plane * idir + adj can make large relative to t, so the
// near-plane offset is lowered and the far-plane offset raised by 4 ulps of |adj|; the remaining error (one rounding of the
// FMA result) is relative to t and covered by the relative padding in node_test. The slabs therefore never reject a box
// that the exact ray enters, which is what makes the closest hit independent of the tree.
// Node8 derives its offsets per node (the quantisation origin enters them), so only ro / idir / sgn are kept.
#if GK_FOLD_PAD
GK_DEV void fold_idir(Trav& T) {
    T.idn = f3(T.idir.x * 0.999999f, T.idir.y * 0.999999f, T.idir.z * 0.999999f);
    T.idf = f3(T.idir.x * 1.000001f, T.idir.y * 1.000001f, T.idir.z * 1.000001f);
}
GK_DEV void fold_pad(Trav& T) {
    fold_idir(T);
    T.adjn = f3(__fmaf_rn(T.adjn.x, 0.999999f, -1e-7f), __fmaf_rn(T.adjn.y, 0.999999f, -1e-7f), __fmaf_rn(T.adjn.z, 0.999999f, -1e-7f));
    T.adjf = f3(__fmaf_rn(T.adjf.x, 1.000001f, 1e-7f), __fmaf_rn(T.adjf.y, 1.000001f, 1e-7f), __fmaf_rn(T.adjf.z, 1.000001f, 1e-7f));
}
#endif
template <int FMT>
GK_DEV void set_space(Trav& T, float3 ro, float3 idir) {
    T.ro = ro;
    T.idir = idir;
    if (FMT == 4) {
        const float ax = -(ro.x * idir.x), ay = -(ro.y * idir.y), az = -(ro.z * idir.z);
        // + 1e-14 world units (|idir| * 1e-14 in t): a ray lying in a face plane of a box (direction component 0, origin on the
        // plane) is inside it, like in the CPU checker's slab test
        const float kUlps = GK_FOLD_PAD ? 3.6e-7f : 2.4e-7f;
        const float ex = __fmaf_rn(fabsf(ax), kUlps, fabsf(idir.x) * 1e-14f), ey = __fmaf_rn(fabsf(ay), kUlps, fabsf(idir.y) * 1e-14f),
                    ez = __fmaf_rn(fabsf(az), kUlps, fabsf(idir.z) * 1e-14f);
        T.adjn = f3(ax - ex, ay - ey, az - ez);
        T.adjf = f3(ax + ex, ay + ey, az + ez);
#if GK_FOLD_PAD
        // The node step pads the entry distance n = max(plane * idir + adjn) to n * 0.999999 - 1e-7 and the exit distance likewise
        // upwards. Both are linear in n, so the factors move into the constants: plane * (idir * 0.999999) + (adjn * 0.999999 - 1e-7).
        // The two extra roundings (of the scaled idir and of the scaled offset) are each half an ulp of a term bounded by |adj| + |t|;
        // the 6-ulp lowering of |adj| above (3.6e-7 instead of the 4 ulps the unfolded form needs) and the 1e-6 relative padding cover them.
        fold_pad(T);
#endif
    }
    T.sgn = (idir.x < 0.0f ? 1u : 0u) | (idir.y < 0.0f ? 2u : 0u) | (idir.z < 0.0f ? 4u : 0u);
}

template <int FMT, class W>
GK_DEV void trav_init(Trav& T, W& world, const TraceScene& S, float3 o, float3 d, float tmin, float tmax, bool any) {
    const float3 idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
    world.set(o, d, idir);
    set_space<FMT>(T, o, idir);
    if constexpr (W::kKeepsTlasSpace && FMT == 4) world.keep(T.adjn, T.adjf, T.sgn);
    T.tmin = tmin; T.tmax = tmax;
#if GK_FOLD_PAD
    T.tminp = __fmaf_rn(tmin, 0.999999f, -1e-7f);
#endif
    T.best.t = tmax; T.best.u = 0.0f; T.best.v = 0.0f; T.best.inst = GK_NO_HIT; T.best.prim = 0;
    T.sh.kx = T.sh.ky = T.sh.kz = 0;
    T.sh.Sx = T.sh.Sy = T.sh.Sz = 0.0f;
    T.nodes = reinterpret_cast<const uint4*>(S.tlas);
    T.tris = S.tris;
    T.leaf = S.tlas_leaf;
    T.cur = 0;
    T.cur_inst = GK_NO_HIT;
    T.sp = 0;
    T.sp_inst = 0;
    T.any = any;
    T.done = S.tlas_nodes == 0;
#ifdef GK_TRAV_STATS
    for (int k = 0; k < 8; ++k) T.st[k] = 0;
#endif
}

// Pops the next item to visit; once every entry the current instance pushed has been consumed the ray goes back to world
// space and the TLAS. Node4 entries are (child reference, entry distance) and are culled here when they lie beyond the best
// hit; Node8 entries are (base, slot mask) groups that hand out one child per pop (a child beyond the best hit is rejected by
// its own node step or triangle test).
template <int FMT, class W, class St>
GK_DEV void trav_pop(Trav& T, const W& world, const TraceScene& S, St& stack) {
    for (;;) {
        if (T.cur_inst != GK_NO_HIT && T.sp == T.sp_inst) {
            if constexpr (W::kKeepsTlasSpace && FMT == 4) {  // T.ro is not read in world space (Node4 slabs use adjn / adjf, triangles live in object space)
                T.idir = world.widir(); T.adjn = world.adjn(); T.adjf = world.adjf(); T.sgn = world.sgn();
#if GK_FOLD_PAD
                fold_idir(T);  // the kept offsets are the folded ones
#endif
            } else {
                set_space<FMT>(T, world.o(), world.widir());
            }
            T.nodes = reinterpret_cast<const uint4*>(S.tlas);
            if (FMT == 8) T.leaf = S.tlas_leaf;
            T.cur_inst = GK_NO_HIT;
        }
        if (T.sp == 0) { T.done = true; return; }
        if constexpr (FMT == 8) {
            const uint2 e = stack.e[T.sp - 1];
            if (GK_N8_TMIN && __uint_as_float(e.y & ~0xFFu) > T.best.t) { GK_STAT(T, 4, 1); --T.sp; continue; }
            const uint32_t m = e.y & 0xFFu, rest = m & (m - 1u);
            if (rest) stack.e[T.sp - 1].y = (e.y & ~0xFFu) | rest;
            else --T.sp;
            T.cur = e.x + (uint32_t)(__ffs((int)m) - 1);
            return;
        } else {
            const uint2 e = stack.pop(T.sp);
            if (__uint_as_float(e.y) > T.best.t) { GK_STAT(T, 4, 1); continue; }
            T.cur = e.x;
            return;
        }
    }
}

// Internal node, Node4: test the 4 child boxes, continue with the nearest, push the others (second nearest on top).
// Node4 keeps the planes as fp32 in structure-of-arrays order, so the near and far planes of an axis are two 16-byte loads
// picked by the sign of the ray direction and every plane costs exactly one FMA. The interval is padded by a few ulps on
// both ends (see set_space).
// trav_node_core4 returns false when no child is hit (the caller pops); trav_node_step4 = core + pop.
template <class St>
GK_DEV bool trav_node_core4(Trav& T, St& stack) {
    const uint4* np = T.nodes + (size_t)T.cur * 8;
    const uint32_t ix = T.sgn & 1u, iy = (T.sgn >> 1) & 1u, iz = T.sgn >> 2;
#if GK_NODE_LD256
    // One 256-bit load per axis (sm_100: LDG.E.256) brings the lower and the upper planes of the four children; the ray's sign picks
    // which half is the near side. Four load instructions per node instead of seven: the L1 data pipe (one wavefront per lane and
    // load, since every lane reads its own node) is as busy as the issue slots in this kernel.
    float pnx[4], pfx[4], pny[4], pfy[4], pnz[4], pfz[4];
    {
        float a[8], b[8], c[8];
        ld256(np, a);
        ld256(np + 2, b);
        ld256(np + 4, c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            pnx[k] = ix ? a[4 + k] : a[k]; pfx[k] = ix ? a[k] : a[4 + k];
            pny[k] = iy ? b[4 + k] : b[k]; pfy[k] = iy ? b[k] : b[4 + k];
            pnz[k] = iz ? c[4 + k] : c[k]; pfz[k] = iz ? c[k] : c[4 + k];
        }
    }
    const uint4 n1 = __ldg(np + 6);
#else
    const uint4 nx = ld_node(np + ix), fx = ld_node(np + (ix ^ 1u));
    const uint4 ny = ld_node(np + 2 + iy), fy = ld_node(np + 2 + (iy ^ 1u));
    const uint4 nz = ld_node(np + 4 + iz), fz = ld_node(np + 4 + (iz ^ 1u));
    const uint4 n1 = ld_node(np + 6);
    const float pnx[4] = {__uint_as_float(nx.x), __uint_as_float(nx.y), __uint_as_float(nx.z), __uint_as_float(nx.w)};
    const float pfx[4] = {__uint_as_float(fx.x), __uint_as_float(fx.y), __uint_as_float(fx.z), __uint_as_float(fx.w)};
    const float pny[4] = {__uint_as_float(ny.x), __uint_as_float(ny.y), __uint_as_float(ny.z), __uint_as_float(ny.w)};
    const float pfy[4] = {__uint_as_float(fy.x), __uint_as_float(fy.y), __uint_as_float(fy.z), __uint_as_float(fy.w)};
    const float pnz[4] = {__uint_as_float(nz.x), __uint_as_float(nz.y), __uint_as_float(nz.z), __uint_as_float(nz.w)};
    const float pfz[4] = {__uint_as_float(fz.x), __uint_as_float(fz.y), __uint_as_float(fz.z), __uint_as_float(fz.w)};
#endif
    const float tfcap = __fmaf_rn(T.best.t, 1.000001f, 1e-7f);
    const float inf = __int_as_float(0x7f800000);
    float tn[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#if GK_FOLD_PAD
        const float x0 = __fmaf_rn(pnx[c], T.idn.x, T.adjn.x), x1 = __fmaf_rn(pfx[c], T.idf.x, T.adjf.x);
        const float y0 = __fmaf_rn(pny[c], T.idn.y, T.adjn.y), y1 = __fmaf_rn(pfy[c], T.idf.y, T.adjf.y);
        const float z0 = __fmaf_rn(pnz[c], T.idn.z, T.adjn.z), z1 = __fmaf_rn(pfz[c], T.idf.z, T.adjf.z);
        const float npad = fmaxf(fmaxf(x0, y0), fmaxf(z0, T.tminp));
        const float fpad = fminf(fminf(x1, y1), z1);
        const bool hit = (npad <= fpad) && (npad <= tfcap);  // an empty slot: npad = +inf, fpad = -inf
#else
        const float x0 = __fmaf_rn(pnx[c], T.idir.x, T.adjn.x), x1 = __fmaf_rn(pfx[c], T.idir.x, T.adjf.x);
        const float y0 = __fmaf_rn(pny[c], T.idir.y, T.adjn.y), y1 = __fmaf_rn(pfy[c], T.idir.y, T.adjf.y);
        const float z0 = __fmaf_rn(pnz[c], T.idir.z, T.adjn.z), z1 = __fmaf_rn(pfz[c], T.idir.z, T.adjf.z);
        const float n = fmaxf(fmaxf(x0, y0), fmaxf(z0, T.tmin));
        const float f = fminf(fminf(x1, y1), z1);
        const float npad = __fmaf_rn(n, 0.999999f, -1e-7f);
        // an empty slot has lo = +inf, hi = -inf: n = +inf, f = -inf, never a hit
        const bool hit = (npad <= __fmaf_rn(f, 1.000001f, 1e-7f)) && (npad <= tfcap);
#endif
        tn[c] = hit ? npad : inf;
    }
    const uint32_t ref[4] = {n1.x, n1.y, n1.z, n1.w};
    GK_STAT(T, T.cur_inst == GK_NO_HIT ? 0 : 1, 1);
    GK_STAT(T, 2, (tn[0] < inf) + (tn[1] < inf) + (tn[2] < inf) + (tn[3] < inf));
#if GK_NODE_BF
    {
        // Branch-free ordering: every hit child gets its rank among the hit children from six comparisons of distinct keys (entry
        // distance, slot index in the two lowest mantissa bits); rank 0 is visited next, rank r > 0 is stored at stack position
        // sp + nhit - 1 - r with a predicated store, i.e. the stack holds them far to near. All lanes of the node step run the same
        // ~45 instructions; the compare-and-push ladder this replaces ran its pushes with 4-5 of 32 lanes (24 % of the kernel's warp
        // instructions, profiles/r02_base_k_trace.txt). The stored entry distance is the key with the slot bits cleared: rounded
        // down, so the cull at pop time stays conservative.
        const uint32_t kInf = 0x7F800000u;
        uint32_t key[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
            key[c] = tn[c] < inf ? ((__float_as_uint(fmaxf(tn[c], 0.0f)) & ~3u) | (uint32_t)c) : (kInf | (uint32_t)c);
        const int p01 = key[0] < key[1], p02 = key[0] < key[2], p03 = key[0] < key[3], p12 = key[1] < key[2], p13 = key[1] < key[3], p23 = key[2] < key[3];
        const int r[4] = {3 - p01 - p02 - p03, p01 + 2 - p12 - p13, p02 + p12 + 1 - p23, p03 + p13 + p23};
        const int nhit = (key[0] < kInf) + (key[1] < kInf) + (key[2] < kInf) + (key[3] < kInf);
        if (nhit == 0) return false;
        const int top = T.sp + nhit - 1;
        uint32_t next = ref[3];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (key[c] < kInf && r[c] > 0) stack.put(top - r[c], make_uint2(ref[c], key[c] & ~3u));
            if (c < 3 && r[c] == 0) next = ref[c];
        }
        GK_STAT(T, 3, nhit - 1);
        T.sp = top;
        T.cur = next;
        return true;
    }
#endif
    // continue with the nearest child; the second nearest goes on top of the stack, the rest below it in slot order
    // (entries carry their entry distance, so a far one is culled when it is popped)
    // The child references travel with the distances through the two searches (selects): indexing ref[] with i0 / i1 would put the array
    // in local memory - a 16-byte store and two loads per node step and lane, a third of the kernel's L1 wavefronts, which are what
    // bounds it (l1tex data pipe 77 % busy, profiles/r02_experiments.md section 9).
    int i0 = 0;
    float t0 = tn[0];
    uint32_t r0 = ref[0];
    if (tn[1] < t0) { t0 = tn[1]; i0 = 1; r0 = ref[1]; }
    if (tn[2] < t0) { t0 = tn[2]; i0 = 2; r0 = ref[2]; }
    if (tn[3] < t0) { t0 = tn[3]; i0 = 3; r0 = ref[3]; }
    if (!(t0 < inf)) return false;
    if (GK_PUSH_ORDERED) {
        int i1 = -1;
        float t1 = inf;
        uint32_t r1 = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c != i0 && tn[c] < t1) { t1 = tn[c]; i1 = c; r1 = ref[c]; }
        if (i1 >= 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c != i0 && c != i1 && tn[c] < inf) { GK_STAT(T, 3, 1); stack.push(T.sp, make_uint2(ref[c], __float_as_uint(tn[c]))); }
            GK_STAT(T, 3, 1);
            stack.push(T.sp, make_uint2(r1, __float_as_uint(t1)));
        }
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c != i0 && tn[c] < inf) stack.push(T.sp, make_uint2(ref[c], __float_as_uint(tn[c])));
    }
    T.cur = r0;
    return true;
}
template <class W, class St>
GK_DEV void trav_node_step4(Trav& T, const W& world, const TraceScene& S, St& stack) {
    if (!trav_node_core4(T, stack)) trav_pop<4>(T, world, S, stack);
}

// Internal node, Node8. Per axis the plane distances are t = q * S + C with S = s * idir (exact: s is a power of two) and
// C = (o - ro) * idir; q * S is exact in the FMA, so the only rounding errors are the two roundings of C, the rounding of idir
// and the final rounding of the FMA, together at most 4 * 2^-24 * (|C| + 2048 |S|): the near offset is lowered and the far
// offset raised by P = 2.5e-7 |C| + 5e-4 |S| (5e-4 of one grid cell). The slabs therefore never reject a box the exact ray
// enters, which keeps the closest hit independent of the tree. The warp continues with the nearest hit child; the others stay
// behind as at most two stack entries (internal children, leaf children), each a (base, slot mask) group.
GK_DEV float2 gk_h2(uint32_t w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
#define GK_N8_SENT 0x7F000000u
template <class W, class St>
GK_DEV void trav_node_step8(Trav& T, const W& world, const TraceScene& S, St& stack) {
    const uint4* np = T.nodes + (size_t)T.cur * 8;
    const uint32_t ix = T.sgn & 1u, iy = (T.sgn >> 1) & 1u, iz = T.sgn >> 2;
    const uint4 h0 = __ldg(np), h1 = __ldg(np + 1);
    const uint4 nx = __ldg(np + 2 + ix), fx = __ldg(np + 2 + (ix ^ 1u));
    const uint4 ny = __ldg(np + 4 + iy), fy = __ldg(np + 4 + (iy ^ 1u));
    const uint4 nz = __ldg(np + 6 + iz), fz = __ldg(np + 6 + (iz ^ 1u));
    // per-axis scale and the two padded offsets
    const float Sx = __uint_as_float(h1.x) * T.idir.x, Sy = __uint_as_float(h1.y) * T.idir.y, Sz = __uint_as_float(h1.z) * T.idir.z;
    const float Cx = (__uint_as_float(h0.x) - T.ro.x) * T.idir.x, Cy = (__uint_as_float(h0.y) - T.ro.y) * T.idir.y, Cz = (__uint_as_float(h0.z) - T.ro.z) * T.idir.z;
    const float Px = __fmaf_rn(fabsf(Cx), 2.5e-7f, fabsf(Sx) * 5e-4f), Py = __fmaf_rn(fabsf(Cy), 2.5e-7f, fabsf(Sy) * 5e-4f),
                Pz = __fmaf_rn(fabsf(Cz), 2.5e-7f, fabsf(Sz) * 5e-4f);
    const float Cnx = Cx - Px, Cfx = Cx + Px, Cny = Cy - Py, Cfy = Cy + Py, Cnz = Cz - Pz, Cfz = Cz + Pz;
    const uint32_t wnx[4] = {nx.x, nx.y, nx.z, nx.w}, wfx[4] = {fx.x, fx.y, fx.z, fx.w};
    const uint32_t wny[4] = {ny.x, ny.y, ny.z, ny.w}, wfy[4] = {fy.x, fy.y, fy.z, fy.w};
    const uint32_t wnz[4] = {nz.x, nz.y, nz.z, nz.w}, wfz[4] = {fz.x, fz.y, fz.z, fz.w};
    const float tcap = T.best.t;
    uint32_t kmin = GK_N8_SENT | 7u, k2 = GK_N8_SENT | 7u, mask = 0u;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const float2 qnx = gk_h2(wnx[w]), qfx = gk_h2(wfx[w]), qny = gk_h2(wny[w]), qfy = gk_h2(wfy[w]), qnz = gk_h2(wnz[w]), qfz = gk_h2(wfz[w]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = 2 * w + h;
            const float x0 = __fmaf_rn(h ? qnx.y : qnx.x, Sx, Cnx), x1 = __fmaf_rn(h ? qfx.y : qfx.x, Sx, Cfx);
            const float y0 = __fmaf_rn(h ? qny.y : qny.x, Sy, Cny), y1 = __fmaf_rn(h ? qfy.y : qfy.x, Sy, Cfy);
            const float z0 = __fmaf_rn(h ? qnz.y : qnz.x, Sz, Cnz), z1 = __fmaf_rn(h ? qfz.y : qfz.x, Sz, Cfz);
            const float n = fmaxf(fmaxf(x0, y0), fmaxf(z0, T.tmin));
            const float f = fminf(fminf(x1, y1), fminf(z1, tcap));
            const bool hit = n <= f;  // an empty slot has qlo = +inf, qhi = -inf: n = +inf, f = -inf
            // n >= tmin > 0, so its bit pattern orders like the float: the slot index rides in the three lowest mantissa bits
            const uint32_t key = hit ? ((__float_as_uint(n) & ~7u) | (uint32_t)c) : (GK_N8_SENT | (uint32_t)c);
            if (GK_N8_TMIN) {  // smallest and second smallest key
                const uint32_t hi = max(kmin, key);
                kmin = min(kmin, key);
                k2 = min(k2, hi);
            } else {
                kmin = min(kmin, key);
            }
            mask |= hit ? (1u << c) : 0u;
        }
    }
    GK_STAT(T, T.cur_inst == GK_NO_HIT ? 0 : 1, 1);
    GK_STAT(T, 2, __popc(mask));
    if (kmin >= GK_N8_SENT) { trav_pop<8>(T, world, S, stack); return; }
    const uint32_t s1 = kmin & 7u, ni = h0.w >> 28, cb = h0.w & 0x00FFFFFFu, lb = GK_LEAF_BIT | h1.w;
    const uint32_t rest = mask & ~(1u << s1), low = (1u << ni) - 1u;
    // the entry distance of the nearest child that stays behind bounds both groups from below (the slot bits and the low byte of
    // the float are cleared: rounded down, i.e. conservative)
    const uint32_t tb = GK_N8_TMIN ? (k2 & ~0xFFu) : 0u;
    const uint32_t im = rest & low, lm = rest & ~low;
    if (im) { GK_STAT(T, 3, 1); stack.push(T.sp, make_uint2(cb, tb | im)); }
    if (lm) { GK_STAT(T, 3, 1); stack.push(T.sp, make_uint2(lb, tb | lm)); }  // leaves on top: they tighten the best hit before the deeper nodes
    T.cur = (s1 < ni ? cb : lb) + s1;
}

template <int FMT, class W, class St>
GK_DEV void trav_node_step(Trav& T, const W& world, const TraceScene& S, St& stack) {
    if (FMT == 8) trav_node_step8(T, world, S, stack);
    else trav_node_step4(T, world, S, stack);
}

// A TLAS leaf enters the instance's BLAS in object space (t is preserved by the affine map); false for an inactive instance.
template <int FMT, class W>
GK_DEV bool trav_enter_instance(Trav& T, const W& world, const TraceScene& S, uint32_t ii) {
    const float4* ir = reinterpret_cast<const float4*>(S.inst + ii);
    const float4 r0 = __ldg(ir), r1 = __ldg(ir + 1), r2 = __ldg(ir + 2);
    const uint4 meta = __ldg(reinterpret_cast<const uint4*>(ir) + 3);
    GK_STAT(T, 5, 1);
    if (!meta.z) return false;
    const float m[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
    const float3 rd = xform_vec(m, world.d());
    set_space<FMT>(T, xform_point(m, world.o()), f3(safe_rcp(rd.x), safe_rcp(rd.y), safe_rcp(rd.z)));
    T.sh = make_shear(rd);
    T.sp_inst = T.sp;
    T.nodes = reinterpret_cast<const uint4*>(S.blas + meta.x);
    if (FMT == 8) T.leaf = S.blas_leaf + meta.y;
    else T.tris = S.tris + (size_t)meta.y * 3;
    T.cur_inst = ii;
    T.cur = 0;
    return true;
}

// Leaf: a TLAS leaf enters the instance, a BLAS leaf tests its triangles.
template <int FMT, class W, class St>
GK_DEV void trav_leaf_step(Trav& T, const W& world, const TraceScene& S, St& stack) {
    uint32_t ref = T.cur & 0x0FFFFFFFu;
    if (FMT == 8) ref = __ldg(T.leaf + (T.cur & 0x7FFFFFFFu));  // leaf-table entry: instance index or triangle range
    if (T.cur_inst == GK_NO_HIT) {
        if (trav_enter_instance<FMT>(T, world, S, FMT == 8 ? ref : (T.cur & 0x0FFFFFFFu))) return;
    } else {
        const uint32_t first = FMT == 8 ? (ref & 0x0FFFFFFFu) : (T.cur & 0x0FFFFFFFu);
        const uint32_t count = FMT == 8 ? (ref >> 28) + 1u : ((T.cur >> 28) & 7u) + 1u;
        GK_STAT(T, 6, 1);
        GK_STAT(T, 7, count);
        for (uint32_t i = 0; i < count; ++i) {
            const float4* tp = T.tris + (size_t)(first + i) * 3;
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            float t, b1, b2;
            if (!tri_test(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), T.ro, T.sh, t, b1, b2)) continue;
            if (!(t > T.tmin && t < T.tmax)) continue;
            const uint32_t p = __float_as_uint(a.w);
            if (T.any) {
                T.best.t = t;
                T.best.inst = T.cur_inst;
                T.best.prim = p;
                T.done = true;
                return;
            }
            const bool better = T.best.inst == GK_NO_HIT || t < T.best.t ||
                                (t == T.best.t && (T.cur_inst < T.best.inst || (T.cur_inst == T.best.inst && p < T.best.prim)));
            if (better) {
                T.best.t = t;
                T.best.u = b1;
                T.best.v = b2;
                T.best.inst = T.cur_inst;
                T.best.prim = p;
            }
        }
    }
    trav_pop<FMT>(T, world, S, stack);
}

// Closest hit (ANY = false) or first hit (ANY = true; best.inst != GK_NO_HIT means occluded): the simple per-ray loop.
template <bool ANY, int FMT>
GK_DEV void trace_ray_fmt(const TraceScene& S, float3 o, float3 d, float tmin, float tmax, Hit& best) {
    LocalStack stack;
    Trav T;
    WorldRegs world;
    trav_init<FMT>(T, world, S, o, d, tmin, tmax, ANY);
    while (!T.done) {
        if (!(T.cur & GK_LEAF_BIT)) trav_node_step<FMT>(T, world, S, stack);
        else trav_leaf_step<FMT>(T, world, S, stack);
    }
    best = T.best;
}
// The same query for a whole warp: every lane of the warp calls it together (lanes without a ray pass active = false), and each
// iteration the warp executes only the step kind most of its rays wait for - node walk, instance entry or triangle leaf - like
// k_trace does, instead of serialising the three code paths whenever the lanes disagree. Per ray the sequence of steps is that of
// trace_ray_fmt, so the result is the same bits. Used by k_tail, where every thread runs a whole path and the rays of a warp
// are traced in lockstep anyway.
template <int FMT>
GK_DEV void trace_ray_warp_fmt(const TraceScene& S, float3 o, float3 d, float tmin, float tmax, bool any, bool active, Hit& best) {
    LocalStack stack;
    Trav T;
    WorldRegs world;
    trav_init<FMT>(T, world, S, o, d, tmin, tmax, any);
    if (!active) T.done = true;
    for (;;) {
        const bool live = !T.done;
        const bool leaf = (T.cur & GK_LEAF_BIT) != 0;
        const bool wN = live && !leaf;
        const bool wI = live && leaf && T.cur_inst == GK_NO_HIT;
        const bool wT = live && leaf && T.cur_inst != GK_NO_HIT;
        const int nN = __popc(__ballot_sync(0xffffffffu, wN)), nI = __popc(__ballot_sync(0xffffffffu, wI)), nT = __popc(__ballot_sync(0xffffffffu, wT));
        if (nN + nI + nT == 0) break;
        if (nN >= nT && nN >= nI) {
            if (wN) trav_node_step<FMT>(T, world, S, stack);
        } else if (nT >= nI) {
            if (wT) trav_leaf_step<FMT>(T, world, S, stack);
        } else {
            if (wI) trav_leaf_step<FMT>(T, world, S, stack);
        }
    }
    best = T.best;
}
GK_DEV void trace_ray_warp(const TraceScene& S, float3 o, float3 d, float tmin, float tmax, bool any, bool active, Hit& best) {
    if (S.fmt == 8) trace_ray_warp_fmt<8>(S, o, d, tmin, tmax, any, active, best);
    else trace_ray_warp_fmt<4>(S, o, d, tmin, tmax, any, active, best);
}
// The kernels off the hot path pick the format at run time (both loops are compiled in); k_trace is instantiated per format.
template <bool ANY>
GK_DEV void trace_ray(const TraceScene& S, float3 o, float3 d, float tmin, float tmax, Hit& best) {
    if (S.fmt == 8) trace_ray_fmt<ANY, 8>(S, o, d, tmin, tmax, best);
    else trace_ray_fmt<ANY, 4>(S, o, d, tmin, tmax, best);
}

}  // namespace gkd
