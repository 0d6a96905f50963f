// gk_bvh.cuh — acceleration-structure layouts and the stack traversal used by every tracing kernel.
// Replaces what VK_KHR_ray_query + RT cores do for the reference (RayQuery.TraceRayInline/Proceed at
// assets/shaders/common/Shading.slang:749-756, 794-801); B200 has no RT cores, so this is the hot loop.
//
// Layout (all 16-byte aligned, read with 128-bit loads):
//   Node4       64 B   4-wide node, children boxes quantised to 8 bits on a per-node power-of-two grid
//   leaf tris   48 B   3 x float4 (v0|prim id, v1, v2), stored in leaf order
//   InstanceRec 64 B   world->object 3x4 + BLAS offsets; TLAS leaves reference one instance each
// Semantics (shared with the CPU checker): a hit needs tmin < t < tmax, no face culling; the committed closest hit
// is the minimum of (t, instance, primitive) over all such hits, which makes the result independent of traversal order.
#pragma once
#include "gk_device.cuh"

namespace gkd {

struct __align__(16) Node4 {
    float ox, oy, oz;       // quantisation origin (node box min)
    uint8_t ex, ey, ez, n;  // biased exponents: cell size = 2^(e-127); n = child count
    uint32_t child[4];      // CHILD_EMPTY | leaf (bit31, count-1 in 30:28, first in 27:0) | node index
    uint32_t qlo[3];        // qlo[axis] byte c = child c
    uint32_t qhi[3];
    uint32_t pad[2];
};
static_assert(sizeof(Node4) == 64, "Node4 must be 64 bytes");

struct __align__(16) InstanceRec {
    float w2o[12];       // world->object, 3 rows x 4
    uint32_t node_base;  // BLAS root index in the BLAS node pool
    uint32_t tri_base;   // first leaf-ordered triangle of the BLAS in the triangle pool
    uint32_t active;     // visible && !nort && BLAS non-empty (VkAccelerationStructureInstanceKHR.mask)
    uint32_t blas;
};
static_assert(sizeof(InstanceRec) == 64, "InstanceRec must be 64 bytes");

#define GK_CHILD_EMPTY 0xFFFFFFFFu
#define GK_LEAF_BIT 0x80000000u
#define GK_STACK_SENTINEL 0xFFFFFFFEu
#define GK_TRAV_STACK 64
#define GK_NO_HIT 0xFFFFFFFFu
#ifndef GK_PUSH_ORDERED
#define GK_PUSH_ORDERED 1
#endif

struct TraceScene {
    const Node4* tlas;
    const Node4* blas;
    const float4* tris;
    const InstanceRec* inst;
    uint32_t tlas_nodes;  // 0 -> empty scene
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

GK_DEV float byte_f(uint32_t w, int c) { return (float)((w >> (8 * c)) & 0xFFu); }

// Tests the 4 quantised child boxes of a node. tn[c] = padded entry distance or +inf on a miss.
// Planes are dequantised relative to the node origin: t = q * (scale * idir) + (origin - o) * idir, one FMA per plane; the
// near/far plane words are picked per axis from the sign of the ray direction, so no per-plane min/max is needed. The
// interval is padded by a few ulps on both ends (the quantised boxes are conservative, the slab arithmetic is not).
GK_DEV void node_test(const uint4 n0, const uint4 n1, const uint4 n2, const uint4 n3, float3 ro, float3 idir, float tmin, float tbest, float tn[4]) {
    const float ox = __uint_as_float(n0.x), oy = __uint_as_float(n0.y), oz = __uint_as_float(n0.z);
    const float sx = __uint_as_float((n0.w & 0xFFu) << 23), sy = __uint_as_float(((n0.w >> 8) & 0xFFu) << 23),
                sz = __uint_as_float(((n0.w >> 16) & 0xFFu) << 23);
    const float adjx = (ox - ro.x) * idir.x, adjy = (oy - ro.y) * idir.y, adjz = (oz - ro.z) * idir.z;
    const float scx = sx * idir.x, scy = sy * idir.y, scz = sz * idir.z;
    const uint32_t child[4] = {n1.x, n1.y, n1.z, n1.w};
    const bool px = idir.x >= 0.0f, py = idir.y >= 0.0f, pz = idir.z >= 0.0f;
    const uint32_t nx = px ? n2.x : n2.w, fx = px ? n2.w : n2.x;  // qlo.x / qhi.x
    const uint32_t ny = py ? n2.y : n3.x, fy = py ? n3.x : n2.y;
    const uint32_t nz = pz ? n2.z : n3.y, fz = pz ? n3.y : n2.z;
    const float tfcap = __fmaf_rn(tbest, 1.000001f, 1e-7f);
    const float inf = __int_as_float(0x7f800000);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float x0 = __fmaf_rn(byte_f(nx, c), scx, adjx), x1 = __fmaf_rn(byte_f(fx, c), scx, adjx);
        const float y0 = __fmaf_rn(byte_f(ny, c), scy, adjy), y1 = __fmaf_rn(byte_f(fy, c), scy, adjy);
        const float z0 = __fmaf_rn(byte_f(nz, c), scz, adjz), z1 = __fmaf_rn(byte_f(fz, c), scz, adjz);
        const float n = fmaxf(fmaxf(x0, y0), fmaxf(z0, tmin));
        const float f = fminf(fminf(x1, y1), z1);
        const float npad = __fmaf_rn(n, 0.999999f, -1e-7f);
        const bool hit = (child[c] != GK_CHILD_EMPTY) && (npad <= __fmaf_rn(f, 1.000001f, 1e-7f)) && (npad <= tfcap);
        tn[c] = hit ? npad : inf;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Traversal state machine. One ray per lane; the same three steps serve the simple per-ray loop (trace_ray) and the
// persistent-thread kernel (k_trace), which interleaves them "while-while" style so that a warp runs node steps together
// and leaf steps together, and refills finished lanes with new rays.
struct Trav {
    float3 o, d;       // world-space ray
    float3 ro, rd;     // current-space ray (world or object)
    float3 idir;
    float3 widir;      // reciprocal world-space direction (restored when an instance is left)
    float tmin, tmax;
    Hit best;
    RayShear sh;
    const uint4* nodes;
    const float4* tris;
    uint32_t cur;      // current node reference
    uint32_t cur_inst; // GK_NO_HIT while in the TLAS
    int sp;
    bool any;          // first-hit query
    bool done;
};

GK_DEV void trav_init(Trav& T, const TraceScene& S, float3 o, float3 d, float tmin, float tmax, bool any) {
    T.o = o; T.d = d; T.ro = o; T.rd = d;
    T.idir = f3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
    T.widir = T.idir;
    T.tmin = tmin; T.tmax = tmax;
    T.best.t = tmax; T.best.u = 0.0f; T.best.v = 0.0f; T.best.inst = GK_NO_HIT; T.best.prim = 0;
    T.sh.kx = T.sh.ky = T.sh.kz = 0;
    T.sh.Sx = T.sh.Sy = T.sh.Sz = 0.0f;
    T.nodes = reinterpret_cast<const uint4*>(S.tlas);
    T.tris = S.tris;
    T.cur = 0;
    T.cur_inst = GK_NO_HIT;
    T.sp = 0;
    T.any = any;
    T.done = S.tlas_nodes == 0;
}

// Pops the next node to visit (culling entries that are already farther than the best hit); leaves instances on sentinels.
GK_DEV void trav_pop(Trav& T, const TraceScene& S, uint2* stack) {
    for (;;) {
        if (T.sp == 0) { T.done = true; return; }
        const uint2 e = stack[--T.sp];
        if (e.x == GK_STACK_SENTINEL) {
            T.ro = T.o;
            T.rd = T.d;
            T.idir = T.widir;
            T.nodes = reinterpret_cast<const uint4*>(S.tlas);
            T.cur_inst = GK_NO_HIT;
            continue;
        }
        if (__uint_as_float(e.y) > T.best.t) continue;
        T.cur = e.x;
        return;
    }
}

// Internal node: test the 4 children, continue with the nearest, push the others (second nearest on top).
GK_DEV void trav_node_step(Trav& T, const TraceScene& S, uint2* stack) {
    const uint4* np = T.nodes + (size_t)T.cur * 4;
    const uint4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
    float tn[4];
    node_test(n0, n1, n2, n3, T.ro, T.idir, T.tmin, T.best.t, tn);
    const uint32_t ref[4] = {n1.x, n1.y, n1.z, n1.w};
    const float inf = __int_as_float(0x7f800000);
    // continue with the nearest child; the second nearest goes on top of the stack, the rest below it in slot order
    // (entries carry their entry distance, so a far one is culled when it is popped)
    int i0 = 0;
    float t0 = tn[0];
    if (tn[1] < t0) { t0 = tn[1]; i0 = 1; }
    if (tn[2] < t0) { t0 = tn[2]; i0 = 2; }
    if (tn[3] < t0) { t0 = tn[3]; i0 = 3; }
    if (!(t0 < inf)) { trav_pop(T, S, stack); return; }
    if (GK_PUSH_ORDERED) {
        int i1 = -1;
        float t1 = inf;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c != i0 && tn[c] < t1) { t1 = tn[c]; i1 = c; }
        if (i1 >= 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c != i0 && c != i1 && tn[c] < inf) stack[T.sp++] = make_uint2(ref[c], __float_as_uint(tn[c]));
            stack[T.sp++] = make_uint2(ref[i1], __float_as_uint(t1));
        }
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c != i0 && tn[c] < inf) stack[T.sp++] = make_uint2(ref[c], __float_as_uint(tn[c]));
    }
    T.cur = ref[i0];
}

// Leaf: a TLAS leaf enters the instance's BLAS in object space (t is preserved by the affine map); a BLAS leaf tests triangles.
GK_DEV void trav_leaf_step(Trav& T, const TraceScene& S, uint2* stack) {
    if (T.cur_inst == GK_NO_HIT) {
        const uint32_t ii = T.cur & 0x0FFFFFFFu;
        const float4* ir = reinterpret_cast<const float4*>(S.inst + ii);
        const float4 r0 = __ldg(ir), r1 = __ldg(ir + 1), r2 = __ldg(ir + 2);
        const uint4 meta = __ldg(reinterpret_cast<const uint4*>(ir) + 3);
        if (meta.z) {
            const float m[12] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w};
            T.ro = xform_point(m, T.o);
            T.rd = xform_vec(m, T.d);
            T.idir = f3(safe_rcp(T.rd.x), safe_rcp(T.rd.y), safe_rcp(T.rd.z));
            T.sh = make_shear(T.rd);
            stack[T.sp++] = make_uint2(GK_STACK_SENTINEL, 0u);
            T.nodes = reinterpret_cast<const uint4*>(S.blas + meta.x);
            T.tris = S.tris + (size_t)meta.y * 3;
            T.cur_inst = ii;
            T.cur = 0;
            return;
        }
    } else {
        const uint32_t first = T.cur & 0x0FFFFFFFu, count = ((T.cur >> 28) & 7u) + 1u;
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
    trav_pop(T, S, stack);
}

// Closest hit (ANY = false) or first hit (ANY = true; best.inst != GK_NO_HIT means occluded): the simple per-ray loop.
template <bool ANY>
GK_DEV void trace_ray(const TraceScene& S, float3 o, float3 d, float tmin, float tmax, Hit& best) {
    uint2 stack[GK_TRAV_STACK];
    Trav T;
    trav_init(T, S, o, d, tmin, tmax, ANY);
    while (!T.done) {
        if (!(T.cur & GK_LEAF_BIT)) trav_node_step(T, S, stack);
        else trav_leaf_step(T, S, stack);
    }
    best = T.best;
}

}  // namespace gkd
