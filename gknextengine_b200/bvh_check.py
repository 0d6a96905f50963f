"""Host-side structural validation of the GPU-built acceleration structures (used by tests/ and bring-up scripts).

Node4 / Node8 layouts: gknextengine_b200/csrc/gk_bvh.cuh. Checks: every primitive / instance is referenced exactly once, and every
child box contains what lies below it (for the quantised Node8: the de-quantised box, evaluated in float64).
"""
import numpy as np

NODE_DTYPE = np.dtype([("planes", "<f4", (3, 2, 4)), ("child", "<u4", 4), ("n", "<u4"), ("pad", "<u4", 3)])  # planes[axis][lo|hi][child]
NODE_BYTES = NODE_DTYPE.itemsize
assert NODE_BYTES == 128
INST_DTYPE = np.dtype([("w2o", "<f4", (3, 4)), ("node_base", "<u4"), ("tri_base", "<u4"), ("active", "<u4"), ("blas", "<u4")])
NODE8_DTYPE = np.dtype([("o", "<f4", 3), ("link", "<u4"), ("s", "<f4", 3), ("lb", "<u4"), ("q", "<f2", (6, 8))])  # q[lo_x, hi_x, lo_y, hi_y, lo_z, hi_z][child]
assert NODE8_DTYPE.itemsize == 128
LEAF = 0x80000000
EMPTY = 0xFFFFFFFF


def child_boxes8(node):
    """De-quantised child boxes of a Node8 and its (n, ni, child_base, lb)."""
    link = int(node["link"])
    n, ni, cb = (link >> 24) & 15, link >> 28, link & 0xFFFFFF
    o, s, q = node["o"].astype(np.float64), node["s"].astype(np.float64), node["q"].astype(np.float64)
    lo = np.stack([o[a] + q[2 * a] * s[a] for a in range(3)], axis=1)  # [child][axis]
    hi = np.stack([o[a] + q[2 * a + 1] * s[a] for a in range(3)], axis=1)
    return lo, hi, n, ni, cb, int(node["lb"])


def walk8(nodes, leaf_table, leaf_box_fn, node_lo, node_hi, tol=1e-5):
    """Node8 counterpart of walk(): returns (list of leaf-table entries reached, max depth, containment violations). A child node is
    checked with its exact fp32 box (the builder's side array): its own children are rounded outward on ITS grid and may stick out of the
    box its parent stores for it by a cell, which is harmless - the geometry lies inside both."""
    leaves, viol, maxdepth = [], 0, 0
    stack = [(0, 1)]
    visited = set()
    while stack:
        idx, depth = stack.pop()
        assert idx not in visited, f"node {idx} referenced twice"
        visited.add(idx)
        maxdepth = max(maxdepth, depth)
        lo, hi, n, ni, cb, lb = child_boxes8(nodes[idx])
        q = nodes[idx]["q"]
        assert np.isinf(q[:, n:].astype(np.float64)).all(), "empty slots must hold +-inf planes"
        assert (q[:, :n] >= 0).all() and (q[:, :n] <= 2050).all() and (q[:, :n] == np.round(q[:, :n])).all(), "planes must be integers on the 11-bit grid"
        for c in range(n):
            if c >= ni:
                rec = int(leaf_table[lb + c])
                blo, bhi = leaf_box_fn(rec)
                leaves.append(rec)
            else:
                blo, bhi = node_lo[cb + c, :3].astype(np.float64), node_hi[cb + c, :3].astype(np.float64)
                clo, chi, cn, _, _, _ = child_boxes8(nodes[cb + c])
                if (clo[:cn].min(0) > blo + tol * (1 + np.abs(blo))).any() or (chi[:cn].max(0) < bhi - tol * (1 + np.abs(bhi))).any():
                    viol += 1  # the child's own boxes must cover its exact box
                stack.append((cb + c, depth + 1))
            pad = tol * (1 + np.abs(bhi) + np.abs(blo))
            if (blo < lo[c] - pad).any() or (bhi > hi[c] + pad).any():
                viol += 1
    return leaves, maxdepth, viol


def child_boxes(node):
    pl = node["planes"].astype(np.float64)
    return pl[:, 0, :].T, pl[:, 1, :].T  # [child][axis]


def walk(nodes, leaf_box_fn, tol=1e-5):
    """Returns (list of leaf refs, max depth, number of containment violations)."""
    leaves, viol, maxdepth = [], 0, 0
    stack = [(0, None, None, 1)]
    visited = set()
    while stack:
        idx, plo, phi, depth = stack.pop()
        assert idx not in visited, f"node {idx} referenced twice"
        visited.add(idx)
        maxdepth = max(maxdepth, depth)
        nd = nodes[idx]
        lo, hi = child_boxes(nd)
        for c in range(4):
            ref = int(nd["child"][c])
            if ref == EMPTY:
                continue
            if ref & LEAF:
                first, count = ref & 0x0FFFFFFF, ((ref >> 28) & 7) + 1
                blo, bhi = leaf_box_fn(first, count)
                pad = tol * (1 + np.abs(bhi) + np.abs(blo))
                if (blo < lo[c] - pad).any() or (bhi > hi[c] + pad).any():
                    viol += 1
                leaves.append((first, count))
            else:
                # the union of the child node's own child boxes must lie inside the box its parent stores for it
                clo, chi = child_boxes(nodes[ref])
                used = nodes[ref]["child"] != EMPTY
                ulo, uhi = clo[used].min(0), chi[used].max(0)
                pad = tol * (1 + np.abs(ulo) + np.abs(uhi))
                if (ulo < lo[c] - pad).any() or (uhi > hi[c] + pad).any():
                    viol += 1
                stack.append((ref, lo[c], hi[c], depth + 1))
    return leaves, maxdepth, viol


def validate(renderer):
    info = renderer.debug_info()
    if info["node_format"] == 8:
        return validate8(renderer, info)
    report = dict(info=info)
    tris_all = None
    blas_boxes = []
    for m, bm in enumerate(info["models"]):
        if bm["prim_count"] == 0:
            blas_boxes.append(None)
            continue
        nodes = renderer.debug_read(1, bm["node_base"] * NODE_BYTES, bm["node_count"], NODE_DTYPE)
        tris = renderer.debug_read(2, bm["tri_base"] * 48, bm["prim_count"] * 12, np.float32).reshape(-1, 3, 4)

        def leaf_box(first, count, tris=tris):
            v = tris[first:first + count, :, :3].reshape(-1, 3).astype(np.float64)
            return v.min(0), v.max(0)

        leaves, depth, viol = walk(nodes, leaf_box)
        covered = np.zeros(bm["prim_count"], dtype=np.int32)
        for f, c in leaves:
            covered[f:f + c] += 1
        prim_ids = tris[:, 0, 3].view(np.uint32)
        report[f"blas{m}"] = dict(nodes=len(nodes), depth=depth, violations=viol, all_tris_once=bool((covered == 1).all()),
                                  prim_ids_are_permutation=bool((np.sort(prim_ids) == np.arange(bm["prim_count"])).all()))
        v = tris[:, :, :3].reshape(-1, 3).astype(np.float64)
        blas_boxes.append((v.min(0), v.max(0)))
    if info["tlas_nodes"]:
        nodes = renderer.debug_read(0, 0, info["tlas_nodes"], NODE_DTYPE)
        inst = renderer.debug_read(3, 0, info["instances"], INST_DTYPE)

        def inst_box(first, count):
            i = first
            bb = blas_boxes[int(inst[i]["blas"])]
            w2o = np.eye(4)
            w2o[:3, :] = inst[i]["w2o"].astype(np.float64)
            o2w = np.linalg.inv(w2o)
            corners = np.array([[bb[(c >> a) & 1][a] for a in range(3)] + [1.0] for c in range(8)])
            w = (o2w @ corners.T).T[:, :3]
            return w.min(0), w.max(0)

        leaves, depth, viol = walk(nodes, inst_box, tol=1e-4)
        ids = sorted(f for f, _ in leaves)
        active = [i for i in range(info["instances"]) if inst[i]["active"]]
        report["tlas"] = dict(nodes=len(nodes), depth=depth, violations=viol, leaves=len(ids), active=len(active),
                              all_instances_once=ids == active)
    return report


def validate8(renderer, info):
    report = dict(info=info)
    blas_boxes = []
    total_tris = sum(bm["prim_count"] for bm in info["models"])
    tris_pool = renderer.debug_read(2, 0, total_tris * 12, np.float32).reshape(-1, 3, 4) if total_tris else None
    for m, bm in enumerate(info["models"]):
        if bm["prim_count"] == 0:
            blas_boxes.append(None)
            continue
        nodes = renderer.debug_read(1, bm["node_base"] * NODE_BYTES, bm["node_count"], NODE8_DTYPE)
        leaf = renderer.debug_read(9, bm["leaf_base"] * 4, bm["leaf_count"], np.uint32)

        def leaf_box(rec):
            first, count = rec & 0x0FFFFFFF, (rec >> 28) + 1
            v = tris_pool[first:first + count, :, :3].reshape(-1, 3).astype(np.float64)
            return v.min(0), v.max(0)

        nlo = renderer.debug_read(10, bm["node_base"] * 16, bm["node_count"] * 4, np.float32).reshape(-1, 4)
        nhi = renderer.debug_read(11, bm["node_base"] * 16, bm["node_count"] * 4, np.float32).reshape(-1, 4)
        leaves, depth, viol = walk8(nodes, leaf, leaf_box, nlo, nhi)
        covered = np.zeros(total_tris, dtype=np.int32)
        for rec in leaves:
            covered[(rec & 0x0FFFFFFF):(rec & 0x0FFFFFFF) + (rec >> 28) + 1] += 1
        mine = slice(bm["tri_base"], bm["tri_base"] + bm["prim_count"])
        tris = tris_pool[mine]
        prim_ids = tris[:, 0, 3].view(np.uint32)
        report[f"blas{m}"] = dict(nodes=len(nodes), depth=depth, violations=viol, all_tris_once=bool((covered[mine] == 1).all() and covered.sum() == bm["prim_count"]),
                                  prim_ids_are_permutation=bool((np.sort(prim_ids) == np.arange(bm["prim_count"])).all()))
        v = tris[:, :, :3].reshape(-1, 3).astype(np.float64)
        blas_boxes.append((v.min(0), v.max(0)))
    if info["tlas_nodes"]:
        nodes = renderer.debug_read(0, 0, info["tlas_nodes"], NODE8_DTYPE)
        leaf = renderer.debug_read(8, 0, info["tlas_leaf_count"], np.uint32)
        inst = renderer.debug_read(3, 0, info["instances"], INST_DTYPE)

        def inst_box(i):
            bb = blas_boxes[int(inst[i]["blas"])]
            w2o = np.eye(4)
            w2o[:3, :] = inst[i]["w2o"].astype(np.float64)
            o2w = np.linalg.inv(w2o)
            corners = np.array([[bb[(c >> a) & 1][a] for a in range(3)] + [1.0] for c in range(8)])
            w = (o2w @ corners.T).T[:, :3]
            return w.min(0), w.max(0)

        nlo = renderer.debug_read(4, 0, info["tlas_nodes"] * 4, np.float32).reshape(-1, 4)
        nhi = renderer.debug_read(5, 0, info["tlas_nodes"] * 4, np.float32).reshape(-1, 4)
        leaves, depth, viol = walk8(nodes, leaf, inst_box, nlo, nhi, tol=1e-4)
        ids = sorted(leaves)
        active = [i for i in range(info["instances"]) if inst[i]["active"]]
        report["tlas"] = dict(nodes=len(nodes), depth=depth, violations=viol, leaves=len(ids), active=len(active),
                              all_instances_once=ids == active)
    return report
