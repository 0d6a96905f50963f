"""Host-side structural validation of the GPU-built acceleration structures (used by tests/ and bring-up scripts).

Node4 layout: gknextengine_b200/csrc/gk_bvh.cuh. Checks: every primitive / instance is referenced exactly once, and every
child box contains what lies below it.
"""
import numpy as np

NODE_DTYPE = np.dtype([("planes", "<f4", (3, 2, 4)), ("child", "<u4", 4), ("n", "<u4"), ("pad", "<u4", 3)])  # planes[axis][lo|hi][child]
NODE_BYTES = NODE_DTYPE.itemsize
assert NODE_BYTES == 128
INST_DTYPE = np.dtype([("w2o", "<f4", (3, 4)), ("node_base", "<u4"), ("tri_base", "<u4"), ("active", "<u4"), ("blas", "<u4")])
LEAF = 0x80000000
EMPTY = 0xFFFFFFFF


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
