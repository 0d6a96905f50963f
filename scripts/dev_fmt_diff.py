"""Bring-up: traces the voxeliser's rays (probe-grid origins, axis and diagonal directions, tmin 0) through both node formats and, for rays whose
hits differ, walks the Node8 tree on the host (float64) along the path to the Node4 hit to find the box that lost it.
python scripts/dev_fmt_diff.py [scene]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gknextengine_b200 import CudaPathTracingRenderer, host
from gknextengine_b200._abi import RAY_DTYPE
from gknextengine_b200.bvh_check import NODE8_DTYPE, INST_DTYPE, NODE_BYTES, child_boxes8

scene_name = sys.argv[1] if len(sys.argv) > 1 else "RTIO.proc"
scene = host.Scene(scene_name, samples=1, bounces=4, taa=False)
scene.update_nodes()
rs = {}
for fmt in (4, 8):
    os.environ["GKCU_NODE_FORMAT"] = str(fmt)
    r = CudaPathTracingRenderer(0)
    r.CreateSwapChain((64, 64))
    r.OnPostLoadScene(scene, host.make_sky(64, 32))
    r.PreRender()
    rs[fmt] = r
rng = np.random.default_rng(5)
n = 400000
gx = rng.integers(0, 192, n); gy = rng.integers(0, 48, n); gz = rng.integers(0, 192, n)
o = np.stack([(gx - 96) * 0.25, (gy - 1.375) * 0.25, (gz - 96) * 0.25], 1).astype(np.float32)
dirs = np.array([[0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1], [1, 1, 1], [-1, 1, 1], [-1, -1, 1], [1, 1, -1], [-1, 1, -1], [-1, -1, -1]], np.float32)
d = dirs[rng.integers(0, len(dirs), n)]
d = (d / np.sqrt((d * d).sum(1, keepdims=True)).astype(np.float32)).astype(np.float32)
rays = np.zeros(n, RAY_DTYPE)
rays["ox"], rays["oy"], rays["oz"] = o.T
rays["dx"], rays["dy"], rays["dz"] = d.T
rays["tmin"] = 0.0
rays["tmax"] = 16.0
h4, h8 = rs[4].trace(rays), rs[8].trace(rays)
bad = np.nonzero((h4["inst"] != h8["inst"]) | (h4["prim"] != h8["prim"]) | (h4["t"].view(np.uint32) != h8["t"].view(np.uint32)))[0]
print("rays", n, "mismatches", len(bad))
r8 = rs[8]
info = r8.debug_info()
tn = r8.debug_read(0, 0, info["tlas_nodes"], NODE8_DTYPE)
tl = r8.debug_read(8, 0, info["tlas_leaf_count"], np.uint32)
inst = r8.debug_read(3, 0, info["instances"], INST_DTYPE)
total_tris = sum(m["prim_count"] for m in info["models"])
tris = r8.debug_read(2, 0, total_tris * 12, np.float32).reshape(-1, 3, 4)


def chain(nodes, leaf, pred):
    """path of (node, slot) from the root to the leaf entry satisfying pred(rec)"""
    def rec(idx, path):
        lo, hi, nk, ni, cb, lb = child_boxes8(nodes[idx])
        for c in range(nk):
            if c >= ni:
                if pred(int(leaf[lb + c])):
                    return path + [(idx, c)]
            else:
                p = rec(cb + c, path + [(idx, c)])
                if p:
                    return p
        return None
    return rec(0, [])


def slab(node, c, ro, dr):
    lo, hi, *_ = child_boxes8(node)
    with np.errstate(divide="ignore", invalid="ignore"):
        t0 = (lo[c] - ro) / dr
        t1 = (hi[c] - ro) / dr
    near = np.where(dr == 0, np.where((ro >= lo[c]) & (ro <= hi[c]), -np.inf, np.inf), np.minimum(t0, t1))
    far = np.where(dr == 0, np.where((ro >= lo[c]) & (ro <= hi[c]), np.inf, -np.inf), np.maximum(t0, t1))
    return near.max(), far.min(), lo[c], hi[c]


for i in bad[:4]:
    ray = rays[i]
    print("ray", ray, "fmt4", h4[i], "fmt8", h8[i])
    good = h4[i] if h4[i]["inst"] != 0xFFFFFFFF else h8[i]
    ii = int(good["inst"])
    ro = np.array([ray["ox"], ray["oy"], ray["oz"]], np.float64); dr = np.array([ray["dx"], ray["dy"], ray["dz"]], np.float64)
    path = chain(tn, tl, lambda rec: rec == ii)
    print(" TLAS path", path)
    for idx, c in path or []:
        print("   node", idx, "slot", c, "slab near/far/lo/hi", slab(tn[idx], c, ro, dr), "raw", tn[idx]["o"], tn[idx]["s"], tn[idx]["q"][:, c], hex(int(tn[idx]["link"])))
    m = int(inst[ii]["blas"])
    bm = info["models"][m]
    bn = r8.debug_read(1, bm["node_base"] * NODE_BYTES, bm["node_count"], NODE8_DTYPE)
    bl = r8.debug_read(9, bm["leaf_base"] * 4, bm["leaf_count"], np.uint32)
    w2o = inst[ii]["w2o"].astype(np.float32)
    ro_o = (w2o[:, :3] @ ro.astype(np.float32) + w2o[:, 3]).astype(np.float64); dr_o = (w2o[:, :3] @ dr.astype(np.float32)).astype(np.float64)
    prim = int(good["prim"])
    def has_prim(rec):
        f, cnt = rec & 0x0FFFFFFF, (rec >> 28) + 1
        return prim in tris[f:f + cnt, 0, 3].view(np.uint32).tolist() and bm["tri_base"] <= f < bm["tri_base"] + bm["prim_count"]
    path = chain(bn, bl, has_prim)
    print(" BLAS", m, "path", path, "obj ray", ro_o, dr_o, "t", good["t"])
    for idx, c in path or []:
        print("   node", idx, "slot", c, "slab near/far/lo/hi", slab(bn[idx], c, ro_o, dr_o), "raw", bn[idx]["o"], bn[idx]["s"], bn[idx]["q"][:, c], hex(int(bn[idx]["link"])))
