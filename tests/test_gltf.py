"""The asset path: the from-scratch GLB reader with the reference's scene mapping rules (FSceneLoader.cpp) and the Radiance
.hdr reader, checked against independent Python restatements. tests/golden/pbr_like.glb is synthetic (tests/golden/make_glb.py); the
reference's own pbr.glb / river_road_2.hdr are read where the reference is mounted."""
import ctypes as C
import json
import os
import struct

import numpy as np
import pytest

from gknextengine_b200 import host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLB = os.path.join(ROOT, "tests", "golden", "pbr_like.glb")
REF = "/root/reference/assets"


def parse_glb(path):
    f = open(path, "rb").read()
    magic, ver, _ = struct.unpack("<III", f[:12])
    assert magic == 0x46546C67 and ver == 2
    n, t = struct.unpack("<II", f[12:20])
    doc = json.loads(f[20:20 + n])
    off = 20 + n
    m, t = struct.unpack("<II", f[off:off + 8])
    return doc, f[off + 8:off + 8 + m]


def read_accessor(doc, blob, idx):
    a = doc["accessors"][idx]
    v = doc["bufferViews"][a["bufferView"]]
    comps = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4}[a["type"]]
    dt = {5126: np.float32, 5123: np.uint16, 5125: np.uint32}[a["componentType"]]
    stride = v.get("byteStride", comps * np.dtype(dt).itemsize)
    base = v.get("byteOffset", 0) + a.get("byteOffset", 0)
    out = np.zeros((a["count"], comps), dt)
    for i in range(a["count"]):
        out[i] = np.frombuffer(blob, dt, comps, base + i * stride)
    return out


def to_half(x):
    return np.array([host.lib().gkhost_to_float16(float(v)) for v in np.asarray(x, np.float32).ravel()], np.uint16)


def test_glb_reader_follows_the_reference_mapping_rules():
    doc, blob = parse_glb(GLB)
    s = host.Scene(GLB)
    d = s.desc()
    assert d.model_count == len(doc["meshes"])
    verts = np.ctypeslib.as_array(C.cast(d.vertices, C.POINTER(C.c_uint16)), shape=(d.vertex_count, 12))
    prim = np.ctypeslib.as_array(d.indices, shape=(d.index_count,))
    # ---- one Model per mesh, primitives concatenated, MaterialIndex = section (FSceneLoader.cpp:394-510)
    for m, mesh in enumerate(doc["meshes"]):
        off = d.offsets[m * 10]
        pos, nrm, uv, idx, slot = [], [], [], [], []
        vo = 0
        for sidx, p in enumerate(mesh["primitives"]):
            pp = read_accessor(doc, blob, p["attributes"]["POSITION"])
            pos.append(pp); nrm.append(read_accessor(doc, blob, p["attributes"]["NORMAL"])); uv.append(read_accessor(doc, blob, p["attributes"]["TEXCOORD_0"]))
            idx.append(read_accessor(doc, blob, p["indices"]).ravel().astype(np.uint32) + vo)
            slot += [sidx] * len(pp)
            vo += len(pp)
        pos, nrm, uv, idx = np.concatenate(pos), np.concatenate(nrm), np.concatenate(uv), np.concatenate(idx)
        assert off.vertexCount == len(pos) and sum(d.offsets[m * 10 + k].indexCount for k in range(10)) == len(idx)
        v = verts[off.vertexOffset:off.vertexOffset + off.vertexCount]
        assert (v[:, 0:3].ravel() == to_half(pos)).all() and (v[:, 4:7].ravel() == to_half(nrm)).all()
        assert (v[:, 3] == to_half(uv[:, 0])).all() and (v[:, 7] == to_half(uv[:, 1])).all()
        assert ((v[:, 11] & 0xFF) == np.array(slot)).all()
        tri = prim[off.indexOffset:off.indexOffset + len(idx)].reshape(-1, 3)
        src = idx.reshape(-1, 3)
        assert (np.sort(tri, axis=1) == np.sort(src, axis=1)).all()  # same triangles (corners rotated by the provoking-vertex pass)
    # ---- materials (:312-391): root_default first, then the file's
    mats, nm = s.materials()
    assert nm == 1 + len(doc["materials"])
    LAMBERT, METAL, DIELECTRIC, LIGHT, MIXTURE = 0, 1, 2, 4, 5
    for i, gm in enumerate(doc["materials"]):
        m = mats[1 + i]
        pbr = gm["pbrMetallicRoughness"]
        rough, metal = np.float32(pbr.get("roughnessFactor", 1.0)), np.float32(pbr.get("metallicFactor", 1.0))
        ext = gm.get("extensions", {})
        if "KHR_materials_emissive_strength" in ext:
            e = np.float32(gm["emissiveFactor"]) * np.float32(ext["KHR_materials_emissive_strength"]["emissiveStrength"]) * np.float32(100)
            assert m.MaterialModel == LIGHT and np.allclose([m.Diffuse.x, m.Diffuse.y, m.Diffuse.z], e, rtol=1e-6)
            continue
        want = MIXTURE
        if metal > 0.95: want = METAL
        if rough > 0.95 and metal < 0.01: want = LAMBERT
        if ext.get("KHR_materials_transmission", {}).get("transmissionFactor", 0) > 0.8: want = DIELECTRIC
        assert m.MaterialModel == want, gm["name"]
        assert m.Fuzziness == rough and m.Metalness == metal
        assert np.allclose(m.Diffuse.x, np.sqrt(np.float32(pbr["baseColorFactor"][0])), rtol=1e-6) and m.Diffuse.w == 1.0
        ior = np.float32(ext.get("KHR_materials_ior", {}).get("ior", 1.46))
        assert m.RefractionIndex == ior and m.RefractionIndex2 == np.float32(gm.get("extras", {}).get("ior2", ior))
        assert (m.DiffuseTextureId, m.MRATextureId, m.NormalTextureId) == (-1, -1, -1)
    # ---- nodes (:102-187): depth first, instanceId = node order, matId[section] = material + 1, children below their parent
    proxies, n = s.nodes()
    order = []

    def walk(i, parent):
        nd = doc["nodes"][i]
        t, r, sc = nd.get("translation", [0, 0, 0]), nd.get("rotation", [0, 0, 0, 1]), nd.get("scale", [1, 1, 1])
        x, y, z, w = r
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        M = np.eye(4)
        M[:3, :3] = R @ np.diag(sc)
        M[:3, 3] = t
        M = parent @ M
        order.append((i, M))
        for c in nd.get("children", []):
            walk(c, M)

    for i in doc["scenes"][0]["nodes"]:
        if i not in {c for nd in doc["nodes"] for c in nd.get("children", [])}:
            walk(i, np.eye(4))
    k = 0
    for inst, (i, M) in enumerate(order):
        nd = doc["nodes"][i]
        if "mesh" not in nd:
            continue
        p = proxies[k]
        k += 1
        assert p.instanceId == inst and p.modelId == nd["mesh"] * 10 and p.visible == 1 and p.nort == 0
        got = np.array([[p.worldTS.m[c][r] for c in range(4)] for r in range(4)])
        assert np.allclose(got, M, atol=1e-5), nd["name"]
        prims = doc["meshes"][nd["mesh"]]["primitives"]
        assert list(p.matId[:len(prims)]) == [q["material"] + 1 for q in prims]
    assert k == n
    # ---- camera (:126-138): fixed 40 degree field of view; environment from the scene extras (:517-552)
    ubo = s.ubo(1920, 1080, 1)
    proj11 = ubo.Projection.m[1][1]
    assert np.isclose(abs(proj11), 1.0 / np.tan(np.radians(40.0) / 2), rtol=1e-5)
    eye = np.array([ubo.ModelViewInverse.m[3][r] for r in range(3)])
    assert np.allclose(eye, doc["nodes"][0]["translation"], atol=1e-3)
    assert ubo.HasSky == 1 and ubo.HasSun == 1 and np.isclose(ubo.SkyIntensity, 60.0)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "models", "pbr.glb")), reason="reference assets not mounted")
def test_reference_pbr_glb_loads_with_the_surveyed_counts():
    s = host.Scene("pbr.glb")  # looked up in the reference's assets/models
    assert s.counts() == dict(nodes=11, models=10, materials=11, instanced_triangles=8764)  # "pbr.glb (8 764 tris)"
    mats, nm = s.materials()
    assert [mats[i].MaterialModel for i in range(nm)] == [0, 0, 5, 5, 5, 5, 0, 5, 1, 1, 1]
    proxies, n = s.nodes()
    assert n == 10 and [proxies[i].instanceId for i in range(n)] == list(range(1, 11))


def test_broken_and_unsupported_files_are_rejected(tmp_path):
    p = tmp_path / "bad.glb"
    p.write_bytes(b"not a glb at all")
    with pytest.raises(ValueError, match="magic"):
        host.Scene(str(p))
    doc, blob = parse_glb(GLB)
    doc["extensionsRequired"] = ["KHR_draco_mesh_compression"]
    js = json.dumps(doc).encode()
    js += b" " * ((-len(js)) % 4)
    p = tmp_path / "draco.glb"
    p.write_bytes(struct.pack("<III", 0x46546C67, 2, 28 + len(js) + len(blob)) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(blob), 0x004E4942) + blob)
    with pytest.raises(ValueError, match="compressed"):
        host.Scene(str(p))
    with pytest.raises(ValueError):
        host.Scene(str(tmp_path / "missing.glb"))


def rgbe_encode(img):
    """float32 (H, W, 3) -> RGBE bytes (H, W, 4), Ward's float2rgbe."""
    m = img.max(axis=2)
    e = np.zeros(m.shape, np.int32)
    out = np.zeros(img.shape[:2] + (4,), np.uint8)
    nz = m > 1e-32
    mant, ex = np.frexp(m[nz])
    scale = mant * 256.0 / m[nz]
    out[nz, :3] = (img[nz] * scale[:, None]).astype(np.uint8)
    out[nz, 3] = (ex + 128).astype(np.uint8)
    return out


def rle_channel(row):
    out = bytearray()
    i, n = 0, len(row)
    while i < n:
        run = 1
        while i + run < n and run < 127 and row[i + run] == row[i]:
            run += 1
        if run >= 4:
            out += bytes([128 + run, int(row[i])])
            i += run
        else:
            j = i
            while j < n and j - i < 128 and not (j + 3 < n and row[j] == row[j + 1] == row[j + 2] == row[j + 3]):
                j += 1
            j = max(j, i + 1)
            out += bytes([j - i]) + bytes(int(v) for v in row[i:j])
            i = j
    return bytes(out)


def rgbe_decode(px):
    e = px[..., 3].astype(np.int32)
    f = np.where(e > 0, np.ldexp(np.float32(1.0), e - 136), 0).astype(np.float32)
    out = np.ones(px.shape[:2] + (4,), np.float32)
    out[..., :3] = px[..., :3].astype(np.float32) * f[..., None]
    return out


@pytest.mark.parametrize("rle", [True, False])
def test_radiance_hdr_reader(tmp_path, rle):
    """host.load_hdr against an independent numpy RGBE decoder on a synthetic picture, adaptive run-length and flat scanlines."""
    rng = np.random.default_rng(3)
    W, H = 64, 20
    img = (rng.random((H, W, 3)) ** 4 * 50).astype(np.float32)
    img[5:9, 10:40] = [0.25, 0.5, 2.0]   # constant runs
    img[12, :] = 0.0                      # zero exponent
    px = rgbe_encode(img)
    body = bytearray()
    for y in range(H):
        if rle:
            body += bytes([2, 2, W >> 8, W & 255])
            for k in range(4):
                body += rle_channel(px[y, :, k])
        else:
            body += px[y].tobytes()
    p = tmp_path / "sky.hdr"
    p.write_bytes(b"#?RADIANCE\n# made by the test\nFORMAT=32-bit_rle_rgbe\n\n-Y %d +X %d\n" % (H, W) + bytes(body))
    if not rle:  # a flat file must not look like an RLE header by accident
        assert not (px[:, 0, 0] == 2).any()
    got = host.load_hdr(str(p))
    assert got.shape == (H, W, 4) and (got.view(np.uint32) == rgbe_decode(px).view(np.uint32)).all()
    assert np.abs(got[..., :3] - img).max() <= img.max() / 128
    (tmp_path / "bad.hdr").write_bytes(b"P6\n")
    with pytest.raises(ValueError):
        host.load_hdr(str(tmp_path / "bad.hdr"))


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "textures", "river_road_2.hdr")), reason="reference assets not mounted")
def test_reference_sky_hdr_decodes_like_opencv():
    """Sky id 0 of the reference (Engine.cpp:855) through our reader and through OpenCV's Radiance decoder (an independent implementation)."""
    path = os.path.join(REF, "textures", "river_road_2.hdr")
    got = host.load_hdr(path)
    assert got.shape[2] == 4 and got.shape[1] == 2 * got.shape[0] and (got[..., 3] == 1).all() and got[..., :3].max() > 10
    cv2 = pytest.importorskip("cv2")
    ref = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert ref is not None and ref.shape[:2] == got.shape[:2]
    assert np.allclose(got[..., :3], ref[..., ::-1], rtol=1e-6, atol=0)
    sh = host.project_sh(got)  # feeds GPUScene.HDRSHs
    assert np.isfinite(np.ctypeslib.as_array(C.cast(C.byref(sh), C.POINTER(C.c_float)), shape=(27,))).all()
