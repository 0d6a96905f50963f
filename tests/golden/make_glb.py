"""Writes tests/golden/pbr_like.glb: a synthetic glTF 2.0 binary with the LAYOUT of the reference's assets/models/pbr.glb (the declared
stand-in of BASELINE config 3) - a ground cylinder and a 3 x 3 matrix of unit spheres, one mesh per sphere sharing one set of
accessors, ten PBR materials spanning the metallic / roughness matrix, a camera node - generated from scratch with numpy (no reference
bytes), plus what pbr.glb does not exercise: a parent / child node pair, a two-primitive mesh (two sections), 32-bit indices, an interleaved
(strided) vertex buffer view, a transmission material, an emissive material, and sky / sun extras on the scene.
    python tests/golden/make_glb.py"""
import json
import os
import struct

import numpy as np


def uv_sphere(stacks=12, slices=24, radius=1.0):
    pos, nrm, uv = [], [], []
    for i in range(stacks + 1):
        th = np.pi * i / stacks
        for j in range(slices + 1):
            ph = 2 * np.pi * j / slices
            n = np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)])
            pos.append(n * radius); nrm.append(n); uv.append([j / slices, i / stacks])
    idx = []
    for i in range(stacks):
        for j in range(slices):
            a, b = i * (slices + 1) + j, (i + 1) * (slices + 1) + j
            if i > 0: idx += [a, a + 1, b]
            if i < stacks - 1: idx += [a + 1, b + 1, b]
    return np.array(pos, np.float32), np.array(nrm, np.float32), np.array(uv, np.float32), np.array(idx, np.uint32)


def cylinder(radius=10.0, y0=-3.17, y1=0.05, slices=48):
    pos, nrm, uv, idx = [], [], [], []
    for j in range(slices):  # top cap fan + side quads
        a0, a1 = 2 * np.pi * j / slices, 2 * np.pi * (j + 1) / slices
        p0, p1 = [radius * np.cos(a0), y1, radius * np.sin(a0)], [radius * np.cos(a1), y1, radius * np.sin(a1)]
        b = len(pos)
        pos += [[0, y1, 0], p1, p0]; nrm += [[0, 1, 0]] * 3; uv += [[0.5, 0.5], [0.5 + 0.5 * np.cos(a1), 0.5 + 0.5 * np.sin(a1)], [0.5 + 0.5 * np.cos(a0), 0.5 + 0.5 * np.sin(a0)]]
        idx += [b, b + 1, b + 2]
        b = len(pos)
        q0, q1 = [p0[0], y0, p0[2]], [p1[0], y0, p1[2]]
        n0, n1 = [np.cos(a0), 0, np.sin(a0)], [np.cos(a1), 0, np.sin(a1)]
        pos += [p0, p1, q1, q0]; nrm += [n0, n1, n1, n0]; uv += [[j / slices, 0], [(j + 1) / slices, 0], [(j + 1) / slices, 1], [j / slices, 1]]
        idx += [b, b + 1, b + 2, b, b + 2, b + 3]
    return np.array(pos, np.float32), np.array(nrm, np.float32), np.array(uv, np.float32), np.array(idx, np.uint32)


def main():
    bin_chunks, views, accessors = [], [], []

    def add_view(data, stride=None):
        off = sum(len(c) for c in bin_chunks)
        pad = (-len(data)) % 4
        bin_chunks.append(data + b"\0" * pad)
        v = {"buffer": 0, "byteOffset": off, "byteLength": len(data)}
        if stride: v["byteStride"] = stride
        views.append(v)
        return len(views) - 1

    def add_acc(view, ctype, count, typ, offset=0, minmax=None):
        a = {"bufferView": view, "componentType": ctype, "count": int(count), "type": typ}
        if offset: a["byteOffset"] = offset
        if minmax is not None: a["min"], a["max"] = [float(x) for x in minmax[0]], [float(x) for x in minmax[1]]
        accessors.append(a)
        return len(accessors) - 1

    def add_mesh_data(pos, nrm, uv, idx, short=True, interleave=False):
        if interleave:  # one strided view: position | normal | uv = 32 bytes per vertex
            inter = np.concatenate([pos, nrm, uv], axis=1).astype(np.float32)
            v = add_view(inter.tobytes(), stride=32)
            ap = add_acc(v, 5126, len(pos), "VEC3", 0, (pos.min(0), pos.max(0)))
            an = add_acc(v, 5126, len(pos), "VEC3", 12)
            au = add_acc(v, 5126, len(pos), "VEC2", 24)
        else:
            ap = add_acc(add_view(pos.tobytes()), 5126, len(pos), "VEC3", 0, (pos.min(0), pos.max(0)))
            an = add_acc(add_view(nrm.tobytes()), 5126, len(pos), "VEC3")
            au = add_acc(add_view(uv.tobytes()), 5126, len(pos), "VEC2")
        ai = add_acc(add_view(idx.astype(np.uint16 if short else np.uint32).tobytes()), 5123 if short else 5125, len(idx), "SCALAR")
        return {"POSITION": ap, "NORMAL": an, "TEXCOORD_0": au}, ai

    cyl = add_mesh_data(*cylinder())
    sph = add_mesh_data(*uv_sphere())
    small = add_mesh_data(*uv_sphere(6, 10, 0.5), short=False, interleave=True)
    box_pos = np.array([[-1, 0, -1], [1, 0, -1], [1, 0, 1], [-1, 0, 1]], np.float32)
    quad = add_mesh_data(box_pos, np.tile([0, 1, 0], (4, 1)).astype(np.float32), np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32), np.array([0, 2, 1, 0, 3, 2], np.uint32))

    def mat(name, base, metallic=None, roughness=None, **extra):
        p = {"baseColorFactor": base + [1]}
        if metallic is not None: p["metallicFactor"] = metallic
        if roughness is not None: p["roughnessFactor"] = roughness
        m = {"doubleSided": True, "name": name, "pbrMetallicRoughness": p}
        m.update(extra)
        return m

    g3 = [0.3, 0.3, 0.3]
    materials = [mat("Ground", [0.8, 0.8, 0.8], 0), mat("Roughness0.5m05", g3, 0.5, 0.5), mat("Roughness1m05", g3, 0.5), mat("Roughness0m05", g3, 0.5, 0.01),
                 mat("Roughness0.5m0", g3, 0, 0.5), mat("Roughness1m0", g3, 0), mat("Roughness0m0", g3, 0, 0.01), mat("Roughness0.5m1", g3, None, 0.5),
                 mat("Roughness1m1", g3), mat("Roughness0m1", g3, None, 0.01),
                 mat("Glass", [1.0, 1.0, 1.0], 0, 0.0, extensions={"KHR_materials_transmission": {"transmissionFactor": 1.0}, "KHR_materials_ior": {"ior": 1.5}},
                     extras={"ior2": 1.33}),
                 mat("Lamp", [1.0, 1.0, 1.0], 0, 1.0, emissiveFactor=[1.0, 0.9, 0.8], extensions={"KHR_materials_emissive_strength": {"emissiveStrength": 4.0}})]
    meshes = [{"name": "Cylinder", "primitives": [{"attributes": cyl[0], "indices": cyl[1], "material": 0}]}]
    for k in range(9):
        meshes.append({"name": "Sphere", "primitives": [{"attributes": sph[0], "indices": sph[1], "material": 1 + k}]})
    meshes.append({"name": "TwoSections", "primitives": [{"attributes": small[0], "indices": small[1], "material": 10},
                                                          {"attributes": quad[0], "indices": quad[1], "material": 11}]})
    nodes = [{"camera": 0, "name": "Camera", "rotation": [-0.254135, -0.435918, -0.129790, 0.853549], "translation": [-18.2153, 14.5449, 13.0995]},
             {"mesh": 0, "name": "Cylinder"}]
    grid = [(0, 0), (3, 0), (-3, 0), (0, -3), (3, -3), (-3, -3), (0, 3), (3, 3), (-3, 3)]
    for k, (x, z) in enumerate(grid):
        nodes.append({"mesh": 1 + k, "name": f"Sphere.{k:03d}", "translation": [x, 1, z]})
    nodes.append({"name": "Rig", "translation": [6, 2.5, 0], "rotation": [0, 0.38268343, 0, 0.92387953], "scale": [1.5, 1.5, 1.5], "children": [len(nodes) + 1]})
    nodes.append({"mesh": 10, "name": "Hanging", "translation": [0, 0.5, 1]})
    doc = {"asset": {"version": "2.0", "generator": "tests/golden/make_glb.py"}, "scene": 0,
           "extensionsUsed": ["KHR_materials_transmission", "KHR_materials_ior", "KHR_materials_emissive_strength"],
           "scenes": [{"name": "Scene", "nodes": list(range(0, 12)), "extras": {"SkyIdx": 0, "SkyIntensity": 60.0, "WithSun": 1, "SunRotation": 0.75}}],
           "nodes": nodes, "cameras": [{"name": "Camera", "type": "perspective", "perspective": {"aspectRatio": 1.7777, "yfov": 0.4, "zfar": 100, "znear": 0.1}}],
           "materials": materials, "meshes": meshes, "accessors": accessors, "bufferViews": views,
           "buffers": [{"byteLength": sum(len(c) for c in bin_chunks)}]}
    js = json.dumps(doc, separators=(",", ":")).encode()
    js += b" " * ((-len(js)) % 4)
    binary = b"".join(bin_chunks)
    out = struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(binary)) + struct.pack("<II", len(js), 0x4E4F534A) + js + struct.pack("<II", len(binary), 0x004E4942) + binary
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pbr_like.glb")
    with open(path, "wb") as f:
        f.write(out)
    print(path, len(out), "bytes", len(meshes), "meshes", len(nodes), "nodes")


if __name__ == "__main__":
    main()
