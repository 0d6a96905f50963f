// gk_gltf.cpp — the asset path the configs need, from scratch: a GLB (glTF 2.0 binary) reader with the
// reference's scene mapping rules, and a Radiance .hdr (RGBE) reader for the equirect sky.
//
// Mapping rules restated from src/Assets/FSceneLoader.cpp (tinygltf, the reference's parser, is a vcpkg dependency that is not in
// the mount; the container format is the published glTF 2.0 specification):
//   * one Model per glTF mesh; every primitive is a "section" whose vertices carry MaterialIndex = section index (:394-510)
//   * the node stores matId[section] = primitive.material + materialOffset, materialOffset = the number of materials loaded
//     before the file (1: root_default, SceneList.cpp:270) (:147-156); instanceId = node order (:143); children are parsed depth first
//     and hang below their parent (:181-186)
//   * material rule (:312-391): Mixture by default with Fuzziness = roughnessFactor, Metalness = metallicFactor, ior 1.46;
//     Diffuse = sqrt(baseColorFactor); Metallic if metalness > 0.95, Lambertian if roughness > 0.95 and metalness < 0.01 (both only
//     without an MRA texture); KHR_materials_ior, KHR_materials_transmission > 0.8 -> Dielectric, extras.ior2,
//     KHR_materials_emissive_strength -> DiffuseLight(emissive * strength * 100)
//   * a camera node becomes lookAt(eye, eye + forward, +Y) with a FIXED 40 degree field of view (:126-138); without one the
//     AutoFocusCamera looks at the bounds' centre from one diagonal away (:755-784)
//   * sky / sun come from the extras of scene 0 (:517-552)
// Not restated: textures (KTX2 / BC7 / PNG decoding; gkcu_set_sky-style RGBA32F uploads exist for them), Draco-compressed
// primitives (rejected with an error), animations, MikkTSpace tangent generation (tangents only feed normal maps).
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>

#include "gk_scene.hpp"

namespace gk {

// ------------------------------------------------------------------------------------------------ minimal JSON
namespace {

struct Json {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;
    const Json* find(const std::string& k) const {
        for (auto& kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
    bool has(const std::string& k) const { return find(k) != nullptr; }
    const Json& operator[](const std::string& k) const {
        static const Json null;
        const Json* j = find(k);
        return j ? *j : null;
    }
    const Json& operator[](size_t i) const {
        static const Json null;
        return i < arr.size() ? arr[i] : null;
    }
    double number(double dflt) const { return type == Num ? num : dflt; }
    int integer(int dflt) const { return type == Num ? (int)num : dflt; }
    size_t size() const { return type == Arr ? arr.size() : obj.size(); }
};

struct JsonParser {
    const char* p;
    const char* end;
    bool ok = true;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
    }
    bool eat(char c) {
        ws();
        if (p < end && *p == c) { ++p; return true; }
        return false;
    }
    std::string string() {
        std::string s;
        if (!eat('"')) { ok = false; return s; }
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                ++p;
                switch (*p) {
                    case 'n': s += '\n'; break;
                    case 't': s += '\t'; break;
                    case 'r': s += '\r'; break;
                    case 'b': s += '\b'; break;
                    case 'f': s += '\f'; break;
                    case 'u': {  // names only: keep the BMP code point as UTF-8
                        unsigned cp = 0;
                        for (int i = 0; i < 4 && p + 1 < end; ++i) { ++p; cp = cp * 16 + (unsigned)(std::isdigit((unsigned char)*p) ? *p - '0' : (std::tolower((unsigned char)*p) - 'a' + 10)); }
                        if (cp < 0x80) s += (char)cp;
                        else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
                        else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: s += *p;
                }
                ++p;
            } else {
                s += *p++;
            }
        }
        if (p < end) ++p; else ok = false;
        return s;
    }
    Json value() {
        Json j;
        ws();
        if (p >= end) { ok = false; return j; }
        if (*p == '{') {
            ++p;
            j.type = Json::Obj;
            if (eat('}')) return j;
            do {
                std::string k = string();
                if (!eat(':')) { ok = false; return j; }
                j.obj.emplace_back(std::move(k), value());
            } while (ok && eat(','));
            if (!eat('}')) ok = false;
        } else if (*p == '[') {
            ++p;
            j.type = Json::Arr;
            if (eat(']')) return j;
            do j.arr.push_back(value());
            while (ok && eat(','));
            if (!eat(']')) ok = false;
        } else if (*p == '"') {
            j.type = Json::Str;
            j.str = string();
        } else if (!std::strncmp(p, "true", std::min<size_t>(4, end - p))) { j.type = Json::Bool; j.b = true; p += 4; }
        else if (!std::strncmp(p, "false", std::min<size_t>(5, end - p))) { j.type = Json::Bool; j.b = false; p += 5; }
        else if (!std::strncmp(p, "null", std::min<size_t>(4, end - p))) { p += 4; }
        else {
            char* e = nullptr;
            j.num = std::strtod(p, &e);
            if (e == p) { ok = false; return j; }
            j.type = Json::Num;
            p = e;
        }
        return j;
    }
};

struct Accessor {
    const uint8_t* data = nullptr;
    size_t count = 0, stride = 0;
    int componentType = 0;
    int components = 0;
};

struct Glb {
    Json doc;
    std::vector<uint8_t> bin;
    std::string error;

    bool accessor(int idx, Accessor& out) {
        const Json& a = doc["accessors"][(size_t)idx];
        if (idx < 0 || a.type != Json::Obj) { error = "accessor index out of range"; return false; }
        if (a.has("sparse")) { error = "sparse accessors are not supported"; return false; }
        const Json& v = doc["bufferViews"][(size_t)a["bufferView"].integer(-1)];
        if (v.type != Json::Obj || v["buffer"].integer(0) != 0) { error = "accessor without a buffer view in the BIN chunk"; return false; }
        const std::string& type = a["type"].str;
        out.components = type == "SCALAR" ? 1 : type == "VEC2" ? 2 : type == "VEC3" ? 3 : type == "VEC4" ? 4 : 0;
        out.componentType = a["componentType"].integer(0);
        const size_t csize = (out.componentType == 5120 || out.componentType == 5121) ? 1 : (out.componentType == 5122 || out.componentType == 5123) ? 2 : 4;
        out.count = (size_t)a["count"].number(0);
        out.stride = (size_t)v["byteStride"].number(0);
        if (out.stride == 0) out.stride = csize * (size_t)out.components;  // tinygltf Accessor::ByteStride
        const size_t off = (size_t)v["byteOffset"].number(0) + (size_t)a["byteOffset"].number(0);
        if (out.components == 0 || (out.count && off + (out.count - 1) * out.stride + csize * (size_t)out.components > bin.size())) {
            error = "accessor exceeds the BIN chunk";
            return false;
        }
        out.data = bin.data() + off;
        return true;
    }
};

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    f.seekg(0, std::ios::end);
    const std::streamoff n = f.tellg();
    f.seekg(0);
    out.resize((size_t)n);
    f.read(reinterpret_cast<char*>(out.data()), n);
    return (bool)f;
}

uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

bool parse_glb(const std::vector<uint8_t>& file, Glb& g) {
    if (file.size() < 20 || rd32(file.data()) != 0x46546C67u) { g.error = "not a GLB file (magic)"; return false; }
    if (rd32(file.data() + 4) != 2) { g.error = "GLB container version is not 2"; return false; }
    size_t pos = 12;
    bool have_json = false;
    while (pos + 8 <= file.size()) {
        const uint32_t len = rd32(file.data() + pos), type = rd32(file.data() + pos + 4);
        pos += 8;
        if (pos + len > file.size()) { g.error = "GLB chunk exceeds the file"; return false; }
        if (type == 0x4E4F534Au) {  // JSON
            JsonParser jp{reinterpret_cast<const char*>(file.data() + pos), reinterpret_cast<const char*>(file.data() + pos + len)};
            g.doc = jp.value();
            if (!jp.ok || g.doc.type != Json::Obj) { g.error = "GLB JSON chunk does not parse"; return false; }
            have_json = true;
        } else if (type == 0x004E4942u) {  // BIN
            g.bin.assign(file.begin() + (std::ptrdiff_t)pos, file.begin() + (std::ptrdiff_t)(pos + len));
        }
        pos += (len + 3u) & ~3u;
    }
    if (!have_json) { g.error = "GLB without a JSON chunk"; return false; }
    return true;
}

float f32_at(const uint8_t* p) {
    float f;
    std::memcpy(&f, p, 4);
    return f;
}

struct NodeCtx {
    Scene& s;
    Glb& g;
    uint32_t modelIdx, materialOffset;
    std::vector<Camera> cameras;
};

// ParseGltfNode, FSceneLoader.cpp:102-187
void parse_node(NodeCtx& c, int nodeIdx, const mat4* parentWorld) {
    const Json& node = c.g.doc["nodes"][(size_t)nodeIdx];
    auto vec = [&](const char* key, int n, double dflt, double* out) {
        const Json& a = node[key];
        for (int i = 0; i < n; ++i) out[i] = a.type == Json::Arr && a.size() == (size_t)n ? a[(size_t)i].number(dflt) : dflt;
    };
    double t[3], sc[3], r[4];
    vec("translation", 3, 0.0, t);
    vec("scale", 3, 1.0, sc);
    const Json& ra = node["rotation"];
    if (ra.type == Json::Arr && ra.size() == 4) for (int i = 0; i < 4; ++i) r[i] = ra[(size_t)i].number(0);
    else { r[0] = r[1] = r[2] = 0; r[3] = 1; }
    const vec3 translation((float)t[0], (float)t[1], (float)t[2]), scaling((float)sc[0], (float)sc[1], (float)sc[2]);
    quat rotation{(float)r[3], (float)r[0], (float)r[1], (float)r[2]};
    const float ql = std::sqrt(rotation.w * rotation.w + rotation.x * rotation.x + rotation.y * rotation.y + rotation.z * rotation.z);
    if (ql > 0) { rotation.w /= ql; rotation.x /= ql; rotation.y /= ql; rotation.z /= ql; }  // glm::normalize
    const int mesh = node["mesh"].integer(-1), camera = node["camera"].integer(-1);
    uint32_t meshId = 0xFFFFFFFFu;
    const mat4 local = translate(mat4(1.0f), translation) * mat4_cast(rotation) * scale(mat4(1.0f), scaling);
    if (mesh != -1) meshId = (uint32_t)mesh + c.modelIdx;
    else if (camera >= 0) {
        const vec4 eye = local * vec4(0, 0, 0, 1), fwd = local * vec4(0, 0, -1, 0);
        Camera cam;
        cam.ModelView = look_at_rh(vec3(eye.x, eye.y, eye.z), vec3(eye.x + fwd.x, eye.y + fwd.y, eye.z + fwd.z), vec3(0, 1, 0));
        cam.FieldOfView = 40;
        c.cameras.push_back(cam);
    }
    Node n;
    n.name = node["name"].str;
    n.translation = translation;
    n.rotation = rotation;
    n.scaling = scaling;
    n.modelId = meshId;
    n.instanceId = (uint32_t)c.s.nodes.size();
    n.transform = parentWorld ? (*parentWorld) * local : local;  // Node::RecalcTransform with a parent (Node.cpp:31-55)
    n.prevTransform = translate(mat4(1.0f), vec3(0, -100, 0));   // CreateNode(..., replace = false), Node.cpp:154-171
    n.visible = false;
    if (meshId != 0xFFFFFFFFu) {
        n.visible = true;
        const Json& prims = c.g.doc["meshes"][(size_t)mesh]["primitives"];
        for (size_t i = 0; i < prims.size() && i < 16; ++i)
            n.materialIdx[i] = (uint32_t)std::max(0, prims[i]["material"].integer(-1) + (int)c.materialOffset);
    }
    const mat4 world = n.transform;
    c.s.nodes.push_back(n);
    const Json& children = node["children"];
    for (size_t i = 0; i < children.size(); ++i) parse_node(c, children[i].integer(0), &world);
}

}  // namespace

// FSceneLoader::LoadGLTFScene, FSceneLoader.cpp:210-640 (geometry, materials, nodes, cameras, environment)
bool LoadGlbScene(Scene& s, const std::string& path, std::string& error) {
    std::vector<uint8_t> file;
    if (!read_file(path, file)) { error = "cannot read " + path; return false; }
    Glb g;
    if (!parse_glb(file, g)) { error = g.error; return false; }
    const uint32_t materialOffset = (uint32_t)s.materials.size(), modelIdx = (uint32_t)s.models.size();
    const Json& used = g.doc["extensionsRequired"];
    for (size_t i = 0; i < used.size(); ++i)
        if (used[i].str == "KHR_draco_mesh_compression" || used[i].str == "EXT_meshopt_compression") {
            error = "compressed geometry (" + used[i].str + ") is not supported";
            return false;
        }
    // ---- materials (:312-391)
    const Json& mats = g.doc["materials"];
    for (size_t i = 0; i < mats.size(); ++i) {
        const Json& mat = mats[i];
        const Json& pbr = mat["pbrMetallicRoughness"];
        GkMaterial m{};
        m.DiffuseTextureId = -1;  // textures are not loaded by this reader (see the header comment)
        m.MRATextureId = -1;
        m.NormalTextureId = -1;
        m.NormalTextureScale = 1.0f;
        m.MaterialModel = GK_MATERIAL_MIXTURE;
        m.Fuzziness = (float)pbr["roughnessFactor"].number(1.0);  // glTF defaults: roughness 1, metallic 1, base colour 1
        m.Metalness = (float)pbr["metallicFactor"].number(1.0);
        m.RefractionIndex = 1.46f;
        m.RefractionIndex2 = 1.46f;
        const Json& bc = pbr["baseColorFactor"];
        vec3 diffuse(1, 1, 1);
        if (bc.type == Json::Arr && bc.size() >= 3) diffuse = vec3((float)bc[(size_t)0].number(1), (float)bc[(size_t)1].number(1), (float)bc[(size_t)2].number(1));
        m.Diffuse = {std::sqrt(diffuse.x), std::sqrt(diffuse.y), std::sqrt(diffuse.z), 1.0f};
        if (m.Metalness > .95) m.MaterialModel = GK_MATERIAL_METALLIC;
        if (m.Fuzziness > .95 && m.Metalness < 0.01) m.MaterialModel = GK_MATERIAL_LAMBERTIAN;
        const Json& ext = mat["extensions"];
        if (ext.has("KHR_materials_ior")) m.RefractionIndex = m.RefractionIndex2 = (float)ext["KHR_materials_ior"]["ior"].number(1.5);
        if (ext.has("KHR_materials_transmission") && (float)ext["KHR_materials_transmission"]["transmissionFactor"].number(0) > 0.8f)
            m.MaterialModel = GK_MATERIAL_DIELECTRIC;
        if (mat["extras"].has("ior2")) m.RefractionIndex2 = (float)mat["extras"]["ior2"].number(m.RefractionIndex2);
        if (ext.has("KHR_materials_emissive_strength")) {
            const Json& ef = mat["emissiveFactor"];
            vec3 e(0, 0, 0);
            if (ef.type == Json::Arr && ef.size() >= 3) e = vec3((float)ef[(size_t)0].number(0), (float)ef[(size_t)1].number(0), (float)ef[(size_t)2].number(0));
            const float power = (float)ext["KHR_materials_emissive_strength"]["emissiveStrength"].number(1.0);
            m = MaterialDiffuseLight(e * power * 100.0f);
        }
        s.materials.push_back(m);
    }
    // ---- meshes -> models (:394-510)
    const Json& meshes = g.doc["meshes"];
    for (size_t mi = 0; mi < meshes.size(); ++mi) {
        Model model;
        model.name = meshes[mi]["name"].str;
        uint32_t vertexOffset = 0, sectionIdx = 0;
        const Json& prims = meshes[mi]["primitives"];
        for (size_t pi = 0; pi < prims.size(); ++pi) {
            const Json& prim = prims[pi];
            if (prim["extensions"].has("KHR_draco_mesh_compression")) { error = "Draco-compressed primitive in " + path; return false; }
            Accessor idx, pos, nrm, uv, tan;
            if (prim["mode"].integer(4) != 4 || !prim.has("indices")) continue;  // TINYGLTF_MODE_TRIANGLES only
            if (!g.accessor(prim["indices"].integer(-1), idx)) { error = g.error; return false; }
            if (idx.count == 0) continue;
            const Json& at = prim["attributes"];
            if (!at.has("POSITION") || !g.accessor(at["POSITION"].integer(-1), pos)) { error = "primitive without POSITION: " + g.error; return false; }
            const bool hasN = at.has("NORMAL") && g.accessor(at["NORMAL"].integer(-1), nrm);
            const bool hasT = at.has("TEXCOORD_0") && g.accessor(at["TEXCOORD_0"].integer(-1), uv) && uv.componentType == 5126;
            const bool hasTan = at.has("TANGENT") && g.accessor(at["TANGENT"].integer(-1), tan);
            if (pos.componentType != 5126 || (hasN && nrm.componentType != 5126)) { error = "non-float vertex attributes are not supported"; return false; }
            for (size_t i = 0; i < pos.count; ++i) {
                Vertex v;
                const uint8_t* p = pos.data + i * pos.stride;
                v.Position = vec3(f32_at(p), f32_at(p + 4), f32_at(p + 8));
                if (hasN && i < nrm.count) {
                    const uint8_t* q = nrm.data + i * nrm.stride;
                    v.Normal = vec3(f32_at(q), f32_at(q + 4), f32_at(q + 8));
                }
                if (hasTan && i < tan.count) {
                    const uint8_t* q = tan.data + i * tan.stride;
                    v.Tangent = vec4(f32_at(q), f32_at(q + 4), f32_at(q + 8), f32_at(q + 12));
                }
                if (hasT && i < uv.count) {
                    const uint8_t* q = uv.data + i * uv.stride;
                    v.TexCoord = {f32_at(q), f32_at(q + 4)};
                }
                v.MaterialIndex = sectionIdx;
                model.vertices.push_back(v);
            }
            sectionIdx++;
            for (size_t i = 0; i < idx.count; ++i) {
                const uint8_t* q = idx.data + i * idx.stride;
                uint32_t k;
                if (idx.componentType == 5123) k = (uint32_t)q[0] | ((uint32_t)q[1] << 8);
                else if (idx.componentType == 5125 || idx.componentType == 5124) k = rd32(q);
                else if (idx.componentType == 5121) k = q[0];
                else { error = "unsupported index component type"; return false; }
                if (k >= pos.count) { error = "index beyond the primitive's vertices"; return false; }
                model.indices.push_back(k + vertexOffset);
            }
            vertexOffset += (uint32_t)pos.count;
        }
        model.aabb_min = vec3(3.4e38f);
        model.aabb_max = vec3(-3.4e38f);
        for (auto& v : model.vertices) {
            model.aabb_min = {std::min(model.aabb_min.x, v.Position.x), std::min(model.aabb_min.y, v.Position.y), std::min(model.aabb_min.z, v.Position.z)};
            model.aabb_max = {std::max(model.aabb_max.x, v.Position.x), std::max(model.aabb_max.y, v.Position.y), std::max(model.aabb_max.z, v.Position.z)};
        }
        s.models.push_back(std::move(model));
    }
    // ---- environment from the extras of scene 0 (:517-552)
    const Json& scene0 = g.doc["scenes"][(size_t)0];
    const Json& ex = scene0["extras"];
    if (ex.has("SkyIdx")) { s.env.HasSky = true; s.env.SkyIdx = ex["SkyIdx"].integer(0); }
    if (ex.has("SkyIntensity")) { s.env.HasSky = true; s.env.SkyIntensity = (float)ex["SkyIntensity"].number(100); }
    if (ex.has("SkyRotation")) { s.env.HasSky = true; s.env.SkyRotation = (float)ex["SkyRotation"].number(0); }
    if (ex.has("SunIntensity")) { s.env.HasSun = true; s.env.SunIntensity = (float)ex["SunIntensity"].number(500); }
    if (ex.has("WithSun")) s.env.HasSun = ex["WithSun"].integer(0) != 0;
    if (ex.has("SunRotation")) s.env.SunRotation = (float)ex["SunRotation"].number(0.5);
    if (ex.has("NoSky")) s.env.HasSky = false;
    // ---- nodes, depth first from the roots of scene 0 (:554-566)
    NodeCtx ctx{s, g, modelIdx, materialOffset, {}};
    const Json& roots = scene0["nodes"];
    for (size_t i = 0; i < roots.size(); ++i) parse_node(ctx, roots[i].integer(0), nullptr);
    if (!ctx.cameras.empty()) {
        s.camera = ctx.cameras[0];
    } else {  // AutoFocusCamera, :755-784
        vec3 mn(3.4e38f), mx(-3.4e38f);
        for (size_t i = modelIdx; i < s.models.size(); ++i) {
            mn = {std::min(mn.x, s.models[i].aabb_min.x), std::min(mn.y, s.models[i].aabb_min.y), std::min(mn.z, s.models[i].aabb_min.z)};
            mx = {std::max(mx.x, s.models[i].aabb_max.x), std::max(mx.y, s.models[i].aabb_max.y), std::max(mx.z, s.models[i].aabb_max.z)};
        }
        const vec3 centre = (mx - mn) * 0.5f + mn;
        s.camera = Camera();
        s.camera.ModelView = look_at_rh(vec3(centre.x, centre.y, centre.z + length(mx - mn)), centre, vec3(0, 1, 0));
        s.camera.FieldOfView = 40;
        s.camera.Aperture = 0.0f;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ Radiance .hdr
// The equirect sky textures are Radiance RGBE pictures (assets/textures/*.hdr, loaded through stb_image's stbi_loadf at
// src/Assets/Texture.cpp:390-432 into RGBA32F). The format is public (Ward, "Real Pixels", Graphics Gems II): a text header ending
// in an empty line, "-Y h +X w", then per scanline either flat RGBE or the adaptive run-length code (2 2 hi lo, then each of the four
// channels as runs). A pixel is (r, g, b) * 2^(e - 136) with the mantissas as integers - stb's hdr_convert, alpha = 1.
bool LoadHdrImage(const std::string& path, std::vector<float>& rgba, uint32_t& width, uint32_t& height, std::string& error) {
    std::vector<uint8_t> f;
    if (!read_file(path, f)) { error = "cannot read " + path; return false; }
    size_t pos = 0;
    auto line = [&]() {
        std::string s;
        while (pos < f.size() && f[pos] != '\n') s += (char)f[pos++];
        if (pos < f.size()) ++pos;
        return s;
    };
    const std::string magic = line();
    if (magic != "#?RADIANCE" && magic != "#?RGBE") { error = "not a Radiance .hdr file"; return false; }
    bool format = false;
    for (;;) {
        const std::string l = line();
        if (l.empty()) break;
        if (l == "FORMAT=32-bit_rle_rgbe") format = true;
        if (pos >= f.size()) { error = "truncated .hdr header"; return false; }
    }
    if (!format) { error = "unsupported .hdr FORMAT"; return false; }
    int h = 0, w = 0;
    const std::string res = line();
    if (std::sscanf(res.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) { error = "unsupported .hdr orientation: " + res; return false; }
    width = (uint32_t)w;
    height = (uint32_t)h;
    rgba.assign((size_t)w * h * 4, 1.0f);
    std::vector<uint8_t> scan((size_t)w * 4);
    auto convert = [&](const uint8_t* in, float* out) {
        if (in[3] != 0) {
            const float f1 = std::ldexp(1.0f, (int)in[3] - (128 + 8));
            out[0] = in[0] * f1; out[1] = in[1] * f1; out[2] = in[2] * f1;
        } else {
            out[0] = out[1] = out[2] = 0.0f;
        }
        out[3] = 1.0f;
    };
    for (int y = 0; y < h; ++y) {
        if (pos + 4 > f.size()) { error = "truncated .hdr data"; return false; }
        const bool rle = w >= 8 && w < 32768 && f[pos] == 2 && f[pos + 1] == 2 && !(f[pos + 2] & 0x80) && (((int)f[pos + 2] << 8) | f[pos + 3]) == w;
        if (!rle) {  // flat scanline
            if (pos + (size_t)w * 4 > f.size()) { error = "truncated .hdr data"; return false; }
            for (int x = 0; x < w; ++x) convert(&f[pos + (size_t)x * 4], &rgba[((size_t)y * w + x) * 4]);
            pos += (size_t)w * 4;
            continue;
        }
        pos += 4;
        for (int k = 0; k < 4; ++k) {
            int x = 0;
            while (x < w) {
                if (pos >= f.size()) { error = "truncated .hdr data"; return false; }
                int count = f[pos++];
                if (count > 128) {  // run
                    count -= 128;
                    if (pos >= f.size() || x + count > w) { error = "corrupt .hdr run"; return false; }
                    const uint8_t v = f[pos++];
                    for (int i = 0; i < count; ++i) scan[(size_t)(x++) * 4 + k] = v;
                } else {  // dump
                    if (count == 0 || pos + (size_t)count > f.size() || x + count > w) { error = "corrupt .hdr dump"; return false; }
                    for (int i = 0; i < count; ++i) scan[(size_t)(x++) * 4 + k] = f[pos++];
                }
            }
        }
        for (int x = 0; x < w; ++x) convert(&scan[(size_t)x * 4], &rgba[((size_t)y * w + x) * 4]);
    }
    return true;
}

}  // namespace gk
