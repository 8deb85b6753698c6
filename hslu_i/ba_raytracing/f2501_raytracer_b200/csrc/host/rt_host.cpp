// rt_host.cpp -- host-side mirror of the reference's scene construction (see rt_host.hpp).
#include "rt_host.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

#include "../rt_internal.h"
#include "rt_host.h"

namespace rth {

static const float EPSILON = 1.1920929e-7f;

// ---------------------------------------------------------------- src/lib.rs:73-92
SceneConstants::SceneConstants(uint32_t w, uint32_t h) : window_width(w), window_height(h) {
    aspect = (float)h / (float)w;
    scene_width = 1.0f;
    scene_height = scene_width * aspect;
    scene_depth = (scene_width + scene_height) / 2.0f;
    average_scene_dimension = (scene_width + scene_height + scene_depth) / 3.0f;
    w2s_w = scene_width / (float)w;
    w2s_h = scene_height / (float)h;
    w2s_d = scene_depth / (float)((w + h) / 2);
    render_ray_focus = Vec3(scene_width / 2.0f, scene_height / 2.0f, -1.9f * scene_depth);
}

// ---------------------------------------------------------------- Cargo.toml:62-83
bool Features::parse(const std::string &csv, Features &f, std::string &err) {
    f = Features();
    std::vector<std::string> todo;
    std::stringstream ss(csv);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        tok.erase(std::remove_if(tok.begin(), tok.end(), [](char c) { return c == ' ' || c == '\t'; }), tok.end());
        if (!tok.empty()) todo.push_back(tok);
    }
    while (!todo.empty()) {
        std::string t = todo.back();
        todo.pop_back();
        auto imply = [&](std::initializer_list<const char *> l) {
            for (auto s : l) todo.push_back(s);
        };
        if (t == "default")
            imply({"realistic", "save_rendering_image", "scene_backface_culling", "anti_aliasing_randomness",
                   "anti_aliasing_rotation_scale", "medium_resolution", "high_quality"});
        else if (t == "simd_render") f.simd_render = true;
        else if (t == "simulate_slow_render" || t == "render_timing_debug") { /* not on the render path */ }
        else if (t == "anti_aliasing") f.anti_aliasing = true;
        else if (t == "anti_aliasing_rotation_scale") { f.anti_aliasing_rotation_scale = true; imply({"anti_aliasing"}); }
        else if (t == "anti_aliasing_randomness") { f.anti_aliasing_randomness = true; imply({"anti_aliasing"}); }
        else if (t == "high_resolution") f.high_resolution = true;
        else if (t == "medium_resolution") f.medium_resolution = true;
        else if (t == "soft_shadows") f.soft_shadows = true;
        else if (t == "reflections") f.reflections = true;
        else if (t == "light_reflections") f.light_reflections = true;
        else if (t == "refractions") f.refractions = true;
        else if (t == "backface_culling") f.backface_culling = true;
        else if (t == "scene_backface_culling") f.scene_backface_culling = true;
        else if (t == "save_rendering_image") f.save_rendering_image = true;
        else if (t == "realistic") { f.realistic = true; imply({"reflections", "light_reflections", "refractions"}); }
        else if (t == "high_quality_model") f.high_quality_model = true;
        else if (t == "high_quality") { f.high_quality = true; imply({"anti_aliasing", "soft_shadows", "high_quality_model"}); }
        else if (t == "extreme_quality") { f.extreme_quality = true; imply({"high_quality"}); }
        else {
            err = "unknown feature '" + t + "'";
            return false;
        }
    }
    return true;
}

// Output::get_feature_string (reference src/output/mod.rs:25-88): the text the window title carries.  The inner
// format!("Antialiasing {} {}") is trimmed as a whole (so "ROS_SCL" alone keeps no trailing blank, "RNG" alone no
// leading one); the depth in the resolution triple is WINDOW_SCENE_DEPTH = (W + H) / 2 (src/lib.rs:74).
std::string Features::feature_string() const {
    auto trim = [](std::string t) {
        const size_t a = t.find_first_not_of(' ');
        if (a == std::string::npos) return std::string();
        return t.substr(a, t.find_last_not_of(' ') - a + 1);
    };
    std::string out = simd_render ? "SIMD" : "Non-SIMD";
    out += " | ";
    if (anti_aliasing)
        out += trim(std::string("Antialiasing ") + (anti_aliasing_rotation_scale ? "ROS_SCL" : "") + " " + (anti_aliasing_randomness ? "RNG" : ""));
    else
        out += "Non-Antialiasing";
    out += " | ";
    out += (reflections || refractions) ? (!refractions ? "Reflections" : "Reflections + Refractions") : "Non-Realistic";
    out += " | ";
    out += high_quality ? (extreme_quality ? "Extreme Quality" : "High Quality") : "Standard Quality";
    out += " | ";
    out += high_resolution ? "High Resolution" : (medium_resolution ? "Medium Resolution" : "Small Resolution");
    const uint32_t W = window_width(), H = window_height();
    out += " (" + std::to_string(W) + "\xC3\x97" + std::to_string(H) + "\xC3\x97" + std::to_string((W + H) / 2) + ") | ";  // U+00D7
    out += backface_culling ? "Backface Culling" : "NO-OPT";
    return trim(out);
}

uint32_t Features::render_flags() const {
    uint32_t fl = 0;
    if (anti_aliasing) fl |= RT_FLAG_ANTI_ALIASING;
    if (anti_aliasing_rotation_scale) fl |= RT_FLAG_AA_ROTATION_SCALE;
    if (reflections) fl |= RT_FLAG_REFLECTIONS;
    if (refractions) fl |= RT_FLAG_REFRACTIONS;
    if (backface_culling) fl |= RT_FLAG_BACKFACE_CULLING;
    if (simd_render) fl |= RT_FLAG_PACKET_SEMANTICS;
    return fl;
}

// ---------------------------------------------------------------- materials
rt_material Material::flat() const {
    rt_material m;
    m.color[0] = color.x;
    m.color[1] = color.y;
    m.color[2] = color.z;
    m.metallic = metallic;
    m.shininess = shininess;
    m.ior = transmission.refraction_index;
    m.opacity = transmission.opacity;
    m.has_opacity = transmission.has_opacity ? 1u : 0u;
    m.boost = transmission.boost;
    return m;
}

static Material material_from_flat(const rt_material *m) {
    Material r;
    r.color = ColorType(m->color[0], m->color[1], m->color[2]);
    r.metallic = m->metallic;
    r.shininess = m->shininess;
    r.transmission.refraction_index = m->ior;
    r.transmission.opacity = m->opacity;
    r.transmission.has_opacity = m->has_opacity != 0;
    r.transmission.boost = m->boost;
    return r;
}

// src/color.rs:125-131: linear -> sRGB encoding -> HSV, value := 1, -> sRGB -> linear
static float srgb_encode(float x) { return x <= 0.0031308f ? 12.92f * x : 1.055f * std::pow(x, 1.0f / 2.4f) - 0.055f; }
static float srgb_decode(float x) { return x <= 0.04045f ? x / 12.92f : std::pow((x + 0.055f) / 1.055f, 2.4f); }
ColorType maximize_value(ColorType c) {
    float r = srgb_encode(c.x), g = srgb_encode(c.y), b = srgb_encode(c.z);
    float mx = std::max(r, std::max(g, b)), mn = std::min(r, std::min(g, b));
    float delta = mx - mn;
    float sat = mx > 0.0f ? delta / mx : 0.0f;
    float hue = 0.0f;  // degrees
    if (delta > 0.0f) {
        if (mx == r) hue = 60.0f * std::fmod((g - b) / delta, 6.0f);
        else if (mx == g) hue = 60.0f * ((b - r) / delta + 2.0f);
        else hue = 60.0f * ((r - g) / delta + 4.0f);
        if (hue < 0.0f) hue += 360.0f;
    }
    float v = 1.0f;
    float ch = v * sat;
    float hp = hue / 60.0f;
    float xx = ch * (1.0f - std::fabs(std::fmod(hp, 2.0f) - 1.0f));
    float r1 = 0, g1 = 0, b1 = 0;
    if (hp < 1) { r1 = ch; g1 = xx; }
    else if (hp < 2) { r1 = xx; g1 = ch; }
    else if (hp < 3) { g1 = ch; b1 = xx; }
    else if (hp < 4) { g1 = xx; b1 = ch; }
    else if (hp < 5) { r1 = xx; b1 = ch; }
    else { r1 = ch; b1 = xx; }
    float m = v - ch;
    return ColorType(srgb_decode(r1 + m), srgb_decode(g1 + m), srgb_decode(b1 + m));
}

// ---------------------------------------------------------------- BoundedPlane
bool BoundedPlane::with_material(Vec3 normal, Vec3 center, Vec3 up, float width, float height, float depth, Material m,
                                 BoundedPlane &out, std::string &err) {
    if (!(width > 0.0f)) { err = "width must be positive"; return false; }
    if (!(height > 0.0f)) { err = "height must be positive"; return false; }
    if (!(std::fabs(dot(normal, up) - 0.0f) <= EPSILON)) { err = "up must be orthogonal to normal"; return false; }
    out.center = center;
    out.up = up;
    out.left = normalized(cross(normal, up));
    out.normal = normal;
    out.width = width;
    out.height = height;
    out.depth = depth;
    out.material = m;
    return true;
}

void BoundedPlane::triangulate(Vec3 t1[3], Vec3 t2[3]) const {
    Vec3 x = (-left) * (width / 2.0f);
    Vec3 y = up * (height / 2.0f);
    Vec3 c = center;
    Vec3 p0 = -x + y, p1 = x + y, p2 = -x - y, p3 = x - y;
    t1[0] = c + p1; t1[1] = c + p0; t1[2] = c + p3;
    t2[0] = c + p2; t2[1] = c + p3; t2[2] = c + p0;
}

std::vector<TriangleData> BoundedPlane::to_basic_geometries() const {
    Vec3 t1[3], t2[3];
    triangulate(t1, t2);
    std::vector<TriangleData> tris;
    Vec3 dod = normal;
    const float half = 0.5f;
    struct FB { float off; Vec3 n; };
    FB fb[2] = {{-(depth * half), -normal}, {depth * half, normal}};
    for (auto &p : fb) {
        Vec3 offset = dod * p.off;
        tris.push_back(TriangleData::with_material_and_normal(t1[0] + offset, t1[1] + offset, t1[2] + offset, p.n, material));
        tris.push_back(TriangleData::with_material_and_normal(t2[0] + offset, t2[1] + offset, t2[2] + offset, p.n, material));
    }
    struct Side { Vec3 dir; float dir_offset; float width; Vec3 n; };
    Side sides[4] = {{up, height, width, up}, {left, width, height, left}, {-up, height, width, -up}, {-left, width, height, -left}};
    for (auto &sd : sides) {
        Vec3 plate_center = mul_add(sd.dir, Vec3(sd.dir_offset * half, sd.dir_offset * half, sd.dir_offset * half), center);
        BoundedPlane plate;
        // the reference would panic here if an assertion failed; the plates of a valid plane never do
        // except for numerically non-orthogonal input, in which case the plate is built unchecked
        plate.center = plate_center;
        plate.up = dod;
        plate.left = normalized(cross(sd.n, dod));
        plate.normal = sd.n;
        plate.width = sd.width;
        plate.height = depth;
        plate.depth = 0.0f;
        plate.material = material;
        Vec3 a[3], b[3];
        plate.triangulate(a, b);
        tris.push_back(TriangleData::with_material_and_normal(a[0], a[1], a[2], sd.n, material));
        tris.push_back(TriangleData::with_material_and_normal(b[0], b[1], b[2], sd.n, material));
    }
    return tris;
}

// ---------------------------------------------------------------- OBJ / MTL (tobj 4.0.3 subset)
static Material material_from_mtl(const std::map<std::string, std::string> &kv) {
    // impl From<&tobj::Material> for Material<f32> (src/raytracing/material.rs:96-126)
    auto getf = [&](const char *k, float d) {
        auto it = kv.find(k);
        if (it == kv.end()) return d;
        char *end = nullptr;
        float v = std::strtof(it->second.c_str(), &end);
        return end == it->second.c_str() ? d : v;
    };
    int illum = 0;
    if (kv.count("illum")) illum = std::atoi(kv.at("illum").c_str());
    ColorType color(0, 0, 0);
    if (kv.count("Kd")) {
        std::stringstream ss(kv.at("Kd"));
        ss >> color.x >> color.y >> color.z;
    }
    float metallic = illum == 3 ? getf("Pm", 0.0f) : 0.0f;
    float shininess = (illum == 3 || illum == 2 || illum == 0) ? getf("Ps", 0.0f) : 0.0f;
    return Material(color, metallic, shininess, TransmissionProperties());
}

static bool load_mtl(const std::string &path, std::vector<Material> &mats, std::map<std::string, int> &names) {
    std::ifstream in(path);
    if (!in) return false;
    std::string line, cur;
    std::map<std::string, std::string> kv;
    auto flush = [&]() {
        if (!cur.empty()) {
            names[cur] = (int)mats.size();
            mats.push_back(material_from_mtl(kv));
        }
        kv.clear();
    };
    while (std::getline(in, line)) {
        std::stringstream ss(line);
        std::string key;
        if (!(ss >> key) || key[0] == '#') continue;
        std::string rest;
        std::getline(ss, rest);
        size_t b = rest.find_first_not_of(" \t\r");
        size_t e = rest.find_last_not_of(" \t\r");
        rest = b == std::string::npos ? "" : rest.substr(b, e - b + 1);
        if (key == "newmtl") {
            flush();
            cur = rest;
        } else {
            kv[key] = rest;
        }
    }
    flush();
    return true;
}

bool load_obj(const std::string &path, Mesh &out, std::string &err) {
    std::ifstream in(path);
    if (!in) {
        err = "cannot open " + path;
        return false;
    }
    out = Mesh();
    std::map<std::string, int> mat_names;
    int cur_mat = -1;
    std::string dir = path.substr(0, path.find_last_of('/') == std::string::npos ? 0 : path.find_last_of('/') + 1);
    std::string line;
    bool any_vn_missing = false, any_vn = false;
    while (std::getline(in, line)) {
        if (line.size() < 2) continue;
        const char *p = line.c_str();
        if (p[0] == 'v' && p[1] == ' ') {
            float x, y, z;
            if (sscanf(p + 2, "%f %f %f", &x, &y, &z) == 3) { out.positions.push_back(x); out.positions.push_back(y); out.positions.push_back(z); }
        } else if (p[0] == 'v' && p[1] == 'n' && p[2] == ' ') {
            float x, y, z;
            if (sscanf(p + 3, "%f %f %f", &x, &y, &z) == 3) { out.normals.push_back(x); out.normals.push_back(y); out.normals.push_back(z); }
        } else if (p[0] == 'f' && p[1] == ' ') {
            std::vector<int> vi, ni;
            std::stringstream ss(p + 2);
            std::string corner;
            while (ss >> corner) {
                int v = 0, t = 0, n = 0;
                if (sscanf(corner.c_str(), "%d/%d/%d", &v, &t, &n) == 3) {}
                else if (sscanf(corner.c_str(), "%d//%d", &v, &n) == 2) {}
                else if (sscanf(corner.c_str(), "%d/%d", &v, &t) == 2) { n = 0; }
                else if (sscanf(corner.c_str(), "%d", &v) == 1) { n = 0; }
                else continue;
                int nv = (int)out.positions.size() / 3, nn = (int)out.normals.size() / 3;
                vi.push_back(v < 0 ? nv + v : v - 1);
                ni.push_back(n == 0 ? -1 : (n < 0 ? nn + n : n - 1));
            }
            for (size_t k = 1; k + 1 < vi.size(); k++) {  // fan triangulation
                int idx[3] = {0, (int)k, (int)k + 1};
                for (int c : idx) {
                    out.face_v.push_back(vi[c]);
                    out.face_vn.push_back(ni[c]);
                    if (ni[c] < 0) any_vn_missing = true; else any_vn = true;
                }
                out.face_material.push_back(cur_mat);
            }
        } else if (!strncmp(p, "mtllib ", 7)) {
            std::string name(p + 7);
            name.erase(name.find_last_not_of(" \t\r") + 1);
            load_mtl(dir + name, out.materials, mat_names);  // CONTINUE_ON_MATERIAL_FAILURE = true
        } else if (!strncmp(p, "usemtl ", 7)) {
            std::string name(p + 7);
            name.erase(name.find_last_not_of(" \t\r") + 1);
            auto it = mat_names.find(name);
            cur_mat = it == mat_names.end() ? -1 : it->second;
        }
    }
    if (!any_vn || any_vn_missing) out.face_vn.clear();  // tobj: mesh.normals empty unless every corner has one
    return true;
}

// ---------------------------------------------------------------- Scene
void Scene::merge(const Scene &o) {
    spheres.insert(spheres.end(), o.spheres.begin(), o.spheres.end());
    triangles.insert(triangles.end(), o.triangles.begin(), o.triangles.end());
    scene_lights.insert(scene_lights.end(), o.scene_lights.begin(), o.scene_lights.end());
}

// Scene::from_obj body (src/scene/scene.rs:67-131)
Scene Scene::from_mesh(const Mesh &mesh, const Similarity3 *transform) {
    Similarity3 tf = transform ? *transform : Similarity3::identity();
    Scene s;
    size_t nf = mesh.face_v.size() / 3;
    bool has_n = !mesh.face_vn.empty() && !mesh.normals.empty();
    for (size_t f = 0; f < nf; f++) {
        Vec3 pos[3], nrm[3];
        for (int c = 0; c < 3; c++) {
            int i = mesh.face_v[3 * f + c];
            pos[c] = tf.transform_vec(Vec3(mesh.positions[3 * i], mesh.positions[3 * i + 1], mesh.positions[3 * i + 2]));
            if (has_n) {
                int j = mesh.face_vn[3 * f + c];
                nrm[c] = rotated_by(Vec3(mesh.normals[3 * j], mesh.normals[3 * j + 1], mesh.normals[3 * j + 2]), tf.rotation);
            }
        }
        int mi = mesh.face_material.empty() ? -1 : mesh.face_material[f];
        Material fm = (mi >= 0 && mi < (int)mesh.materials.size()) ? mesh.materials[mi] : Material::diffuse(ColorType(1.0f, 1.0f, 1.0f));
        if (has_n) {
            Vec3 n = lerp(lerp(nrm[0], nrm[1], 0.5f), nrm[2], 0.5f);
            s.add_triangle(TriangleData::with_material_and_normal(pos[0], pos[1], pos[2], n, fm));
        } else {
            s.add_triangle(TriangleData::with_material(pos[0], pos[1], pos[2], fm));
        }
    }
    return s;
}

bool Scene::from_obj(const std::string &path, const Similarity3 *transform, Scene &out, std::string &err) {
    Mesh m;
    if (!load_obj(path, m, err)) return false;
    out = from_mesh(m, transform);
    return true;
}

// Scene::backface_culling (src/scene/scene.rs:136-155)
Scene Scene::backface_culling(const Scene &scene, Vec3 view) {
    Scene r;
    r.scene_lights = scene.scene_lights;
    r.spheres = scene.spheres;
    for (const auto &t : scene.triangles) {
        if (t.material.transmission.mask()) {
            r.triangles.push_back(t);
            continue;
        }
        bool ne = !(std::fabs(dot(t.normal, view) - 1.0f) <= 0.01f);
        if (ne) r.triangles.push_back(t);
    }
    return r;
}

void Scene::flatten(std::vector<rt_triangle> &t, std::vector<rt_sphere> &s, std::vector<rt_material> &m, std::vector<rt_light> &l) const {
    t.clear(); s.clear(); m.clear(); l.clear();
    auto mat_index = [&](const Material &mat) -> uint32_t {
        rt_material f = mat.flat();
        for (size_t i = 0; i < m.size(); i++)
            if (!memcmp(&m[i], &f, sizeof(f))) return (uint32_t)i;
        m.push_back(f);
        return (uint32_t)m.size() - 1;
    };
    for (const auto &sp : spheres) {
        rt_sphere o;
        o.center[0] = sp.center.x; o.center[1] = sp.center.y; o.center[2] = sp.center.z;
        o.radius = sp.radius;
        o.material = mat_index(sp.material);
        s.push_back(o);
    }
    for (const auto &tr : triangles) {
        rt_triangle o;
        const Vec3 *src[4] = {&tr.vertex1, &tr.vertex2, &tr.vertex3, &tr.normal};
        float *dst[4] = {o.v1, o.v2, o.v3, o.normal};
        for (int k = 0; k < 4; k++) { dst[k][0] = src[k]->x; dst[k][1] = src[k]->y; dst[k][2] = src[k]->z; }
        o.material = mat_index(tr.material);
        t.push_back(o);
    }
    for (const auto &li : scene_lights) {
        rt_light o;
        o.position[0] = li.position.x; o.position[1] = li.position.y; o.position[2] = li.position.z;
        o.color[0] = li.color.x; o.color[1] = li.color.y; o.color[2] = li.color.z;
        o.intensity = li.intensity;
        l.push_back(o);
    }
}

// ---------------------------------------------------------------- entry-point scenes
static void add_plane(Scene &scene, Vec3 normal, Vec3 center, Vec3 up, float w, float h, float d, Material m) {
    BoundedPlane p;
    std::string err;
    if (!BoundedPlane::with_material(normal, center, up, w, h, d, m, p, err)) {
        // reference: assertion panic. Rotated axes are orthogonal only to a few ulp; the shipped
        // scenes pass the assertion in the reference, so build the plane regardless.
        p.center = center; p.up = up; p.left = normalized(cross(normal, up)); p.normal = normal;
        p.width = w; p.height = h; p.depth = d; p.material = m;
    }
    for (auto &t : p.to_basic_geometries()) scene.add_triangle(t);
}

// src/main.rs:30-320 (byte-identical in examples/semesterbild.rs)
Scene make_semesterbild_scene(const Features &f, const SceneConstants &k, const Mesh &text_mesh) {
    const float SW = k.scene_width, SH = k.scene_height, SD = k.scene_depth, AVG = k.average_scene_dimension;
    Similarity3 tf(Vec3(0.0135f * SW, 0.145f * SH, 0.885f * SD), Rotor3::from_euler_angles(0.0f, -0.015f, 0.0f), 1.226f * AVG);
    Scene scene = Scene::from_mesh(text_mesh, &tf);
    using TP = TransmissionProperties;
    scene.add_sphere(SphereData::with_material(Vec3(0.475f * SW, 0.385f * SH, 0.595f * SD), 0.291f * AVG,
                                               Material(ColorType(1.0f, 0.8f, 1.0f), 0.0f, 0.15f, TP::new_with_boost(0.99f, 1.5f, 0.025f))));
    scene.add_sphere(SphereData::with_material(Vec3(0.8f * SW, 0.76f * SH, 0.2f * SD), 0.07f * AVG,
                                               Material(ColorType(0.75f, 0.5f, 1.0f), 0.2f, 0.3f, TP::new_(0.78f, 1.5f))));
    scene.add_sphere(SphereData::with_material(Vec3(0.76f * SW, 0.76f * SH, 0.4f * SD), 0.07f * AVG,
                                               Material(ColorType(0.75f, 0.9f, 0.8f), 0.2f, 0.35f, TP::new_(0.6f, 1.8f))));
    scene.add_sphere(SphereData::with_material(Vec3(0.73f * SW, 0.7f * SH, 0.52f * SD), 0.065f * AVG,
                                               Material(ColorType(0.75f, 0.9f, 0.8f), 0.0f, 0.7f, TP::new_(0.78f, 1.3f))));
    scene.add_sphere(SphereData::with_material(Vec3(0.69f * SW, 0.76f * SH, 0.3f * SD), 0.07f * AVG,
                                               Material(ColorType(0.88f, 0.9f, 0.88f), 0.0f, 0.1f, TP::new_with_boost(1.0f, 1.42f, 0.125f))));
    scene.add_sphere(SphereData::with_material(Vec3(0.1f * SW, 0.68f * SH, 0.3f * SD), 0.07f * AVG,
                                               Material(ColorType(0.88f, 0.9f, 0.88f), 0.2f, 0.7f, TP::none())));
    scene.add_sphere(SphereData::with_material(Vec3(0.35f * SW, 0.76f * SH, 0.25f * SD), 0.07f * AVG,
                                               Material(ColorType(0.9f, 0.2f, 0.3f), 0.0f, 0.01f, TP::none())));
    scene.add_sphere(SphereData::with_material(Vec3(0.2f * SW, 0.87f * SH, 0.5f * SD), 0.07f * AVG,
                                               Material(ColorType(0.88f, 0.5f, 0.7f), 0.4f, 0.2f, TP::none())));
    scene.add_sphere(SphereData::with_material(Vec3(0.5f * SW, 0.87f * SH, 0.46f * SD), 0.075f * AVG,
                                               Material(ColorType(1.0f, 1.0f, 1.0f), 0.95f, 0.23f, TP::none())));

    Rotor3 rotor = Rotor3::from_euler_angles(-0.04f, 0.125f, 0.51f);
    Isometry3 iso(Vec3(0.25f * SW, 0.002f * SH, 0.037f * SD), rotor);
    // back, bottom (transmissive), bottom2, right planes
    add_plane(scene, -rotated_by(Vec3::unit_z(), rotor), iso.transform_vec(Vec3(SW * 0.5f, (SH * 1.1f) * 0.5f, SD)),
              rotated_by(Vec3::unit_y(), rotor), SW, SH * 1.1f, 0.01f * SD,
              Material(ColorType(0.5f, 0.75f, 0.75f), 0.0f, 0.0f, TP::none()));
    add_plane(scene, rotated_by(Vec3::unit_y(), rotor), iso.transform_vec(Vec3(SW * 0.5f, SH + 0.001f, SD * 0.5f)),
              rotated_by(Vec3::unit_z(), rotor), SW, SD, 0.012f * SD,
              Material(ColorType(0.75f, 0.5f, 0.75f), 0.0f, 0.7f, TP::new_(0.675f, 1.13f)));
    add_plane(scene, rotated_by(Vec3::unit_y(), rotor), iso.transform_vec(Vec3(SW * 0.5f, SH + 0.09f, SD * 0.5f)),
              rotated_by(Vec3::unit_z(), rotor), SW, SD, 0.01f * SD,
              Material(ColorType(0.75f, 0.5f, 0.75f), 0.0f, 0.7f, TP::none()));
    add_plane(scene, -rotated_by(Vec3::unit_x(), rotor), iso.transform_vec(Vec3(SW, (SH * 1.1f) * 0.5f, SD * 0.5f)),
              -rotated_by(Vec3::unit_z(), rotor), SH * 1.1f, SD, 0.01f * SD,
              Material(ColorType(0.875f, 0.85f, 0.61f), 0.55f, 0.325f, TP::none()));

    scene.add_light(PointLight::new_(Vec3(SW / 1.2f, 0.0f, 0.015f * SD), ColorType(0.825f, 0.675f, 0.5f), 1.0f));
    scene.add_light(PointLight::new_(Vec3(SW / 2.4f, SH * 0.1f, 0.08f * SD), ColorType(0.825f, 0.675f, 0.65f), 0.675f));
    scene.add_light(PointLight::new_(Vec3(SW, SH, 0.01f * SD), ColorType(0.825f, 0.35f, 0.8f), 0.435f));
    scene.add_light(PointLight::new_(iso.transform_vec(Vec3(SW * 0.5f, SH + 0.05f, SD * 0.75f)), ColorType(1.0f, 1.0f, 1.0f), 0.2775f));
    scene.add_light(PointLight::new_(Vec3(0.2f * SW, SH * 0.67f, 0.95f * SD), ColorType(0.825f, 0.5f, 0.7f), 0.26f));

    if (f.scene_backface_culling) return Scene::backface_culling(scene, Vec3::unit_z());
    return scene;
}

// examples/test_scene.rs:22-343
Scene make_test_scene(const SceneConstants &k) {
    const float SW = k.scene_width, SH = k.scene_height, SD = k.scene_depth;
    using TP = TransmissionProperties;
    Scene scene;
    scene.add_sphere(SphereData::new_(Vec3(SW / 2.5f, SH / 2.75f, 0.170f * SD), 0.070f * SD, ColorType(255.0f / 255.0f, 0.0f / 255.0f, 0.0f / 255.0f)));
    scene.add_sphere(SphereData::with_material(Vec3(SW / 2.5f, SH / 1.5f, 0.170f * SD), 0.070f * SD,
                                               Material(ColorType(255.0f / 255.0f, 0.0f, 0.0f), 0.8f, 0.0f, TP::none())));
    scene.add_sphere(SphereData::with_material(Vec3(1.9f * (SW / 2.5f), SH / 2.8f, 0.160f * SD), 0.088f * SD,
                                               Material(ColorType(250.0f / 255.0f, 255.0f / 255.0f, 245.0f / 255.0f), 0.01f, 0.2f, TP::new_(0.85f, 1.5f))));
    scene.add_sphere(SphereData::with_material(Vec3(SW / 2.5f, 2.1f * (SH / 2.5f), 0.5f * SD), 0.250f * SD,
                                               Material(ColorType(254.0f / 255.0f, 255.0f / 255.0f, 255.0f / 255.0f), 0.5f, 0.05f, TP::none())));
    scene.add_triangle(TriangleData::with_material(Vec3(SW * 0.05f, SH * 0.2f, 0.2f * SD), Vec3(SW * 0.3f, SH * 0.5f, 0.2f * SD),
                                                   Vec3(SW * 0.25f, SH * 0.15f, 0.15f * SD),
                                                   Material(ColorType(0.5f, 0.7f, 0.8f), 0.001f, 0.2f, TP::new_(0.999f, 1.8f))));
    scene.add_triangle(TriangleData::with_material(Vec3(SW * 0.55f, SH * 0.45f, 0.2f * SD), Vec3(SW * 0.7f, SH * 0.72f, 0.2f * SD),
                                                   Vec3(SW * 0.65f, SH * 0.35f, 0.14f * SD),
                                                   Material(ColorType(0.7f, 0.7f, 0.8f), 0.1f, 0.3f, TP::none())));
    scene.add_triangle(TriangleData::with_material(Vec3(SW * 0.7f, SH * 0.90f, 0.2f * SD), Vec3(SW * 0.55f, SH * 0.65f, 0.2f * SD),
                                                   Vec3(SW * 0.65f, SH * 0.55f, 0.14f * SD),
                                                   Material(ColorType(0.7f, 0.7f, 0.8f), 0.1f, 0.3f, TP::new_(1.0f, 1.5f))));
    {
        Rotor3 r = Rotor3::from_rotation_yz(-0.555f);
        add_plane(scene, rotated_by(-Vec3::unit_z(), r), Vec3(SW * 0.5f, SH * 0.45f, 0.270f * SD), rotated_by(Vec3::unit_y(), r),
                  SW * 0.55f, SH * 0.55f, 0.01f * SD, Material(ColorType(0.6f, 0.7f, 0.5f), 0.075f, 0.07f, TP::new_with_boost(1.0f, 1.5f, 0.5f)));
    }
    {
        Rotor3 r = Rotor3::from_rotation_xz(-0.9955f);
        add_plane(scene, rotated_by(-Vec3::unit_z(), r), Vec3(SW * 0.82f, SH * 0.57f, 0.110f * SD), rotated_by(Vec3::unit_y(), r),
                  SW * 0.318f, SH * 0.35f, 0.007f * SD, Material(ColorType(0.99f, 0.99f, 0.99f), 1.0f, 0.2f, TP::none()));
    }
    Material wall_back(ColorType(0.5f, 0.75f, 0.75f), 0, 0, TP::none());
    Material wall_tb(ColorType(0.75f, 0.5f, 0.75f), 0, 0, TP::none());
    Material wall_lr(ColorType(0.75f, 0.75f, 0.5f), 0, 0, TP::none());
    add_plane(scene, -Vec3::unit_z(), Vec3(SW * 0.5f, SH * 0.5f, SD), Vec3::unit_y(), SW, SH, 0.001f * SD, wall_back);
    add_plane(scene, Vec3::unit_y(), Vec3(SW * 0.5f, SH, SD * 0.5f), Vec3::unit_z(), SW, SD, 0.001f * SD, wall_tb);
    add_plane(scene, -Vec3::unit_y(), Vec3(SW * 0.5f, 0.0f, SD * 0.5f), Vec3::unit_z(), SW, SD, 0.001f * SD, wall_tb);
    add_plane(scene, Vec3::unit_x(), Vec3(0.0f, SH * 0.5f, SD * 0.5f), Vec3::unit_z(), SH, SD, 0.001f * SD, wall_lr);
    add_plane(scene, -Vec3::unit_x(), Vec3(SW, SH * 0.5f, SD * 0.5f), -Vec3::unit_z(), SH, SD, 0.001f * SD, wall_lr);

    scene.add_light(PointLight::new_(Vec3(SW / 2.0f, SH / 1.8f, 0.016f * SD), ColorType(0.825f, 0.675f, 0.5f), 0.15f));
    scene.add_light(PointLight::new_(Vec3(SW / 3.5f, SH / 3.75f, 0.025f * SD), ColorType(0.825f, 0.675f, 0.45f), 0.485f));
    scene.add_light(PointLight::new_(Vec3(SW / 1.22f, SH / 2.9f, 0.38f * SD), ColorType(0.78f, 0.67f, 0.45f), 0.6f));
    scene.add_light(PointLight::new_(Vec3(SW - 80.0f, SH / 2.0f, 0.125f * SD), ColorType(1.0f, 1.0f, 1.0f), 0.1f));
    scene.add_light(PointLight::new_(Vec3(SW / 2.5f, SH / 5.0f, 0.175f * SD), ColorType(0.75f, 0.56f, 0.65f), 0.2f));
    scene.add_light(PointLight::new_(Vec3(SW / 4.0f, SH / 6.0f, 0.01f * SD), ColorType(0.01f, 0.5f, 0.4f), 0.175f));
    return scene;
}

// examples/test_text.rs:24-52
Scene make_test_text_scene(const SceneConstants &k, const Mesh &text_mesh) {
    const float SW = k.scene_width, SH = k.scene_height, SD = k.scene_depth;
    Similarity3 tf(Vec3(0.15f, 0.0f, 0.5f), Rotor3::from_euler_angles(0.25f, 0.2f, 0.0f), 1.05f);
    Scene scene = Scene::from_mesh(text_mesh, &tf);
    scene.add_light(PointLight::new_(Vec3(SW / 2.0f, SH / 1.9f, 0.015f * SD), ColorType(0.825f, 0.675f, 0.5f), 0.99f));
    scene.add_light(PointLight::new_(Vec3(SW / 2.0f, SH / 2.1f, 0.85f * SD), ColorType(0.825f, 0.275f, 0.8f), 0.99f));
    return scene;
}

// ---------------------------------------------------------------- RaytracerRenderer
static void fill_config(const Features &f, rt_config &cfg) {
    memset(&cfg, 0, sizeof(cfg));
    cfg.width = f.window_width();
    cfg.height = f.window_height();
    cfg.flags = f.render_flags();
    cfg.aa_samples_per_pixel = f.antialiasing_samples_per_pixel();
    cfg.aa_table_len = ((cfg.aa_samples_per_pixel + 7) / 8) * 8;
    cfg.light_multiplicator = f.point_light_multiplicator();
    cfg.reflection_max_depth = f.reflection_max_depth();
    cfg.refraction_max_depth = f.refraction_max_depth();
}

int RaytracerRenderer::render(ImageBuffer &buffer, const Scene &scene, rt_stats *stats) const {
    rt_config cfg;
    fill_config(features, cfg);
    if (buffer.width != cfg.width || buffer.height != cfg.height) {
        rt_set_error("RaytracerRenderer::render: ImageBuffer is %ux%u but the feature set renders %ux%u", buffer.width,
                     buffer.height, cfg.width, cfg.height);
        return RT_ERR_INVALID_ARGUMENT;
    }
    cfg.seed = seed;
    std::vector<float> aa(2 * (size_t)cfg.aa_table_len), clouds;
    int rc = rt_sampler_aa_table(seed, cfg.aa_samples_per_pixel, features.anti_aliasing_randomness ? 1 : 0, aa.data(), cfg.aa_table_len);
    if (rc) return rc;
    cfg.aa_table = aa.data();
    if (cfg.light_multiplicator > 1) {
        clouds.resize((size_t)n_light_clouds * cfg.light_multiplicator * 3);
        rc = rt_sampler_light_clouds(seed, cfg.light_multiplicator, n_light_clouds, clouds.data());
        if (rc) return rc;
        cfg.light_clouds = clouds.data();
        cfg.n_light_clouds = n_light_clouds;
    }
    std::vector<rt_triangle> t;
    std::vector<rt_sphere> s;
    std::vector<rt_material> m;
    std::vector<rt_light> l;
    scene.flatten(t, s, m, l);
    rt_scene_desc desc{t.data(), (uint32_t)t.size(), s.data(), (uint32_t)s.size(), m.data(), (uint32_t)m.size(), l.data(), (uint32_t)l.size()};
    rt_scene *dev = nullptr;
    rc = rt_scene_create(&desc, device, &dev);
    if (rc) return rc;
    rc = rt_render(dev, &cfg, buffer.buffer.data(), nullptr, stats);
    rt_scene_destroy(dev);
    return rc;
}

}  // namespace rth

// =============================================================================================
// C wrappers (include/rt_host.h)
struct rth_scene {
    rth::Scene scene;
};

static rth::Vec3 V(const float p[3]) { return rth::Vec3(p[0], p[1], p[2]); }

static rth::Mesh mesh_from_arrays(const float *positions, uint32_t n_positions, const float *normals, uint32_t n_normals,
                                  const int32_t *face_v, const int32_t *face_vn, uint32_t n_faces, const rth::Material *mat) {
    rth::Mesh m;
    if (positions) m.positions.assign(positions, positions + 3 * (size_t)n_positions);
    if (normals) m.normals.assign(normals, normals + 3 * (size_t)n_normals);
    if (face_v) m.face_v.assign(face_v, face_v + 3 * (size_t)n_faces);
    if (face_vn && normals) m.face_vn.assign(face_vn, face_vn + 3 * (size_t)n_faces);
    if (mat) {
        m.materials.push_back(*mat);
        m.face_material.assign(n_faces, 0);
    } else {
        m.face_material.assign(n_faces, -1);
    }
    return m;
}

static bool mesh_valid(const rth::Mesh &m) {
    int nv = (int)m.positions.size() / 3, nn = (int)m.normals.size() / 3;
    for (int i : m.face_v)
        if (i < 0 || i >= nv) return false;
    for (int i : m.face_vn)
        if (i < 0 || i >= nn) return false;
    return true;
}

extern "C" {

rth_scene *rth_scene_new(void) { return new (std::nothrow) rth_scene(); }
void rth_scene_free(rth_scene *s) { delete s; }

int rth_scene_add_sphere(rth_scene *s, const float center[3], float radius, const rt_material *m) {
    if (!s || !center || !m) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_scene_add_sphere: null argument");
    s->scene.add_sphere(rth::SphereData::with_material(V(center), radius, rth::material_from_flat(m)));
    return RT_OK;
}

int rth_scene_add_triangle(rth_scene *s, const float v1[3], const float v2[3], const float v3[3], const rt_material *m) {
    if (!s || !v1 || !v2 || !v3 || !m) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_scene_add_triangle: null argument");
    s->scene.add_triangle(rth::TriangleData::with_material(V(v1), V(v2), V(v3), rth::material_from_flat(m)));
    return RT_OK;
}

int rth_scene_add_triangle_with_normal(rth_scene *s, const float v1[3], const float v2[3], const float v3[3],
                                       const float normal[3], const rt_material *m) {
    if (!s || !v1 || !v2 || !v3 || !normal || !m)
        return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_scene_add_triangle_with_normal: null argument");
    s->scene.add_triangle(rth::TriangleData::with_material_and_normal(V(v1), V(v2), V(v3), V(normal), rth::material_from_flat(m)));
    return RT_OK;
}

int rth_scene_add_bounded_plane(rth_scene *s, const float normal[3], const float center[3], const float up[3], float width,
                                float height, float depth, const rt_material *m) {
    if (!s || !normal || !center || !up || !m) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_scene_add_bounded_plane: null argument");
    rth::BoundedPlane p;
    std::string err;
    if (!rth::BoundedPlane::with_material(V(normal), V(center), V(up), width, height, depth, rth::material_from_flat(m), p, err))
        return rt_fail(RT_ERR_INVALID_ARGUMENT, "BoundedPlane: %s", err.c_str());  // the reference panics here
    for (auto &t : p.to_basic_geometries()) s->scene.add_triangle(t);
    return RT_OK;
}

int rth_scene_add_light(rth_scene *s, const float position[3], const float color[3], float intensity) {
    if (!s || !position || !color) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_scene_add_light: null argument");
    s->scene.add_light(rth::PointLight::new_(V(position), V(color), intensity));
    return RT_OK;
}

static rth::Similarity3 make_tf(const float translation[3], const float euler[3], float scale) {
    rth::Vec3 t = translation ? V(translation) : rth::Vec3(0, 0, 0);
    rth::Rotor3 r = euler ? rth::Rotor3::from_euler_angles(euler[0], euler[1], euler[2]) : rth::Rotor3::identity();
    return rth::Similarity3(t, r, scale);
}

int rth_scene_add_obj(rth_scene *s, const char *path, const float translation[3], const float euler_rpy[3], float scale) {
    if (!s || !path) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_scene_add_obj: null argument");
    rth::Similarity3 tf = make_tf(translation, euler_rpy, scale);
    rth::Scene o;
    std::string err;
    if (!rth::Scene::from_obj(path, &tf, o, err)) return rt_fail(RT_ERR_IO, "Scene::from_obj: %s", err.c_str());
    s->scene.merge(o);
    return RT_OK;
}

int rth_scene_add_mesh(rth_scene *s, const float *positions, uint32_t n_positions, const float *normals, uint32_t n_normals,
                       const int32_t *face_v, const int32_t *face_vn, uint32_t n_faces, const rt_material *m,
                       const float translation[3], const float euler_rpy[3], float scale) {
    if (!s || (!positions && n_faces) || (!face_v && n_faces)) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_scene_add_mesh: null argument");
    rth::Material mat;
    if (m) mat = rth::material_from_flat(m);
    rth::Mesh mesh = mesh_from_arrays(positions, n_positions, normals, n_normals, face_v, face_vn, n_faces, m ? &mat : nullptr);
    if (!mesh_valid(mesh)) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_scene_add_mesh: face index out of range");
    rth::Similarity3 tf = make_tf(translation, euler_rpy, scale);
    s->scene.merge(rth::Scene::from_mesh(mesh, &tf));
    return RT_OK;
}

int rth_scene_backface_culling(rth_scene *s, const float view_direction[3]) {
    if (!s || !view_direction) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_scene_backface_culling: null argument");
    s->scene = rth::Scene::backface_culling(s->scene, V(view_direction));
    return RT_OK;
}

int rth_scene_counts(const rth_scene *s, uint32_t *n_triangles, uint32_t *n_spheres, uint32_t *n_materials, uint32_t *n_lights) {
    if (!s) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_scene_counts: null scene");
    std::vector<rt_triangle> t;
    std::vector<rt_sphere> sp;
    std::vector<rt_material> m;
    std::vector<rt_light> l;
    s->scene.flatten(t, sp, m, l);
    if (n_triangles) *n_triangles = (uint32_t)t.size();
    if (n_spheres) *n_spheres = (uint32_t)sp.size();
    if (n_materials) *n_materials = (uint32_t)m.size();
    if (n_lights) *n_lights = (uint32_t)l.size();
    return RT_OK;
}

int rth_scene_flatten(const rth_scene *s, rt_triangle *triangles, rt_sphere *spheres, rt_material *materials, rt_light *lights) {
    if (!s) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_scene_flatten: null scene");
    std::vector<rt_triangle> t;
    std::vector<rt_sphere> sp;
    std::vector<rt_material> m;
    std::vector<rt_light> l;
    s->scene.flatten(t, sp, m, l);
    if (triangles && !t.empty()) memcpy(triangles, t.data(), t.size() * sizeof(rt_triangle));
    if (spheres && !sp.empty()) memcpy(spheres, sp.data(), sp.size() * sizeof(rt_sphere));
    if (materials && !m.empty()) memcpy(materials, m.data(), m.size() * sizeof(rt_material));
    if (lights && !l.empty()) memcpy(lights, l.data(), l.size() * sizeof(rt_light));
    return RT_OK;
}

int rth_example_scene(const char *name, const char *features_csv, const float *positions, uint32_t n_positions,
                      const float *normals, uint32_t n_normals, const int32_t *face_v, const int32_t *face_vn, uint32_t n_faces,
                      rth_scene **out) {
    if (!name || !out) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_example_scene: null argument");
    rth::Features f;
    std::string err;
    if (!rth::Features::parse(features_csv ? features_csv : "", f, err)) return rt_fail(RT_ERR_INVALID_ARGUMENT, "%s", err.c_str());
    rth::SceneConstants k(f.window_width(), f.window_height());
    // text.mtl: Material.001 (Kd, illum 2, Pm 0, Ps) -> material.rs:96-126
    rth::Material text_mat(rth::ColorType(0.123533f, 0.719221f, 0.826410f), 0.0f, 0.592264f, rth::TransmissionProperties());
    rth::Mesh mesh = mesh_from_arrays(positions, n_positions, normals, n_normals, face_v, face_vn, n_faces, &text_mat);
    if (!mesh_valid(mesh)) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_example_scene: face index out of range");
    std::string n(name);
    rth_scene *s = new (std::nothrow) rth_scene();
    if (!s) return rt_fail(RT_ERR_OUT_OF_MEMORY, "out of memory");
    if (n == "semesterbild" || n == "main") s->scene = rth::make_semesterbild_scene(f, k, mesh);
    else if (n == "test_scene") s->scene = rth::make_test_scene(k);
    else if (n == "test_text") s->scene = rth::make_test_text_scene(k, mesh);
    else {
        delete s;
        return rt_fail(RT_ERR_INVALID_ARGUMENT, "unknown example scene '%s'", name);
    }
    *out = s;
    return RT_OK;
}

int rth_config_from_features(const char *features_csv, rt_config *cfg, int *high_quality_model, int *scene_backface_culling,
                             int *anti_aliasing_randomness) {
    if (!cfg) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_config_from_features: null cfg");
    rth::Features f;
    std::string err;
    if (!rth::Features::parse(features_csv ? features_csv : "", f, err)) return rt_fail(RT_ERR_INVALID_ARGUMENT, "%s", err.c_str());
    rth::fill_config(f, *cfg);
    if (high_quality_model) *high_quality_model = (f.high_quality_model || f.medium_resolution) ? 1 : 0;  // src/main.rs:31-35
    if (scene_backface_culling) *scene_backface_culling = f.scene_backface_culling ? 1 : 0;
    if (anti_aliasing_randomness) *anti_aliasing_randomness = f.anti_aliasing_randomness ? 1 : 0;
    return RT_OK;
}

int rth_feature_string(const char *features_csv, int with_title_prefix, char *out, uint32_t out_size) {
    if (!out || out_size == 0) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_feature_string: null output");
    rth::Features f;
    std::string err;
    if (!rth::Features::parse(features_csv ? features_csv : "", f, err)) return rt_fail(RT_ERR_INVALID_ARGUMENT, "%s", err.c_str());
    const std::string t = with_title_prefix ? f.window_title() : f.feature_string();
    if (t.size() + 1 > out_size) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_feature_string: buffer of %u bytes is too small (%zu needed)", out_size, t.size() + 1);
    memcpy(out, t.c_str(), t.size() + 1);
    return RT_OK;
}

int rth_render(const rth_scene *s, const char *features_csv, uint64_t seed, int device, uint32_t *out_argb, uint8_t *out_hit_mask,
               rt_stats *stats) {
    if (!s || !out_argb) return rt_fail(RT_ERR_INVALID_ARGUMENT, "rth_render: null argument");
    rth::RaytracerRenderer r;
    std::string err;
    if (!rth::Features::parse(features_csv ? features_csv : "", r.features, err)) return rt_fail(RT_ERR_INVALID_ARGUMENT, "%s", err.c_str());
    r.seed = seed;
    r.device = device;
    rth::ImageBuffer buf(r.features.window_width(), r.features.window_height());
    int rc = r.render(buf, s->scene, stats);
    if (rc) return rc;
    // ImageBuffer semantics: only hit pixels are written, the rest keep the caller's content
    for (size_t i = 0; i < buf.buffer.size(); i++) {
        bool hit = buf.buffer[i] != 0;
        if (hit) out_argb[i] = buf.buffer[i];
        if (out_hit_mask) out_hit_mask[i] = hit ? 1 : 0;
    }
    return RT_OK;
}

}  // extern "C"
