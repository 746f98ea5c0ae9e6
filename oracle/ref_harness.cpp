// ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A C-API shim around the UNMODIFIED reference objects (src/scene.cpp,
// src/geometry.cpp, src/parsers.cpp, src/options.cpp, src/exceptions.cpp),
// compiled from where they lie under /root/reference by oracle/Makefile into
// oracle/_ref/libref.so.  No reference source is copied: this file is our own
// code and only *calls* the reference's public members
//   Scene::traceRay / Scene::castRay           (src/scene.h:15-17)
//   Camera::calculateViewingRay                (src/rtbase.h:74-84)
//   RTIParser::parseFile                       (src/parsers.h:36)
// The reference's own Scene::renderScene aborts when W*H % 2000 != 0
// (src/scene.cpp:21-25, block not clamped) and races on lazily cached
// light/camera vectors (src/lights.h:28-33,56-61; src/rtbase.h:86-95), so the
// pixel loop is restated here with the clamp and a cache pre-warm; the
// per-pixel work (traceRay and everything below it) is the reference's code.
// ref_render_stock() calls the reference's renderScene itself where it is safe
// (W*H % 2000 == 0) so tests can check the two loops agree bit for bit.
//
// `#define private public` is applied to the reference headers in THIS TU only
// (class layout is unaffected) to read Scene::geometries_/lights_ for the
// flat-scene export and the cache pre-warm.

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#define private public
#include "scene.h"
#include "parsers.h"
#include "options.h"
#undef private

#include "rt_b200.h"

// ---- optional ray counting (libref_count.so only; linked with
// -Wl,--wrap=_ZN8Geometry27calculateIntersectionNormalE3RayRN5Eigen6MatrixIdLi4ELi1ELi0ELi4ELi1EEES4_b) ----
static std::atomic<uint64_t> g_geomtests(0);
static thread_local uint64_t t_geomtests = 0;
#ifdef REF_COUNT_RAYS
extern "C" bool __real__ZN8Geometry27calculateIntersectionNormalE3RayRN5Eigen6MatrixIdLi4ELi1ELi0ELi4ELi1EEES4_b(
    Geometry*, Ray*, Vector4d&, Vector4d&, bool);
extern "C" bool __wrap__ZN8Geometry27calculateIntersectionNormalE3RayRN5Eigen6MatrixIdLi4ELi1ELi0ELi4ELi1EEES4_b(
    Geometry* self, Ray* r, Vector4d& p, Vector4d& n, bool rev) {
    t_geomtests++;
    return __real__ZN8Geometry27calculateIntersectionNormalE3RayRN5Eigen6MatrixIdLi4ELi1ELi0ELi4ELi1EEES4_b(self, r, p, n, rev);
}
#endif

struct RefScene {
    Scene scene;
    // flat export storage
    std::vector<rt_geometry> geoms;
    std::vector<rt_material> mats;
    std::vector<rt_light> lights;
    std::vector<double> fpts, fnrm;
    rt_scene flat;
    bool flattened = false;
};

static void set_err(char* err, int errlen, const std::string& s) {
    if (err && errlen > 0) {
        std::strncpy(err, s.c_str(), errlen - 1);
        err[errlen - 1] = 0;
    }
}

static void prewarm(Scene& scene) {
    Vector4d p(0.25, 0.5, 0.75, 1.0);
    for (auto& l : scene.lights_) {
        (void)l->calculateDistanceToLight(p);
        (void)l->calculateDirectionToLight(p);
    }
    (void)scene.camera_.calculateViewingRay(0.5, 0.5);
}

extern "C" {

void* ref_scene_load(const char** files, int nfiles, char* err, int errlen) {
    RefScene* rs = new RefScene();
    try {
        for (int i = 0; i < nfiles; i++) {
            RTIParser parser(rs->scene);
            parser.parseFile(files[i]);
        }
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        delete rs;
        return nullptr;
    }
    if (!rs->scene.hasCamera()) {
        set_err(err, errlen, "At least one camera must be specified.");
        delete rs;
        return nullptr;
    }
    prewarm(rs->scene);
    return rs;
}

// Build the REFERENCE object graph from a flat descriptor (used for scenes that exist only
// in memory, i.e. the synthetic benchmark scene): reference Sphere/Mesh/Light/Camera objects
// are created through the reference's own classes; transforms go through
// Transformable::forwardTransform so the reference computes its own inverse/determinant.
// Lights and camera are handed over already transformed (identity transform on the objects).
void* ref_scene_from_flat(const rt_scene* f, char* err, int errlen) {
    RefScene* rs = new RefScene();
    try {
        auto xf_of = [](const double* m) {
            Transform4d T = Transform4d::Identity();
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 4; c++) T.matrix()(r, c) = m[r * 4 + c];
            return T;
        };
        auto P = [](const double* v) { return Vector4d(v[0], v[1], v[2], 1.0); };
        auto C3 = [](const double* v) { return Color3d(v[0], v[1], v[2]); };
        Camera cam;
        cam.forwardTransform(Transform4d::Identity());
        cam.eyePoint(P(f->camera.eye));
        cam.lowerLeftPoint(P(f->camera.ll));
        cam.lowerRightPoint(P(f->camera.lr));
        cam.upperLeftPoint(P(f->camera.ul));
        cam.upperRightPoint(P(f->camera.ur));
        rs->scene.camera(cam);
        for (int i = 0; i < f->num_lights; i++) {
            const rt_light& l = f->lights[i];
            Light* out;
            if (l.type == RT_LIGHT_POINT) {
                PointLight* pl = new PointLight();
                pl->point(P(l.v));
                pl->falloffExponent_ = l.falloff;
                out = pl;
            } else if (l.type == RT_LIGHT_DIRECTIONAL) {
                DirectionalLight* dl = new DirectionalLight();
                dl->direction(Vector4d(l.v[0], l.v[1], l.v[2], 0.0));
                out = dl;
            } else {
                out = new AmbientLight();
            }
            out->forwardTransform(Transform4d::Identity());
            out->color_ = C3(l.color);
            rs->scene.addLight(std::unique_ptr<Light>(out));
        }
        for (int i = 0; i < f->num_geometries; i++) {
            const rt_geometry& g = f->geometries[i];
            const rt_material& m = f->materials[g.material];
            Geometry* out;
            if (g.type == RT_GEOM_SPHERE) {
                Sphere* s = new Sphere();
                s->center_ = P(g.center);
                s->radius_ = (float)g.radius;
                out = s;
            } else {
                Mesh* me = new Mesh();
                me->faces_.resize((size_t)g.num_faces);
                for (int64_t k = 0; k < g.num_faces; k++) {
                    const double* fp = f->face_points + 9 * (g.first_face + k);
                    const double* fn = f->face_normals + 9 * (g.first_face + k);
                    for (int v = 0; v < 3; v++) {
                        me->faces_[(size_t)k].points_[v] = Vector4d(fp[3 * v], fp[3 * v + 1], fp[3 * v + 2], 1.0);
                        me->faces_[(size_t)k].normals_[v] = Vector4d(fn[3 * v], fn[3 * v + 1], fn[3 * v + 2], 0.0);
                    }
                }
                if (g.type == RT_GEOM_MESH) me->updateBoundingBox();
                out = me;
            }
            out->forwardTransform(xf_of(g.fwd));
            out->material_.ambientColor_ = C3(m.ka);
            out->material_.diffuseColor_ = C3(m.kd);
            out->material_.specularColor_ = C3(m.ks);
            out->material_.reflectiveColor_ = C3(m.kr);
            out->material_.translucencyColor_ = C3(m.kt);
            out->material_.specularCoefficient_ = m.sp;
            out->material_.indexOfRefractivity_ = m.ior;
            rs->scene.addGeometry(std::unique_ptr<Geometry>(out));
        }
    } catch (const std::exception& e) {
        set_err(err, errlen, e.what());
        delete rs;
        return nullptr;
    }
    prewarm(rs->scene);
    return rs;
}

void ref_scene_free(void* h) { delete static_cast<RefScene*>(h); }

int ref_num_geometries(void* h) { return (int)static_cast<RefScene*>(h)->scene.geometries_.size(); }

// Render with the reference's traceRay; our loop (clamped blocks, same 2000-px
// dynamic self-scheduling as src/scene.cpp:13-24).  seconds = worker section only.
int ref_render(void* h, int width, int height, int depth, int intersection_only, int threads,
               double* rgb, int32_t* geom_ids, double* seconds, uint64_t* castray_calls) {
    RefScene* rs = static_cast<RefScene*>(h);
    Scene& scene = rs->scene;
    programOptions.bounceDepth_ = depth;
    programOptions.intersectionOnly_ = intersection_only != 0;
    programOptions.renderThreadsCount_ = threads;
    prewarm(scene);
    const long total = (long)width * height;
    std::atomic<long> next(0);
    g_geomtests = 0;
    // the reference hands out 2000-pixel blocks (src/scene.cpp:13); small sample frames get
    // smaller blocks so that every thread has work (scheduling only, no effect on pixels)
    long bs = total / ((long)threads * 4);
    const int blockSize = (int)std::max(1L, std::min(2000L, bs));
    auto worker = [&]() {
        t_geomtests = 0;
        Camera& cam = scene.camera_;
        while (true) {
            long start = next.fetch_add(blockSize);
            if (start >= total) break;
            long end = std::min(start + blockSize, total);
            for (long i = start; i < end; i++) {
                int r = (int)(i / width), c = (int)(i % width);
                double rowFrac = (r + 0.5) / height;
                double colFrac = (c + 0.5) / width;
                Ray ray = cam.calculateViewingRay(rowFrac, colFrac);
                Color3d col = scene.traceRay(ray, depth);
                rgb[3 * i + 0] = col[0];
                rgb[3 * i + 1] = col[1];
                rgb[3 * i + 2] = col[2];
                if (geom_ids) {
                    double d; Geometry* g = nullptr; Vector4d P, N;
                    uint64_t save = t_geomtests;
                    bool hit = scene.castRay(ray, &d, &g, &P, &N, false);
                    t_geomtests = save;
                    int gi = -1;
                    if (hit)
                        for (size_t k = 0; k < scene.geometries_.size(); k++)
                            if (scene.geometries_[k].get() == g) { gi = (int)k; break; }
                    geom_ids[i] = gi;
                }
            }
        }
        g_geomtests += t_geomtests;
    };
    auto t0 = std::chrono::steady_clock::now();
    {
        std::vector<std::thread> pool;
        for (int i = 0; i < threads; i++) pool.emplace_back(worker);
        for (auto& t : pool) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (castray_calls) {
        size_t ng = scene.geometries_.size();
        *castray_calls = ng ? (uint64_t)g_geomtests / ng : 0;
    }
    if (intersection_only) {   // restated post-pass of src/scene.cpp:50-58
        double mx = std::numeric_limits<double>::min();
        for (long i = 0; i < total * 3; i++) mx = std::max(mx, rgb[i]);
        for (long i = 0; i < total * 3; i++) rgb[i] /= mx;
    }
    return 0;
}

// Face index of every primary hit, obtained from the UNMODIFIED reference code: the reference reports only the
// hit Geometry (Scene::castRay) and, through Geometry::calculateIntersectionNormal, the hit point and normal, but no
// face number.  For a pixel whose primary ray hits mesh g, every face k of g is put alone into a temporary Mesh
// (same transform) and handed to the reference's own Mesh code; the first k whose one-face answer equals the full
// mesh's answer bit for bit (point and normal) is the face the reference picked (its loop keeps the FIRST face of
// least distance, src/geometry.cpp:108-110).  Our only own logic is a generous object-space bounding-box
// pre-filter on the hit point that skips faces far away from it.  Spheres report face 0, misses -1.
int ref_primary_faces(void* h, int width, int height, const int32_t* geom_ids, int32_t* face_ids, int threads) {
    RefScene* rs = static_cast<RefScene*>(h);
    Scene& scene = rs->scene;
    prewarm(scene);
    const long total = (long)width * height;
    // per-mesh face boxes (object space), padded
    std::vector<std::vector<std::array<double, 6>>> boxes(scene.geometries_.size());
    for (size_t g = 0; g < scene.geometries_.size(); g++) {
        Mesh* me = dynamic_cast<Mesh*>(scene.geometries_[g].get());
        if (!me) continue;
        boxes[g].resize(me->faces_.size());
        for (size_t k = 0; k < me->faces_.size(); k++) {
            std::array<double, 6>& b = boxes[g][k];
            for (int a = 0; a < 3; a++) { b[a] = 1e300; b[3 + a] = -1e300; }
            for (int v = 0; v < 3; v++)
                for (int a = 0; a < 3; a++) {
                    b[a] = std::min(b[a], me->faces_[k].points_[v][a]);
                    b[3 + a] = std::max(b[3 + a], me->faces_[k].points_[v][a]);
                }
            for (int a = 0; a < 3; a++) {
                double pad = 1e-6 * (std::fabs(b[a]) + std::fabs(b[3 + a]) + 1.0);
                b[a] -= pad; b[3 + a] += pad;
            }
        }
    }
    std::atomic<long> next(0);
    std::atomic<int> bad(0);
    auto worker = [&]() {
        Camera& cam = scene.camera_;
        Mesh one;
        one.faces_.resize(1);
        const Geometry* bound = nullptr;
        while (true) {
            long start = next.fetch_add(256);
            if (start >= total) break;
            long end = std::min(start + 256, total);
            for (long i = start; i < end; i++) {
                const int gi = geom_ids[i];
                face_ids[i] = gi < 0 ? -1 : 0;
                if (gi < 0) continue;
                Geometry* g = scene.geometries_[(size_t)gi].get();
                Mesh* me = dynamic_cast<Mesh*>(g);
                if (!me) continue;
                int r = (int)(i / width), c = (int)(i % width);
                Ray ray = cam.calculateViewingRay((r + 0.5) / height, (c + 0.5) / width);
                Vector4d P, N;
                if (!g->calculateIntersectionNormal(ray, P, N, false)) { bad++; face_ids[i] = -2; continue; }
                if (bound != g) { one.forwardTransform(g->forwardTransform()); bound = g; }
                Vector4d Po = g->inverseTransform() * P;
                int found = -2;
                const auto& bx = boxes[(size_t)gi];
                for (size_t k = 0; k < me->faces_.size(); k++) {
                    const std::array<double, 6>& b = bx[k];
                    if (Po[0] < b[0] || Po[0] > b[3] || Po[1] < b[1] || Po[1] > b[4] || Po[2] < b[2] || Po[2] > b[5]) continue;
                    one.faces_[0] = me->faces_[k];
                    Vector4d P1, N1;
                    if (!one.calculateIntersectionNormal(ray, P1, N1, false)) continue;
                    if (P1 == P && N1 == N) { found = (int)k; break; }
                }
                if (found < 0) bad++;
                face_ids[i] = found;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int i = 0; i < std::max(1, threads); i++) pool.emplace_back(worker);
    for (auto& t : pool) t.join();
    return bad.load();
}

// The reference's own Scene::renderScene (only safe when W*H % 2000 == 0).
int ref_render_stock(void* h, int width, int height, int depth, int intersection_only, int threads,
                     double* rgb) {
    RefScene* rs = static_cast<RefScene*>(h);
    if (((long)width * height) % 2000 != 0) return -1;
    programOptions.bounceDepth_ = depth;
    programOptions.intersectionOnly_ = intersection_only != 0;
    programOptions.renderThreadsCount_ = threads;
    prewarm(rs->scene);
    Scene::RasterImage image(height, width);
    rs->scene.renderScene(image, nullptr);
    for (int r = 0; r < height; r++)
        for (int c = 0; c < width; c++)
            for (int k = 0; k < 3; k++)
                rgb[((long)r * width + c) * 3 + k] = image(r, c)[k];
    return 0;
}

// Per-ray Scene::castRay.  dir is handed to the Ray ctor (normalises).
int ref_cast_rays(void* h, int64_t n, const double* org, const double* dir, const uint8_t* reverse,
                  int32_t* geom, double* dist, double* point, double* normal) {
    RefScene* rs = static_cast<RefScene*>(h);
    Scene& scene = rs->scene;
    for (int64_t i = 0; i < n; i++) {
        Vector4d o(org[3 * i], org[3 * i + 1], org[3 * i + 2], 1.0);
        Vector4d d(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2], 0.0);
        int gi = -1;
        double dd = 0; Vector4d P = Vector4d::Zero(), N = Vector4d::Zero();
        try {
            Ray ray(o, d);
            Geometry* g = nullptr;
            bool hit = scene.castRay(ray, &dd, &g, &P, &N, reverse ? reverse[i] != 0 : false);
            if (hit) {
                for (size_t k = 0; k < scene.geometries_.size(); k++)
                    if (scene.geometries_[k].get() == g) { gi = (int)k; break; }
            } else { dd = 0; P.setZero(); N.setZero(); }
        } catch (const MathException&) {
            gi = -2;
        }
        if (geom) geom[i] = gi;
        if (dist) dist[i] = dd;
        if (point) { point[3 * i] = P[0]; point[3 * i + 1] = P[1]; point[3 * i + 2] = P[2]; }
        if (normal) { normal[3 * i] = N[0]; normal[3 * i + 1] = N[1]; normal[3 * i + 2] = N[2]; }
    }
    return 0;
}

// Per-ray Scene::traceRay.
int ref_trace_rays(void* h, int64_t n, const double* org, const double* dir, int depth,
                   const uint8_t* from_inside, double* rgb) {
    RefScene* rs = static_cast<RefScene*>(h);
    programOptions.bounceDepth_ = depth;
    programOptions.intersectionOnly_ = false;
    for (int64_t i = 0; i < n; i++) {
        Vector4d o(org[3 * i], org[3 * i + 1], org[3 * i + 2], 1.0);
        Vector4d d(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2], 0.0);
        Color3d c = Color3d::Zero();
        try {
            c = rs->scene.traceRay(Ray(o, d), depth, from_inside ? from_inside[i] != 0 : false);
        } catch (const MathException&) {
            c = Color3d::Constant(std::numeric_limits<double>::quiet_NaN());
        }
        rgb[3 * i] = c[0]; rgb[3 * i + 1] = c[1]; rgb[3 * i + 2] = c[2];
    }
    return 0;
}

// Camera rays (origin, normalised direction) for pixel indices.
int ref_camera_rays(void* h, int width, int height, int64_t n, const int64_t* pix, double* org, double* dir) {
    RefScene* rs = static_cast<RefScene*>(h);
    for (int64_t k = 0; k < n; k++) {
        int64_t i = pix[k];
        int r = (int)(i / width), c = (int)(i % width);
        Ray ray = rs->scene.camera_.calculateViewingRay((r + 0.5) / height, (c + 0.5) / width);
        for (int j = 0; j < 3; j++) { org[3 * k + j] = ray.origin()[j]; dir[3 * k + j] = ray.direction()[j]; }
    }
    return 0;
}

// Export the reference object graph as the ABI's flat descriptor so tests can
// compare it bit-for-bit with the product host's own flattening.
const rt_scene* ref_scene_flatten(void* h) {
    RefScene* rs = static_cast<RefScene*>(h);
    if (rs->flattened) return &rs->flat;
    Scene& scene = rs->scene;
    prewarm(scene);
    rt_scene& f = rs->flat;
    std::memset(&f, 0, sizeof(f));
    Camera& cam = scene.camera_;
    for (int k = 0; k < 3; k++) {
        f.camera.eye[k] = cam.xfEyePoint_[k];
        f.camera.ll[k] = cam.xfLowerLeftPoint_[k];
        f.camera.lr[k] = cam.xfLowerRightPoint_[k];
        f.camera.ul[k] = cam.xfUpperLeftPoint_[k];
        f.camera.ur[k] = cam.xfUpperRightPoint_[k];
    }
    for (auto& gp : scene.geometries_) {
        Geometry* g = gp.get();
        rt_geometry rg; std::memset(&rg, 0, sizeof(rg));
        rt_material m; std::memset(&m, 0, sizeof(m));
        for (int k = 0; k < 3; k++) {
            m.ka[k] = g->material_.ambientColor_[k];
            m.kd[k] = g->material_.diffuseColor_[k];
            m.ks[k] = g->material_.specularColor_[k];
            m.kr[k] = g->material_.reflectiveColor_[k];
            m.kt[k] = g->material_.translucencyColor_[k];
        }
        m.sp = g->material_.specularCoefficient_;
        m.ior = g->material_.indexOfRefractivity_;
        rg.material = (int32_t)rs->mats.size();
        rs->mats.push_back(m);
        const Matrix4d& F = g->forwardTransform().matrix();
        const Matrix4d& I = g->inverseTransform().matrix();
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) { rg.fwd[r * 4 + c] = F(r, c); rg.inv[r * 4 + c] = I(r, c); }
        rg.det = g->transformDeterminant();
        if (Sphere* s = dynamic_cast<Sphere*>(g)) {
            rg.type = RT_GEOM_SPHERE;
            for (int k = 0; k < 3; k++) rg.center[k] = s->center_[k];
            rg.radius = (double)s->radius_;
            rg.radius2 = (double)(s->radius_ * s->radius_);
        } else if (Mesh* me = dynamic_cast<Mesh*>(g)) {
            bool has_bbox = (me->boundingBoxMin_ != me->boundingBoxMax_) && me->faces_.size() > 1;
            // `tri` meshes never call updateBoundingBox (src/parsers.cpp:213-221): bbox stays 0/0
            bool is_tri = me->faces_.size() == 2 && me->boundingBoxMin_ == me->boundingBoxMax_;
            rg.type = is_tri ? RT_GEOM_TRI : RT_GEOM_MESH;
            rg.first_face = (int64_t)(rs->fpts.size() / 9);
            rg.num_faces = (int64_t)me->faces_.size();
            rg.use_bbox = has_bbox ? 1 : 0;
            for (int k = 0; k < 3; k++) { rg.bbmin[k] = me->boundingBoxMin_[k]; rg.bbmax[k] = me->boundingBoxMax_[k]; }
            for (auto& face : me->faces_)
                for (int v = 0; v < 3; v++)
                    for (int k = 0; k < 3; k++) {
                        rs->fpts.push_back(face.points_[v][k]);
                        rs->fnrm.push_back(face.normals_[v][k]);
                    }
        }
        rs->geoms.push_back(rg);
    }
    Vector4d probe(0, 0, 0, 1);
    for (auto& lp : scene.lights_) {
        Light* l = lp.get();
        rt_light rl; std::memset(&rl, 0, sizeof(rl));
        for (int k = 0; k < 3; k++) rl.color[k] = l->color_[k];
        if (PointLight* pl = dynamic_cast<PointLight*>(l)) {
            rl.type = RT_LIGHT_POINT;
            for (int k = 0; k < 3; k++) rl.v[k] = pl->xfPoint_[k];
            rl.falloff = pl->falloffExponent_;
        } else if (DirectionalLight* dl = dynamic_cast<DirectionalLight*>(l)) {
            rl.type = RT_LIGHT_DIRECTIONAL;
            for (int k = 0; k < 3; k++) rl.v[k] = dl->xfDirection_[k];
        } else {
            rl.type = RT_LIGHT_AMBIENT;
        }
        rs->lights.push_back(rl);
    }
    f.num_geometries = (int32_t)rs->geoms.size();
    f.num_materials = (int32_t)rs->mats.size();
    f.num_lights = (int32_t)rs->lights.size();
    f.num_faces = (int64_t)(rs->fpts.size() / 9);
    f.geometries = rs->geoms.data();
    f.materials = rs->mats.data();
    f.lights = rs->lights.data();
    f.face_points = rs->fpts.data();
    f.face_normals = rs->fnrm.data();
    rs->flattened = true;
    return &rs->flat;
}

int ref_counts_enabled(void) {
#ifdef REF_COUNT_RAYS
    return 1;
#else
    return 0;
#endif
}

}  // extern "C"
