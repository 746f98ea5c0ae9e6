// render_scene_b200.cpp — INTEGRATION.md "path B", compiled: the ONE translation unit a maintainer of the
// reference adds to run the per-pixel trace loop on a B200.  It defines Scene::renderScene (declared at
// src/scene.h:14, called at src/main.cpp:72) and is linked INSTEAD of the reference's src/scene.cpp; every
// other reference file (main.cpp, parsers.cpp, options.cpp, exceptions.cpp, geometry.cpp, writers.cpp + libpng)
// is compiled unmodified from where it lies.  tests/pathb/Makefile is the recipe, tests/test_gpu_round2.py
// (test_pathb_reference_main_on_the_b200) reproduces outputs/image-08.png through the resulting binary.
//
// The reference keeps its object graph private (src/scene.h:35-38) and has no accessor for it; a maintainer
// would add `friend` lines or getters.  This file gets the same access with `#define private public` around the
// reference headers in THIS translation unit only (class layout is unaffected) so that no reference file changes.
// every standard header the reference headers pull in is included BEFORE the access macro
#include <algorithm>
#include <array>
#include <atomic>
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
#include "options.h"
#undef private

#include "rt_b200.h"

namespace {

struct Flat {
    std::vector<rt_geometry> geoms;
    std::vector<rt_material> mats;
    std::vector<rt_light> lights;
    std::vector<double> fpts, fnrm;
    rt_scene desc;
};

// The reference object graph -> the ABI's flat descriptor (insertion order kept: it IS the tie-break order of
// src/scene.cpp:153-155 and src/geometry.cpp:108-110).  Lights and camera are handed over forward-transformed, as
// the reference's lazy caches would compute them (src/lights.h:28-33,56-61; src/rtbase.h:86-95).
void flatten(Scene& scene, Flat& f) {
    std::memset(&f.desc, 0, sizeof(f.desc));
    Camera& cam = scene.camera_;
    (void)cam.calculateViewingRay(0.5, 0.5);                     // fills the xf* caches
    for (int k = 0; k < 3; k++) {
        f.desc.camera.eye[k] = cam.xfEyePoint_[k];
        f.desc.camera.ll[k] = cam.xfLowerLeftPoint_[k];
        f.desc.camera.lr[k] = cam.xfLowerRightPoint_[k];
        f.desc.camera.ul[k] = cam.xfUpperLeftPoint_[k];
        f.desc.camera.ur[k] = cam.xfUpperRightPoint_[k];
    }
    for (auto& gp : scene.geometries_) {
        Geometry* g = gp.get();
        rt_geometry rg;
        rt_material m;
        std::memset(&rg, 0, sizeof(rg));
        std::memset(&m, 0, sizeof(m));
        for (int k = 0; k < 3; k++) {
            m.ka[k] = g->material_.ambientColor_[k];
            m.kd[k] = g->material_.diffuseColor_[k];
            m.ks[k] = g->material_.specularColor_[k];
            m.kr[k] = g->material_.reflectiveColor_[k];
            m.kt[k] = g->material_.translucencyColor_[k];
        }
        m.sp = g->material_.specularCoefficient_;
        m.ior = g->material_.indexOfRefractivity_;
        rg.material = (int32_t)f.mats.size();
        f.mats.push_back(m);
        const Matrix4d& F = g->forwardTransform().matrix();
        const Matrix4d& I = g->inverseTransform().matrix();
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) { rg.fwd[r * 4 + c] = F(r, c); rg.inv[r * 4 + c] = I(r, c); }
        rg.det = g->transformDeterminant();
        if (Sphere* s = dynamic_cast<Sphere*>(g)) {
            rg.type = RT_GEOM_SPHERE;
            for (int k = 0; k < 3; k++) rg.center[k] = s->center_[k];
            rg.radius = (double)s->radius_;
            rg.radius2 = (double)(s->radius_ * s->radius_);       // float product, src/geometry.cpp:54
        } else if (Mesh* me = dynamic_cast<Mesh*>(g)) {
            const bool has_bbox = (me->boundingBoxMin_ != me->boundingBoxMax_) && me->faces_.size() > 1;
            const bool is_tri = me->faces_.size() == 2 && me->boundingBoxMin_ == me->boundingBoxMax_;
            rg.type = is_tri ? RT_GEOM_TRI : RT_GEOM_MESH;
            rg.first_face = (int64_t)(f.fpts.size() / 9);
            rg.num_faces = (int64_t)me->faces_.size();
            rg.use_bbox = has_bbox ? 1 : 0;
            for (int k = 0; k < 3; k++) { rg.bbmin[k] = me->boundingBoxMin_[k]; rg.bbmax[k] = me->boundingBoxMax_[k]; }
            for (auto& face : me->faces_)
                for (int v = 0; v < 3; v++)
                    for (int k = 0; k < 3; k++) {
                        f.fpts.push_back(face.points_[v][k]);
                        f.fnrm.push_back(face.normals_[v][k]);
                    }
        }
        f.geoms.push_back(rg);
    }
    const Vector4d probe(0, 0, 0, 1);
    for (auto& lp : scene.lights_) {
        Light* l = lp.get();
        (void)l->calculateDistanceToLight(probe);
        (void)l->calculateDirectionToLight(probe);               // fill the xf* caches
        rt_light rl;
        std::memset(&rl, 0, sizeof(rl));
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
        f.lights.push_back(rl);
    }
    f.desc.num_geometries = (int32_t)f.geoms.size();
    f.desc.num_materials = (int32_t)f.mats.size();
    f.desc.num_lights = (int32_t)f.lights.size();
    f.desc.num_faces = (int64_t)(f.fpts.size() / 9);
    f.desc.geometries = f.geoms.data();
    f.desc.materials = f.mats.data();
    f.desc.lights = f.lights.data();
    f.desc.face_points = f.fpts.data();
    f.desc.face_normals = f.fnrm.data();
}

struct Thunk { Scene::ProgressHandler handler; };
void forward(int complete, int total, void* user) {
    Thunk* t = static_cast<Thunk*>(user);
    if (t->handler) t->handler(complete, total);
}

}  // namespace

// Replaces src/scene.cpp:10-59.  RasterImage (src/scene.h:11) is a row-major Eigen array of Color3d, i.e.
// double[rows][cols][3] contiguous: exactly what rt_render fills.
void Scene::renderScene(RasterImage& output, ProgressHandler phandler) {
    Flat flat;
    flatten(*this, flat);
    rt_context* ctx = nullptr;
    if (rt_create(-1, &ctx) != RT_OK) throw std::runtime_error(rt_last_error());
    rt_params p;
    std::memset(&p, 0, sizeof(p));
    p.width = (int32_t)output.cols();
    p.height = (int32_t)output.rows();
    p.bounce_depth = programOptions.bounceDepth_;                   // src/options.h:15
    p.intersection_only = programOptions.intersectionOnly_ ? 1 : 0;  // src/options.h:16
    p.tile_world = 1;
    Thunk thunk{phandler};
    int rc = rt_scene_upload(ctx, &flat.desc);
    if (rc == RT_OK) rc = rt_render(ctx, &p, reinterpret_cast<double*>(output.data()), forward, &thunk);
    std::string err = rc == RT_OK ? "" : rt_last_error();
    rt_destroy(ctx);
    if (rc != RT_OK) throw std::runtime_error(err);
}
