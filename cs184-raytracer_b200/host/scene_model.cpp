// scene_model.cpp — object model bookkeeping and the flattening into rt_scene.
#include "scene_model.h"

#include <cfloat>
#include <cstring>
#include <iostream>
#include <limits>

namespace as2 {

std::string ParseException::format(const std::string& msg, int lineno) {
    if (lineno <= 0) return msg;
    return "line " + std::to_string(lineno) + ": " + msg;
}
void ParseException::showWarning(const std::string& msg, int lineno) {
    std::cerr << "Warning: " << format(msg, lineno) << std::endl;
}

// ---- `tri`: the +-eps face pair (src/geometry.cpp:128-143) ----------------------
// n = normalize((v1-v0) x (v2-v0)) as 4-vectors; eps = DBL_EPSILON * |v0+v1+v2| / 3
// (4-vector norm, w = 3 included); face 0 = points - eps n with normals -n,
// face 1 = points + eps n with normals +n.
void Mesh::addTriangle(const std::array<Vec4, 3>& pts) {
    const Vec4 e1 = pts[1] - pts[0], e2 = pts[2] - pts[0];
    const Vec4 n = normalized4(Vec4::dir(cross(e1.head(), e2.head())));
    const Vec4 sum = (pts[0] + pts[1]) + pts[2];
    const Vec4 ep = (DBL_EPSILON * norm4(sum) / 3) * n;
    for (int s = -1; s <= 1; s += 2) {
        Face f;
        const Vec4 sn = (double)s * n, sep = (double)s * ep;
        for (int k = 0; k < 3; k++) {
            f.points_[k] = pts[k] + sep;
            f.normals_[k] = sn;
        }
        faces_.push_back(f);
    }
}

void Mesh::updateBoundingBox() {
    if (faces_.empty()) {
        bbmin_ = Vec4();
        bbmax_ = Vec4();
        return;
    }
    const double inf = std::numeric_limits<double>::infinity();
    Vec4 lo(inf, inf, inf, inf), hi(-inf, -inf, -inf, -inf);
    for (const Face& f : faces_)
        for (const Vec4& p : f.points_)
            for (int k = 0; k < 4; k++) {
                if (p[k] < lo[k]) lo[k] = p[k];
                if (p[k] > hi[k]) hi[k] = p[k];
            }
    if (lo.w != 1.0 || hi.w != 1.0) throw MathException("non-unity-homogeneous bounding box");
    bbmin_ = lo;
    bbmax_ = hi;
}

// ---- eager (race-free) replacements of the reference's lazy xf caches ----------
static void put3(double* dst, const Vec4& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

void Camera::fill(rt_camera& out) const {
    const Affine& T = forwardTransform();
    put3(out.eye, T.apply(eye_));
    put3(out.ll, T.apply(ll_));
    put3(out.lr, T.apply(lr_));
    put3(out.ul, T.apply(ul_));
    put3(out.ur, T.apply(ur_));
}
void PointLight::fill(rt_light& out) const {
    out.type = RT_LIGHT_POINT;
    put3(out.v, forwardTransform().apply(point_));
    out.falloff = falloffExponent_;
}
void DirectionalLight::fill(rt_light& out) const {
    out.type = RT_LIGHT_DIRECTIONAL;
    put3(out.v, forwardTransform().apply(direction_));
}
void AmbientLight::fill(rt_light& out) const { out.type = RT_LIGHT_AMBIENT; }

void FlatScene::seal() {
    desc.num_geometries = (int32_t)geometries.size();
    desc.num_materials = (int32_t)materials.size();
    desc.num_lights = (int32_t)lights.size();
    desc.num_faces = (int64_t)(face_points.size() / 9);
    desc.geometries = geometries.data();
    desc.materials = materials.data();
    desc.lights = lights.data();
    desc.face_points = face_points.data();
    desc.face_normals = face_normals.data();
}

const FlatScene& Scene::flatten() {
    if (flat_) return *flat_;
    std::unique_ptr<FlatScene> fs(new FlatScene());
    camera_.fill(fs->desc.camera);
    size_t nfaces = 0;
    for (auto& g : geometries_)
        if (auto* me = dynamic_cast<Mesh*>(g.get())) nfaces += me->faces_.size();
    fs->face_points.reserve(nfaces * 9);
    fs->face_normals.reserve(nfaces * 9);
    fs->geometries.reserve(geometries_.size());
    fs->materials.reserve(geometries_.size());
    for (auto& gp : geometries_) {
        Geometry* g = gp.get();
        rt_geometry rg;
        std::memset(&rg, 0, sizeof(rg));
        rt_material m;
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
        rg.material = (int32_t)fs->materials.size();
        fs->materials.push_back(m);
        std::memcpy(rg.fwd, g->forwardTransform().m, sizeof(rg.fwd));
        std::memcpy(rg.inv, g->inverseTransform().m, sizeof(rg.inv));
        rg.det = g->transformDeterminant();
        if (auto* s = dynamic_cast<Sphere*>(g)) {
            rg.type = RT_GEOM_SPHERE;
            put3(rg.center, s->center_);
            rg.radius = (double)s->radius_;
            rg.radius2 = (double)(s->radius_ * s->radius_);   // float product, then widened (src/geometry.cpp:54)
        } else if (auto* me = dynamic_cast<Mesh*>(g)) {
            rg.type = me->fromTriStatement_ ? RT_GEOM_TRI : RT_GEOM_MESH;
            rg.first_face = (int64_t)(fs->face_points.size() / 9);
            rg.num_faces = (int64_t)me->faces_.size();
            const Vec4& lo = me->boundingBoxMin();
            const Vec4& hi = me->boundingBoxMax();
            bool differ = lo.x != hi.x || lo.y != hi.y || lo.z != hi.z || lo.w != hi.w;
            rg.use_bbox = (differ && me->faces_.size() > 1) ? 1 : 0;   // src/geometry.cpp:72
            put3(rg.bbmin, lo);
            put3(rg.bbmax, hi);
            for (const Mesh::Face& f : me->faces_)
                for (int v = 0; v < 3; v++)
                    for (int k = 0; k < 3; k++) {
                        fs->face_points.push_back(f.points_[v][k]);
                        fs->face_normals.push_back(f.normals_[v][k]);
                    }
        } else {
            throw MathException("unknown geometry class");
        }
        fs->geometries.push_back(rg);
    }
    for (auto& lp : lights_) {
        rt_light rl;
        std::memset(&rl, 0, sizeof(rl));
        for (int k = 0; k < 3; k++) rl.color[k] = lp->color_[k];
        lp->fill(rl);
        fs->lights.push_back(rl);
    }
    fs->seal();
    flat_ = std::move(fs);
    return *flat_;
}

}  // namespace as2
