// scene_model.h — the host object model: the drop-in surface of the reference's
// Scene / Geometry / Light / Camera / Material classes (src/scene.h:9-39,
// src/geometry.h:6-37, src/lights.h:3-75, src/rtbase.h:30-103) with the same public
// names, but with the trace loop removed: Scene::renderScene flattens the object graph
// into the C-ABI descriptor (include/rt_b200.h) and calls the CUDA library.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "rt_b200.h"
#include "vecmath.h"

namespace as2 {

// ---- error convention (src/exceptions.h:6-30) ------------------------------
class ParseException : public std::runtime_error {
public:
    explicit ParseException(const std::string& msg, int lineno = -1)
        : std::runtime_error(format(msg, lineno)), lineno_(lineno) {}
    int line() const { return lineno_; }
    static void showWarning(const std::string& msg, int lineno = -1);
    static std::string format(const std::string& msg, int lineno);
private:
    int lineno_;
};
class MathException : public std::runtime_error {
public:
    using std::runtime_error::runtime_error;
};
class WriteException : public std::runtime_error {
public:
    using std::runtime_error::runtime_error;
};
// Raised when the device library reports a failure; main prints it as "Error: ...".
class RenderException : public std::runtime_error {
public:
    using std::runtime_error::runtime_error;
};

typedef std::array<double, 3> Color3d;

// ---- Material (src/rtbase.h:30-39); zero-initialised (the reference leaves it
// uninitialised before the first `mat`, which is UB — SURVEY App. D) -----------
struct Material {
    Color3d ambientColor_{{0, 0, 0}};
    Color3d diffuseColor_{{0, 0, 0}};
    Color3d specularColor_{{0, 0, 0}};
    Color3d reflectiveColor_{{0, 0, 0}};
    double specularCoefficient_ = 0;
    Color3d translucencyColor_{{0, 0, 0}};
    double indexOfRefractivity_ = 0;
};

// ---- Transformable (src/rtbase.h:41-64) --------------------------------------
class Transformable {
public:
    const Affine& forwardTransform() const { return fwd_; }
    const Affine& inverseTransform() const { return inv_; }
    double transformDeterminant() const { return det_; }
    void forwardTransform(const Affine& xf) {
        fwd_ = xf;
        inv_ = xf.inverse();
        det_ = fwd_.determinant();
    }
    void inverseTransform(const Affine& xf) {
        fwd_ = xf.inverse();
        inv_ = xf;
        det_ = fwd_.determinant();
    }
private:
    Affine fwd_, inv_;
    double det_ = 1.0;
};

// ---- Camera (src/rtbase.h:66-103): the bilinear image plane; the transformed
// points are computed eagerly at flatten time, never lazily ---------------------
class Camera : public Transformable {
public:
    void eyePoint(const Vec4& p) { eye_ = p; }
    void lowerLeftPoint(const Vec4& p) { ll_ = p; }
    void lowerRightPoint(const Vec4& p) { lr_ = p; }
    void upperLeftPoint(const Vec4& p) { ul_ = p; }
    void upperRightPoint(const Vec4& p) { ur_ = p; }
    void fill(rt_camera& out) const;
private:
    Vec4 eye_, ll_, lr_, ul_, ur_;
};

// ---- Lights (src/lights.h:3-75) ----------------------------------------------
class Light : public Transformable {
public:
    virtual ~Light() = default;
    virtual void fill(rt_light& out) const = 0;
    Color3d color_{{0, 0, 0}};
};
class PointLight : public Light {
public:
    void point(const Vec4& p) { point_ = p; }
    void fill(rt_light& out) const override;
    double falloffExponent_ = 0;
private:
    Vec4 point_;
};
class DirectionalLight : public Light {
public:
    void direction(const Vec4& d) { direction_ = d; }
    void fill(rt_light& out) const override;
private:
    Vec4 direction_;
};
class AmbientLight : public Light {
public:
    void fill(rt_light& out) const override;
};

// ---- Geometry (src/geometry.h:6-37) --------------------------------------------
class Geometry : public Transformable {
public:
    virtual ~Geometry() = default;
    Material material_;
};
class Sphere : public Geometry {
public:
    Vec4 center_;
    float radius_ = 0;   // float on purpose: src/geometry.h:22
};
class Mesh : public Geometry {
public:
    struct Face {
        std::array<Vec4, 3> points_, normals_;
    };
    // Two one-sided faces displaced by -/+ eps along the normal (src/geometry.cpp:128-143).
    void addTriangle(const std::array<Vec4, 3>& points);
    // Object-space AABB over all face vertices (src/geometry.cpp:145-162).
    void updateBoundingBox();
    const Vec4& boundingBoxMin() const { return bbmin_; }
    const Vec4& boundingBoxMax() const { return bbmax_; }
    std::vector<Face> faces_;
    bool fromTriStatement_ = false;
private:
    Vec4 bbmin_, bbmax_;
};

// ---- the flattened scene (owner of the arrays rt_scene points into) -------------
struct FlatScene {
    std::vector<rt_geometry> geometries;
    std::vector<rt_material> materials;
    std::vector<rt_light> lights;
    std::vector<double> face_points, face_normals;
    rt_scene desc;
    FlatScene() { desc = rt_scene(); }
    FlatScene(const FlatScene&) = delete;
    FlatScene& operator=(const FlatScene&) = delete;
    void seal();   // point desc at the vectors
};

// Row-major H x W image of Color3d == Scene::RasterImage (src/scene.h:11).
class RasterImage {
public:
    RasterImage(int rows, int cols) : rows_(rows), cols_(cols), px_((size_t)rows * cols) {}
    int rows() const { return rows_; }
    int cols() const { return cols_; }
    long size() const { return (long)rows_ * cols_; }
    Color3d& operator()(int r, int c) { return px_[(size_t)r * cols_ + c]; }
    const Color3d& operator()(int r, int c) const { return px_[(size_t)r * cols_ + c]; }
    Color3d& operator()(long i) { return px_[(size_t)i]; }
    double* data() { return px_[0].data(); }
    const double* data() const { return px_[0].data(); }
private:
    int rows_, cols_;
    std::vector<Color3d> px_;
};

class Scene {
public:
    typedef as2::RasterImage RasterImage;
    typedef void (*ProgressHandler)(int complete, int total);

    Scene();
    ~Scene();
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;

    // The drop-in seam (src/scene.h:14, called at src/main.cpp:72).  Reads the same
    // global the reference reads: programOptions.{bounceDepth_, intersectionOnly_}.
    void renderScene(RasterImage& output, ProgressHandler phandler = nullptr);
    // Same render with the PNG writer's quantisation done on the device (src/writers.cpp:7).
    void renderSceneRGB8(std::vector<uint8_t>& rgb8, int width, int height, ProgressHandler phandler = nullptr);

    bool hasCamera() const { return hasCamera_; }
    const Camera& camera() const { return camera_; }
    void camera(const Camera& cam) { hasCamera_ = true; camera_ = cam; flat_.reset(); }
    void addGeometry(std::unique_ptr<Geometry>&& g) { geometries_.push_back(std::move(g)); flat_.reset(); }
    void addLight(std::unique_ptr<Light>&& l) { lights_.push_back(std::move(l)); flat_.reset(); }
    size_t geometryCount() const { return geometries_.size(); }
    size_t lightCount() const { return lights_.size(); }

    // Flatten the object graph into the ABI descriptor (cached until the scene changes).
    const FlatScene& flatten();
    const rt_stats& lastStats() const { return stats_; }
    // Starts creating the device context on a helper thread (CUDA initialisation takes far longer than
    // parsing a small scene); renderScene joins it.  Optional: without it the context is created on first use.
    void warmDeviceAsync();

private:
    rt_context* deviceContext();
    std::thread warm_;
    rt_context* warmCtx_ = nullptr;
    int warmRc_ = 0;
    std::string warmErr_;
    bool hasCamera_ = false;
    Camera camera_;
    std::vector<std::unique_ptr<Geometry>> geometries_;
    std::vector<std::unique_ptr<Light>> lights_;
    std::unique_ptr<FlatScene> flat_;
    rt_context* ctx_ = nullptr;
    bool uploaded_ = false;
    rt_stats stats_{};
};

}  // namespace as2
