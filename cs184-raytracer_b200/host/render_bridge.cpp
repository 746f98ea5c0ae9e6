// render_bridge.cpp — Scene::renderScene, the drop-in seam (src/scene.h:14,
// src/main.cpp:72).  The reference runs its thread pool + traceRay here
// (src/scene.cpp:10-59); this implementation flattens the object graph and hands it to
// the CUDA library through the C ABI of include/rt_b200.h.  There is no CPU fallback:
// a missing/unsupported device surfaces as RenderException ("Error: ..." + exit 1).
#include "options.h"
#include "scene_model.h"

namespace as2 {

Scene::Scene() {}
Scene::~Scene() {
    if (warm_.joinable()) warm_.join();
    if (warmCtx_ && warmCtx_ != ctx_) rt_destroy(warmCtx_);
    if (ctx_) rt_destroy(ctx_);
}

void Scene::warmDeviceAsync() {
    if (ctx_ || warm_.joinable()) return;
    warm_ = std::thread([this]() {
        warmRc_ = rt_create(-1, &warmCtx_);
        if (warmRc_ != RT_OK) warmErr_ = rt_last_error();      // rt_last_error() is per thread
    });
}

rt_context* Scene::deviceContext() {
    if (warm_.joinable()) {
        warm_.join();
        if (warmRc_ != RT_OK) throw RenderException(warmErr_);
        ctx_ = warmCtx_;
    }
    if (!ctx_) {
        if (rt_create(-1, &ctx_) != RT_OK) throw RenderException(rt_last_error());
    }
    return ctx_;
}

namespace {
struct ProgressThunk {
    Scene::ProgressHandler handler;
};
void forwardProgress(int complete, int total, void* user) {
    auto* t = static_cast<ProgressThunk*>(user);
    if (t->handler) t->handler(complete, total);
}
rt_params paramsFromOptions(int width, int height) {
    rt_params p = rt_params();
    p.width = width;
    p.height = height;
    p.bounce_depth = programOptions.bounceDepth_;
    p.intersection_only = programOptions.intersectionOnly_ ? 1 : 0;
    p.tile_rank = 0;
    p.tile_world = 1;
    p.flags = programOptions.bruteForce_ ? RT_FLAG_BRUTE_FORCE : 0u;
    p.samples = programOptions.samples_;
    p.n_gpus = programOptions.gpus_;
    return p;
}
}  // namespace

void Scene::renderScene(RasterImage& output, ProgressHandler phandler) {
    rt_context* ctx = deviceContext();
    const bool hadFlat = (bool)flat_;
    const FlatScene& fs = flatten();
    if (!uploaded_ || !hadFlat) {
        if (rt_scene_upload(ctx, &fs.desc) != RT_OK) throw RenderException(rt_last_error());
        uploaded_ = true;
    }
    rt_params p = paramsFromOptions(output.cols(), output.rows());
    ProgressThunk thunk{phandler};
    if (rt_render(ctx, &p, output.data(), forwardProgress, &thunk) != RT_OK)
        throw RenderException(rt_last_error());
    rt_get_stats(ctx, &stats_);
}

void Scene::renderSceneRGB8(std::vector<uint8_t>& rgb8, int width, int height, ProgressHandler phandler) {
    rt_context* ctx = deviceContext();
    const bool hadFlat = (bool)flat_;
    const FlatScene& fs = flatten();
    if (!uploaded_ || !hadFlat) {
        if (rt_scene_upload(ctx, &fs.desc) != RT_OK) throw RenderException(rt_last_error());
        uploaded_ = true;
    }
    rt_params p = paramsFromOptions(width, height);
    rgb8.resize((size_t)width * height * 3);
    ProgressThunk thunk{phandler};
    if (rt_render_rgb8(ctx, &p, rgb8.data(), forwardProgress, &thunk) != RT_OK)
        throw RenderException(rt_last_error());
    rt_get_stats(ctx, &stats_);
}

}  // namespace as2
