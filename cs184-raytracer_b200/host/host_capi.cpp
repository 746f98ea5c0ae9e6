// host_capi.cpp — a small C surface over the C++ host (parsers, object model,
// flattening, PNG writer) so the Python tests/bench can drive it through ctypes.  It
// adds no rendering logic: rendering goes through include/rt_b200.h.
#include <cstring>
#include <string>

#include "options.h"
#include "parsers.h"
#include "scene_model.h"
#include "synth.h"
#include "writers.h"

using namespace as2;

namespace {
void setError(char* err, int errlen, const std::string& msg) {
    if (!err || errlen <= 0) return;
    std::strncpy(err, msg.c_str(), (size_t)errlen - 1);
    err[errlen - 1] = 0;
}
}  // namespace

extern "C" {

// Parse the given .rti files into one Scene (like src/main.cpp:53-66).  NULL + err on failure.
void* as2_scene_load(const char** files, int nfiles, char* err, int errlen) {
    Scene* scene = new Scene();
    try {
        for (int i = 0; i < nfiles; i++) {
            RTIParser parser(*scene);
            parser.parseFile(files[i]);
        }
        if (!scene->hasCamera()) throw ParseException("At least one camera must be specified.");
    } catch (const std::exception& e) {
        setError(err, errlen, e.what());
        delete scene;
        return nullptr;
    }
    return scene;
}

// The synthetic BASELINE config-5 scene built in memory (see synth.h).
void* as2_scene_synthetic(int grid_cells, int num_spheres, uint64_t seed, char* err, int errlen) {
    Scene* scene = new Scene();
    try {
        buildSyntheticScene(*scene, grid_cells, num_spheres, seed);
    } catch (const std::exception& e) {
        setError(err, errlen, e.what());
        delete scene;
        return nullptr;
    }
    return scene;
}

// Same scene written as .rti + .obj text so it can enter through the parsers.
int as2_write_synthetic(const char* rti_path, const char* obj_path, int grid_cells, int num_spheres,
                        uint64_t seed, char* err, int errlen) {
    try {
        writeSyntheticScene(rti_path, obj_path, grid_cells, num_spheres, seed);
    } catch (const std::exception& e) {
        setError(err, errlen, e.what());
        return -1;
    }
    return 0;
}

void as2_scene_free(void* h) { delete static_cast<Scene*>(h); }

const rt_scene* as2_scene_flatten(void* h, char* err, int errlen) {
    try {
        return &static_cast<Scene*>(h)->flatten().desc;
    } catch (const std::exception& e) {
        setError(err, errlen, e.what());
        return nullptr;
    }
}

// Scene::renderScene through the C++ host (the call a user of the reference makes).
int as2_scene_render(void* h, int width, int height, int bounce_depth, int intersection_only,
                     double* rgb, char* err, int errlen) {
    Scene* scene = static_cast<Scene*>(h);
    programOptions.bounceDepth_ = bounce_depth;
    programOptions.intersectionOnly_ = intersection_only != 0;
    try {
        RasterImage image(height, width);
        scene->renderScene(image, nullptr);
        std::memcpy(rgb, image.data(), sizeof(double) * 3 * (size_t)width * height);
    } catch (const std::exception& e) {
        setError(err, errlen, e.what());
        return -1;
    }
    return 0;
}

int as2_write_png_rgb8(const char* path, const uint8_t* rgb, int width, int height, char* err, int errlen) {
    try {
        PNGWriter(path).writeRGB8(rgb, width, height);
    } catch (const std::exception& e) {
        setError(err, errlen, e.what());
        return -1;
    }
    return 0;
}

int as2_write_png_f64(const char* path, const double* rgb, int width, int height, char* err, int errlen) {
    try {
        RasterImage image(height, width);
        std::memcpy(image.data(), rgb, sizeof(double) * 3 * (size_t)width * height);
        PNGWriter(path).writeImage(image);
    } catch (const std::exception& e) {
        setError(err, errlen, e.what());
        return -1;
    }
    return 0;
}

// Encodes a PNG into a caller-owned buffer with an explicit thread count (tests / ingest
// benchmark).  Returns the file size, or -1 (error) / -2 (buffer too small: *needed is set).
int64_t as2_encode_png_rgb8(const uint8_t* rgb, int width, int height, int threads, uint8_t* out, int64_t out_cap,
                            int64_t* needed, char* err, int errlen) {
    try {
        std::vector<uint8_t> file = PNGWriter::encodeRGB8(rgb, width, height, threads, 6);
        if (needed) *needed = (int64_t)file.size();
        if ((int64_t)file.size() > out_cap) return -2;
        std::memcpy(out, file.data(), file.size());
        return (int64_t)file.size();
    } catch (const std::exception& e) {
        setError(err, errlen, e.what());
        return -1;
    }
}

// Quantise like src/writers.cpp:7 (host implementation used by the PNG writer).
void as2_quantize_rgb8(const double* rgb, int64_t n_values, uint8_t* out) {
    RasterImage image(1, (int)(n_values / 3));
    std::memcpy(image.data(), rgb, sizeof(double) * (size_t)n_values);
    std::vector<uint8_t> q = PNGWriter::convertToRGB8(image);
    std::memcpy(out, q.data(), q.size());
}

}  // extern "C"
