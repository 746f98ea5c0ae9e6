// synth.h — the synthetic scene of BASELINE.json configs[4] / SURVEY §8(d) "Config 5":
// a procedural height-field mesh (grid_cells^2 * 2 triangles; 708 -> 1 002 528) with
// per-vertex normals, num_spheres spheres (70 % diffuse+specular, 20 % mirror, 10 %
// glass) floating above it, 6 point + 2 directional shadow lights + ambient, 16:9 camera.
// PRNG: splitmix64.  The SAME description feeds two emitters: an in-memory Scene (fast
// path for timing) and .rti/.obj text (numbers printed %.17g, so parsing the text yields
// bit-identical doubles and the scene enters through the same parsers as any other).
#pragma once
#include <cstdint>
#include <string>

#include "scene_model.h"

namespace as2 {

void buildSyntheticScene(Scene& scene, int grid_cells, int num_spheres, uint64_t seed);
void writeSyntheticScene(const std::string& rti_path, const std::string& obj_path, int grid_cells,
                         int num_spheres, uint64_t seed);

}  // namespace as2
