// writers.h — PNGWriter with the reference's surface (src/writers.h:5-16).  The
// reference goes through vendored libpng; the box has no libpng headers, so the RGB8
// PNG container (IHDR/IDAT/IEND, filter 0 rows, CRC32) is written directly on zlib.
// What is compared with the goldens is the DECODED pixels, and the quantisation rule is
// the reference's: uint8 = (uint8_t)(clamp(v,0,1) * 255.0), truncating (src/writers.cpp:7).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "scene_model.h"

namespace as2 {

class PNGWriter {
public:
    explicit PNGWriter(std::string filename) : filename_(std::move(filename)) {}
    void writeImage(const RasterImage& image);
    // Already-quantised pixels (device-side quantisation path), row-major RGB8.
    void writeRGB8(const uint8_t* rgb, int width, int height);
    static std::vector<uint8_t> convertToRGB8(const RasterImage& image);
private:
    std::string filename_;
};

}  // namespace as2
