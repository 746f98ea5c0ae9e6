// writers.h — PNGWriter with the reference's surface (src/writers.h:5-16).  The
// reference goes through vendored libpng; the box has no libpng headers, so the RGB8
// PNG container (IHDR/IDAT/IEND, filter 0 rows, CRC32) is written directly on zlib, with the
// deflate work spread over the host's threads.
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
    // The whole .png file image in memory.  The rows are cut into `threads` stripes deflated
    // concurrently and stitched into one zlib stream (at 8K the reference's
    // single-threaded libpng encode takes longer than the GPU needs for the frame).
    static std::vector<uint8_t> encodeRGB8(const uint8_t* rgb, int width, int height, int threads, int level);
    // hardware threads (AS2_PNG_THREADS overrides), at most one per 64 rows
    static int encoderThreads(int height);
private:
    std::string filename_;
};

}  // namespace as2
