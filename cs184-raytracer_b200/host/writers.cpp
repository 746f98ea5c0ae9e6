// writers.cpp — see writers.h.
#include "writers.h"

#include <zlib.h>

#include <cstdio>
#include <cstring>

namespace as2 {

std::vector<uint8_t> PNGWriter::convertToRGB8(const RasterImage& image) {
    const long n = image.size() * 3;
    const double* src = image.data();
    std::vector<uint8_t> out((size_t)n);
    for (long i = 0; i < n; i++) {
        double v = src[i];
        // std::min/std::max as Eigen's cwiseMin(1).cwiseMax(0) apply them; a NaN passes
        // through both and the x86 double->int cast of the reference build yields 0.
        v = (1.0 < v) ? 1.0 : v;
        v = (v < 0.0) ? 0.0 : v;
        if (v != v) v = 0.0;
        out[(size_t)i] = (uint8_t)(v * 255.0);
    }
    return out;
}

void PNGWriter::writeImage(const RasterImage& image) {
    std::vector<uint8_t> rgb = convertToRGB8(image);
    writeRGB8(rgb.data(), image.cols(), image.rows());
}

static void putU32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24));
    v.push_back((uint8_t)(x >> 16));
    v.push_back((uint8_t)(x >> 8));
    v.push_back((uint8_t)x);
}
static void putChunk(std::vector<uint8_t>& file, const char type[4], const uint8_t* data, size_t len) {
    putU32(file, (uint32_t)len);
    size_t start = file.size();
    file.insert(file.end(), type, type + 4);
    if (len) file.insert(file.end(), data, data + len);
    uint32_t crc = (uint32_t)crc32(0L, file.data() + start, (uInt)(len + 4));
    putU32(file, crc);
}

void PNGWriter::writeRGB8(const uint8_t* rgb, int width, int height) {
    if (width <= 0 || height <= 0) throw WriteException("invalid image dimensions");
    const size_t stride = (size_t)width * 3;
    std::vector<uint8_t> raw((stride + 1) * (size_t)height);
    for (int r = 0; r < height; r++) {
        raw[(stride + 1) * r] = 0;   // filter type None
        std::memcpy(&raw[(stride + 1) * r + 1], rgb + stride * r, stride);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK)
        throw WriteException("zlib compression failed");
    std::vector<uint8_t> file;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    file.insert(file.end(), sig, sig + 8);
    std::vector<uint8_t> ihdr;
    putU32(ihdr, (uint32_t)width);
    putU32(ihdr, (uint32_t)height);
    ihdr.push_back(8);   // bit depth
    ihdr.push_back(2);   // colour type RGB
    ihdr.push_back(0);
    ihdr.push_back(0);
    ihdr.push_back(0);
    putChunk(file, "IHDR", ihdr.data(), ihdr.size());
    putChunk(file, "IDAT", comp.data(), clen);
    putChunk(file, "IEND", nullptr, 0);
    FILE* f = std::fopen(filename_.c_str(), "wb");
    if (!f) throw WriteException("cannot open " + filename_ + " for writing");
    size_t wrote = std::fwrite(file.data(), 1, file.size(), f);
    int rc = std::fclose(f);
    if (wrote != file.size() || rc != 0) throw WriteException("short write to " + filename_);
}

}  // namespace as2
