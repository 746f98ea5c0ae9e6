// writers.cpp — see writers.h.
#include "writers.h"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>

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

// One horizontal stripe of the image, deflated on its own thread as RAW deflate data that ends
// on a byte boundary (Z_SYNC_FLUSH; the last stripe ends the stream with Z_FINISH), so the
// stripes concatenate into one valid zlib stream (the pigz construction).  adler/crc of the
// pieces are merged with adler32_combine / crc32_combine.
namespace {
struct Stripe {
    int row0 = 0, rows = 0;
    std::vector<uint8_t> comp;
    uLong adler = 1;         // adler32 of this stripe's filtered bytes
    uLong crc = 0;           // crc32 of comp
    bool ok = true;
};

void deflateStripe(const uint8_t* rgb, int width, Stripe& s, bool last, int level) {
    const size_t stride = (size_t)width * 3, rawlen = (stride + 1) * (size_t)s.rows;
    std::vector<uint8_t> raw(rawlen);
    for (int r = 0; r < s.rows; r++) {
        raw[(stride + 1) * r] = 0;   // filter type None
        std::memcpy(&raw[(stride + 1) * r + 1], rgb + stride * (size_t)(s.row0 + r), stride);
    }
    z_stream z;
    std::memset(&z, 0, sizeof(z));
    if (deflateInit2(&z, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { s.ok = false; return; }
    s.comp.resize(deflateBound(&z, (uLong)rawlen) + 16);
    size_t in_off = 0;
    z.next_out = s.comp.data();
    z.avail_out = (uInt)std::min<size_t>(s.comp.size(), 0x7fffffffu);
    // zlib counts in 32-bit uInt: feed the input in < 2 GiB pieces (a stripe is far smaller)
    while (true) {
        const size_t piece = std::min<size_t>(rawlen - in_off, (size_t)1 << 30);
        z.next_in = raw.data() + in_off;
        z.avail_in = (uInt)piece;
        in_off += piece;
        const bool final_piece = in_off == rawlen;
        const int rc = deflate(&z, final_piece ? (last ? Z_FINISH : Z_SYNC_FLUSH) : Z_NO_FLUSH);
        if (rc == Z_STREAM_ERROR || (final_piece && last && rc != Z_STREAM_END) || z.avail_in != 0) {
            s.ok = false;
            break;
        }
        if (final_piece) break;
    }
    s.comp.resize((size_t)(z.next_out - s.comp.data()));
    deflateEnd(&z);
    s.adler = adler32(1L, Z_NULL, 0);
    for (size_t off = 0; off < rawlen; off += (size_t)1 << 30)
        s.adler = adler32(s.adler, raw.data() + off, (uInt)std::min<size_t>(rawlen - off, (size_t)1 << 30));
    s.crc = crc32(0L, Z_NULL, 0);
    for (size_t off = 0; off < s.comp.size(); off += (size_t)1 << 30)
        s.crc = crc32(s.crc, s.comp.data() + off, (uInt)std::min<size_t>(s.comp.size() - off, (size_t)1 << 30));
}
}  // namespace

int PNGWriter::encoderThreads(int height) {
    const char* env = std::getenv("AS2_PNG_THREADS");
    int n = env ? std::atoi(env) : (int)std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    // below ~64 rows a stripe's deflate window has nothing to work with
    const int by_rows = std::max(1, height / 64);
    return std::min(n, by_rows);
}

std::vector<uint8_t> PNGWriter::encodeRGB8(const uint8_t* rgb, int width, int height, int threads, int level) {
    if (width <= 0 || height <= 0) throw WriteException("invalid image dimensions");
    if (threads < 1) threads = 1;
    if (threads > height) threads = height;
    std::vector<Stripe> stripes((size_t)threads);
    for (int t = 0; t < threads; t++) {
        const int r0 = (int)((long long)height * t / threads), r1 = (int)((long long)height * (t + 1) / threads);
        stripes[(size_t)t].row0 = r0;
        stripes[(size_t)t].rows = r1 - r0;
    }
    if (threads == 1) {
        deflateStripe(rgb, width, stripes[0], true, level);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back(deflateStripe, rgb, width, std::ref(stripes[(size_t)t]), t == threads - 1, level);
        for (std::thread& th : pool) th.join();
    }
    size_t clen = 2 + 4;
    for (const Stripe& s : stripes) {
        if (!s.ok) throw WriteException("zlib compression failed");
        clen += s.comp.size();
    }
    if (clen > 0x7fffffffu) throw WriteException("image too large for a single IDAT chunk");
    const size_t stride = (size_t)width * 3;
    // zlib wrapper: CMF/FLG for deflate, 32 KiB window, default compression; adler32 trailer
    const uint8_t zhdr[2] = {0x78, 0x9c};
    uLong adler = adler32(0L, Z_NULL, 0);
    uLong crc = crc32(0L, Z_NULL, 0);
    crc = crc32(crc, (const Bytef*)"IDAT", 4);
    crc = crc32(crc, zhdr, 2);
    for (const Stripe& s : stripes) {
        adler = adler32_combine(adler, s.adler, (z_off_t)((stride + 1) * (size_t)s.rows));
        crc = crc32_combine(crc, s.crc, (z_off_t)s.comp.size());
    }
    uint8_t trailer[4] = {(uint8_t)(adler >> 24), (uint8_t)(adler >> 16), (uint8_t)(adler >> 8), (uint8_t)adler};
    crc = crc32(crc, trailer, 4);

    std::vector<uint8_t> file;
    file.reserve(clen + 64);
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
    putU32(file, (uint32_t)clen);
    file.insert(file.end(), {'I', 'D', 'A', 'T'});
    file.insert(file.end(), zhdr, zhdr + 2);
    for (const Stripe& s : stripes) file.insert(file.end(), s.comp.begin(), s.comp.end());
    file.insert(file.end(), trailer, trailer + 4);
    putU32(file, (uint32_t)crc);
    putChunk(file, "IEND", nullptr, 0);
    return file;
}

void PNGWriter::writeRGB8(const uint8_t* rgb, int width, int height) {
    std::vector<uint8_t> file = encodeRGB8(rgb, width, height, encoderThreads(height), 6);
    FILE* f = std::fopen(filename_.c_str(), "wb");
    if (!f) throw WriteException("cannot open " + filename_ + " for writing");
    size_t wrote = std::fwrite(file.data(), 1, file.size(), f);
    int rc = std::fclose(f);
    if (wrote != file.size() || rc != 0) throw WriteException("short write to " + filename_);
}

}  // namespace as2
