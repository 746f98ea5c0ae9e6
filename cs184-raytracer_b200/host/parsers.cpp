// parsers.cpp — see parsers.h.  Grammar reference: src/parsers.cpp:5-19 (statement
// table), :24-91 (tokens / numbers), :93-251 (.rti), :253-374 (.obj).
#include "parsers.h"

#include <libgen.h>

#include <algorithm>
#include <exception>
#include <thread>
#include <cstdint>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

namespace as2 {

static inline bool isSpaceC(char c) {
    return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r';
}

bool LineLexer::next(const char*& tb, const char*& te) {
    while (p_ < end_ && isSpaceC(*p_)) p_++;
    if (p_ >= end_) return false;
    if (*p_ == '"') {
        const char* q = p_ + 1;
        const char* close = (const char*)std::memchr(q, '"', (size_t)(end_ - q));
        if (!close) throw ParseException("unclosed quotes", lineno_);
        tb = q;
        te = close;
        p_ = close + 1;
        return tb != te;          // "" is an empty token and ends the list
    }
    tb = p_;
    while (p_ < end_ && !isSpaceC(*p_)) p_++;
    te = p_;
    if (p_ < end_) p_++;          // the delimiter is consumed with the token
    if (*tb == '#') {             // comment: drop the rest of the line
        p_ = end_;
        return false;
    }
    return true;
}

double parseNumber(const char* tb, const char* te, int lineno) {
    char stackbuf[64];
    size_t n = (size_t)(te - tb);
    std::unique_ptr<char[]> heap;
    char* buf = stackbuf;
    if (n >= sizeof(stackbuf)) {
        heap.reset(new char[n + 1]);
        buf = heap.get();
    }
    std::memcpy(buf, tb, n);
    buf[n] = 0;
    char* endp = nullptr;
    errno = 0;
    double v = std::strtod(buf, &endp);
    if (endp == buf || errno == ERANGE) throw ParseException(std::string("invalid number ") + buf, lineno);
    return v;
}

std::vector<char> slurpFile(const std::string& filename) {
    FILE* f = std::fopen(filename.c_str(), "rb");
    if (!f) throw ParseException("file not found: " + filename);
    std::vector<char> data;
    // regular file: one allocation, one read; anything else (pipe, /proc): grow as we go
    if (std::fseek(f, 0, SEEK_END) == 0) {
        const long size = std::ftell(f);
        if (size > 0 && std::fseek(f, 0, SEEK_SET) == 0) {
            data.resize((size_t)size);
            const size_t got = std::fread(data.data(), 1, (size_t)size, f);
            data.resize(got);
        } else {
            std::rewind(f);
        }
    }
    char chunk[1 << 16];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0) data.insert(data.end(), chunk, chunk + got);
    std::fclose(f);
    return data;
}

// Calls fn(lineBegin, lineEnd, lineno) for every '\n'-terminated line (and a last
// unterminated one), numbering from 1 like the reference's getline loop.
template <typename Fn>
static void forEachLine(const std::vector<char>& data, Fn fn) {
    const char* p = data.data();
    const char* end = p + data.size();
    int lineno = 1;
    while (p < end) {
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        fn(p, le, lineno);
        lineno++;
        p = nl ? nl + 1 : end;
    }
}

static std::string directoryOf(const std::string& path) {
    std::vector<char> tmp(path.begin(), path.end());
    tmp.push_back(0);
    return std::string(::dirname(tmp.data()));
}

// ---- .rti -------------------------------------------------------------------
namespace {
enum class Stmt { Cam, Sph, Tri, Ltp, Ltd, Lta, Mat, Xft, Xfr, Xfs, Xfz };
struct StmtInfo {
    const char* name;
    Stmt kind;
    int pmin, pmax;
};
const StmtInfo kStatements[] = {
    {"cam", Stmt::Cam, 15, 15}, {"sph", Stmt::Sph, 4, 4},   {"tri", Stmt::Tri, 9, 9},
    {"ltp", Stmt::Ltp, 6, 7},   {"ltd", Stmt::Ltd, 6, 6},   {"lta", Stmt::Lta, 3, 3},
    {"mat", Stmt::Mat, 13, 17}, {"xft", Stmt::Xft, 3, 3},   {"xfr", Stmt::Xfr, 3, 3},
    {"xfs", Stmt::Xfs, 3, 3},   {"xfz", Stmt::Xfz, 0, 0},
};
const StmtInfo* findStatement(const std::string& kw) {
    for (const StmtInfo& s : kStatements)
        if (kw == s.name) return &s;
    return nullptr;
}
}  // namespace

void RTIParser::parseFile(std::string filename) {
    std::vector<char> data = slurpFile(filename);
    forEachLine(data, [&](const char* lb, const char* le, int lineno) {
        LineLexer lex(lb, le, lineno);
        const char *tb, *te;
        if (!lex.next(tb, te)) return;   // blank or comment line
        statement(std::string(tb, te), lex, filename);
    });
}

void RTIParser::statement(const std::string& kw, LineLexer& lex, const std::string& filename) {
    const int lineno = lex.lineno();
    const char *tb, *te;
    if (kw == "obj") {
        if (!lex.next(tb, te)) throw ParseException("obj requires a filename", lineno);
        std::string objname(tb, te);
        if (objname[0] != '/') objname = directoryOf(filename) + "/" + objname;
        std::unique_ptr<Mesh> mesh(new Mesh());
        mesh->forwardTransform(transform_);
        mesh->material_ = material_;
        OBJParser(*mesh).parseFile(objname);
        mesh->updateBoundingBox();
        scene_.addGeometry(std::move(mesh));
        return;
    }
    const StmtInfo* info = findStatement(kw);
    if (!info) {
        ParseException::showWarning("unknown line type " + kw, lineno);
        return;
    }
    std::vector<double> a;
    while (lex.next(tb, te)) a.push_back(parseNumber(tb, te, lineno));
    if ((int)a.size() < info->pmin) {
        std::string msg = kw + " requires " + (info->pmin == info->pmax ? "" : "at least ") +
                          std::to_string(info->pmin) + " parameters";
        throw ParseException(msg, lineno);
    }
    if ((int)a.size() > info->pmax) ParseException::showWarning("extra parameters found", lineno);
    if ((int)a.size() < info->pmax) a.resize((size_t)info->pmax, 0.0);

    auto color = [&](int o) { return Color3d{{a[o], a[o + 1], a[o + 2]}}; };
    auto vec3 = [&](int o) { return Vec3(a[o], a[o + 1], a[o + 2]); };
    auto point = [&](int o) { return Vec4(a[o], a[o + 1], a[o + 2], 1.0); };

    switch (info->kind) {
        case Stmt::Xfz: transform_.setIdentity(); break;
        case Stmt::Xft: transform_.translate(vec3(0)); break;
        case Stmt::Xfs: transform_.scale(vec3(0)); break;
        case Stmt::Xfr: {
            Vec3 r = vec3(0);
            if (!r.isZero()) transform_.rotate(norm3(r) * (2 * M_PI / 360.0), normalized3(r));
            break;
        }
        case Stmt::Mat:
            material_.ambientColor_ = color(0);
            material_.diffuseColor_ = color(3);
            material_.specularColor_ = color(6);
            material_.specularCoefficient_ = a[9];
            material_.reflectiveColor_ = color(10);
            material_.translucencyColor_ = color(13);
            material_.indexOfRefractivity_ = a[16];
            break;
        case Stmt::Cam: {
            Camera cam;
            cam.forwardTransform(transform_);
            cam.eyePoint(point(0));
            cam.lowerLeftPoint(point(3));
            cam.lowerRightPoint(point(6));
            cam.upperLeftPoint(point(9));
            cam.upperRightPoint(point(12));
            scene_.camera(cam);
            break;
        }
        case Stmt::Sph: {
            std::unique_ptr<Sphere> s(new Sphere());
            s->forwardTransform(transform_);
            s->material_ = material_;
            s->center_ = point(0);
            s->radius_ = (float)a[3];
            scene_.addGeometry(std::move(s));
            break;
        }
        case Stmt::Tri: {
            std::unique_ptr<Mesh> m(new Mesh());
            m->forwardTransform(transform_);
            m->material_ = material_;
            m->fromTriStatement_ = true;
            m->addTriangle({{point(0), point(3), point(6)}});
            scene_.addGeometry(std::move(m));
            break;
        }
        case Stmt::Ltp: {
            std::unique_ptr<PointLight> l(new PointLight());
            l->forwardTransform(transform_);
            l->point(point(0));
            l->color_ = color(3);
            l->falloffExponent_ = a[6];
            scene_.addLight(std::move(l));
            break;
        }
        case Stmt::Ltd: {
            Vec3 d = vec3(0);
            if (d.isZero()) throw ParseException("zero direction specified", lineno);
            std::unique_ptr<DirectionalLight> l(new DirectionalLight());
            l->forwardTransform(transform_);
            l->direction(Vec4::dir(normalized3(d)));
            l->color_ = color(3);
            scene_.addLight(std::move(l));
            break;
        }
        case Stmt::Lta: {
            std::unique_ptr<AmbientLight> l(new AmbientLight());
            l->forwardTransform(transform_);
            l->color_ = color(0);
            scene_.addLight(std::move(l));
            break;
        }
    }
}

// ---- .obj -------------------------------------------------------------------
namespace {
struct Corner {
    int v = 0, vt = 0, vn = 0;
};

// "i", "i/j", "i//k", "i/j/k": at most three '/'-separated fields, each std::stoi-like.
Corner parseCorner(const char* tb, const char* te, int lineno) {
    int idx[3] = {0, 0, 0};
    int count = 0;
    const char* p = tb;
    while (count < 3 && p < te) {
        const char* slash = (const char*)std::memchr(p, '/', (size_t)(te - p));
        const char* fe = slash ? slash : te;
        int value = 0;
        if (fe > p) {
            std::string part(p, fe);
            char* endp = nullptr;
            errno = 0;
            long lv = std::strtol(part.c_str(), &endp, 10);
            if (endp == part.c_str() || errno == ERANGE || lv < INT32_MIN || lv > INT32_MAX)
                throw ParseException("invalid integer " + part, lineno);
            if (lv <= 0) throw ParseException("index must be positive", lineno);
            value = (int)lv;
        }
        idx[count++] = value;
        p = fe + 1;
    }
    Corner c;
    c.v = idx[0];
    c.vt = idx[1];
    c.vn = idx[2];
    return c;
}
}  // namespace

// ---- .obj, one contiguous range of lines -------------------------------------------
// Pass 1 (parseObjRange) turns text into numbers without looking at any other part of the
// file: vertices, normals, and for every `f` line its corner indices plus how many vertices /
// normals the range had defined before it.  Pass 2 (buildObjFaces) resolves the indices
// against the whole file's vertex arrays and does the reference's per-face work
// (src/parsers.cpp:279-326).  The serial parser runs both over one range covering the file;
// the parallel parser runs them per chunk on the host's threads.
namespace {
struct ObjWarning {
    int lineno;
    std::string msg;
};
struct ObjFaceRec {
    int lineno;
    uint32_t first, count;          // corners [first, first + count)
    uint32_t nv_before, nn_before;  // vertices / normals defined earlier in this range
    bool partial;                   // a later corner of the line failed to parse: validate these, build nothing
};
struct ObjRange {
    const char *begin = nullptr, *end = nullptr;
    int first_lineno = 1;
    std::vector<Vec4> v, vn;
    std::vector<Corner> corners;
    std::vector<ObjFaceRec> faces;
    std::vector<ObjWarning> warnings;     // in line order
    std::vector<Mesh::Face> built;
    size_t v_base = 0, vn_base = 0;       // vertices / normals defined before this range
    std::exception_ptr error;             // first ParseException of the range, if any
};

template <typename Fn>
void forEachLineIn(const char* p, const char* end, int lineno, Fn fn) {
    while (p < end) {
        const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
        const char* le = nl ? nl : end;
        fn(p, le, lineno);
        lineno++;
        p = nl ? nl + 1 : end;
    }
}

void parseObjRange(ObjRange& r) {
    std::vector<std::pair<const char*, const char*>> spans;
    double nums[8];
    try {
        forEachLineIn(r.begin, r.end, r.first_lineno, [&](const char* lb, const char* le, int lineno) {
            LineLexer lex(lb, le, lineno);
            const char *tb, *te;
            if (!lex.next(tb, te)) return;
            const size_t klen = (size_t)(te - tb);
            if (klen == 1 && tb[0] == 'f') {
                // the reference tokenises the whole line first, checks the count, then
                // parses and validates corner by corner (src/parsers.cpp:279-326)
                spans.clear();
                while (lex.next(tb, te)) spans.push_back({tb, te});
                if (spans.size() < 3) throw ParseException("f requires at least 3 vertices", lineno);
                ObjFaceRec f;
                f.lineno = lineno;
                f.first = (uint32_t)r.corners.size();
                f.count = (uint32_t)spans.size();
                f.nv_before = (uint32_t)r.v.size();
                f.nn_before = (uint32_t)r.vn.size();
                f.partial = false;
                try {
                    for (const auto& sp : spans) r.corners.push_back(parseCorner(sp.first, sp.second, lineno));
                } catch (const ParseException&) {
                    // the reference validates corner k before it parses corner k+1: keep the parsed ones
                    f.count = (uint32_t)r.corners.size() - f.first;
                    f.partial = true;
                    r.faces.push_back(f);
                    throw;
                }
                r.faces.push_back(f);
            } else if (klen == 1 && tb[0] == 'v') {
                size_t n = 0;
                while (lex.next(tb, te)) {
                    const double x = parseNumber(tb, te, lineno);
                    if (n < 8) nums[n] = x;
                    n++;
                }
                if (n != 3 && n != 4) throw ParseException("v requires 3 or 4 parameters", lineno);
                Vec4 v(nums[0], nums[1], nums[2], n == 4 ? nums[3] : 1.0);
                if (v.w == 0) throw ParseException("v must be a point vector", lineno);
                r.v.push_back(v);
            } else if (klen == 2 && tb[0] == 'v' && tb[1] == 'n') {
                size_t n = 0;
                while (lex.next(tb, te)) {
                    const double x = parseNumber(tb, te, lineno);
                    if (n < 8) nums[n] = x;
                    n++;
                }
                if (n != 3) throw ParseException("vn requires 3 parameters", lineno);
                r.vn.push_back(Vec4(nums[0], nums[1], nums[2], 0.0));
            } else {
                r.warnings.push_back({lineno, "unknown obj line type " + std::string(tb, te)});
            }
        });
    } catch (const ParseException&) {
        r.error = std::current_exception();
    }
}

// vertices / normals: the whole file's arrays, 1-indexed (element 0 is a dummy).
// Returns the error of the first failing face (null if none); r.error (pass 1) is left alone.
std::exception_ptr buildObjFaces(ObjRange& r, const std::vector<Vec4>& vertices, const std::vector<Vec4>& normals) {
    std::vector<ObjWarning> merged;
    size_t wi = 0;          // pass-1 warnings of earlier lines come first
    std::exception_ptr error;
    size_t upper = 0;
    for (const ObjFaceRec& f : r.faces) upper += f.partial ? 0 : f.count - 2;
    r.built.reserve(upper);
    try {
        for (const ObjFaceRec& f : r.faces) {
            const size_t nv = 1 + r.v_base + f.nv_before, nn = 1 + r.vn_base + f.nn_before;
            const Corner* c = r.corners.data() + f.first;
            for (uint32_t k = 0; k < f.count; k++) {
                if (c[k].v == 0) throw ParseException("vertex index is required", f.lineno);
                if ((size_t)c[k].v >= nv) throw ParseException("vertex index out of range", f.lineno);
                if (c[k].vn != 0 && (size_t)c[k].vn >= nn) throw ParseException("normal index out of range", f.lineno);
            }
            if (f.partial) continue;
            const Corner& base = c[0];
            for (uint32_t k = 1; k + 1 < f.count; k++) {
                const Corner* tri[3] = {&base, &c[k], &c[k + 1]};
                Vec4 e1 = vertices[tri[1]->v] - vertices[base.v];
                Vec4 e2 = vertices[tri[2]->v] - vertices[base.v];
                Vec4 fn = Vec4::dir(cross(e1.head(), e2.head()));
                if (fn.isZero()) {
                    while (wi < r.warnings.size() && r.warnings[wi].lineno < f.lineno) merged.push_back(r.warnings[wi++]);
                    merged.push_back({f.lineno, "degenerate face"});
                    continue;
                }
                // in-place normalize(): Eigen 3.2 multiplies by the reciprocal of the norm
                fn = (1.0 / norm4(fn)) * fn;
                Mesh::Face face;
                for (int i = 0; i < 3; i++) {
                    face.points_[i] = vertices[tri[i]->v];
                    face.normals_[i] = tri[i]->vn ? normals[tri[i]->vn] : fn;
                }
                r.built.push_back(face);
            }
        }
    } catch (const ParseException&) {
        error = std::current_exception();
    }
    while (wi < r.warnings.size()) merged.push_back(r.warnings[wi++]);
    r.warnings.swap(merged);
    return error;
}

int lineOfError(const std::exception_ptr& e) {
    try {
        std::rethrow_exception(e);
    } catch (const ParseException& pe) {
        return pe.line();
    }
}

int objParserThreads(size_t bytes) {
    const char* env = std::getenv("AS2_PARSE_THREADS");
    int n = env ? std::atoi(env) : (int)std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    // at least 1 MiB of text per thread (AS2_PARSE_CHUNK_BYTES: tests chunk tiny files with it)
    const char* cenv = std::getenv("AS2_PARSE_CHUNK_BYTES");
    const size_t chunk = cenv && std::atoll(cenv) > 0 ? (size_t)std::atoll(cenv) : ((size_t)1 << 20);
    const size_t by_size = bytes / chunk;
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)n, by_size));
}
}  // namespace

// The reference reads line by line, so an error on line N is raised after the warnings of
// lines < N were printed and before anything later is looked at.  The chunked parse defers all
// output; when any chunk fails, the file is simply parsed again serially, which reproduces
// that order exactly (error paths need not be fast).
void OBJParser::parseFile(std::string filename) {
    std::vector<char> data = slurpFile(filename);
    const char* p0 = data.data();
    const char* pend = p0 + data.size();
    const int threads = objParserThreads(data.size());
    if (threads > 1) {
        std::vector<ObjRange> ranges((size_t)threads);
        for (int t = 0; t < threads; t++) {
            const char* cut = p0 + data.size() * (size_t)t / (size_t)threads;
            if (t > 0) {      // move to the start of the next line
                const char* nl = (const char*)std::memchr(cut, '\n', (size_t)(pend - cut));
                cut = nl ? nl + 1 : pend;
            }
            ranges[(size_t)t].begin = cut;
            if (t > 0) ranges[(size_t)t - 1].end = cut;
        }
        ranges.back().end = pend;
        auto onAll = [&](auto fn) {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; t++) pool.emplace_back([&, t] { fn(ranges[(size_t)t]); });
            for (std::thread& th : pool) th.join();
        };
        std::vector<int> newlines((size_t)threads, 0);
        {
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; t++)
                pool.emplace_back([&, t] {
                    int n = 0;
                    for (const char* q = ranges[(size_t)t].begin; q < ranges[(size_t)t].end;) {
                        const char* nl = (const char*)std::memchr(q, '\n', (size_t)(ranges[(size_t)t].end - q));
                        if (!nl) break;
                        n++;
                        q = nl + 1;
                    }
                    newlines[(size_t)t] = n;
                });
            for (std::thread& th : pool) th.join();
        }
        int lineno = 1;
        for (int t = 0; t < threads; t++) {
            ranges[(size_t)t].first_lineno = lineno;
            lineno += newlines[(size_t)t];
        }
        onAll([](ObjRange& r) { parseObjRange(r); });
        bool failed = false;
        size_t nv = 0, nn = 0;
        for (ObjRange& r : ranges) {
            failed = failed || (bool)r.error;
            r.v_base = nv;
            r.vn_base = nn;
            nv += r.v.size();
            nn += r.vn.size();
        }
        if (!failed) {
            std::vector<Vec4> vertices(1 + nv), normals(1 + nn);   // 1-indexed
            onAll([&](ObjRange& r) {
                std::copy(r.v.begin(), r.v.end(), vertices.begin() + 1 + (std::ptrdiff_t)r.v_base);
                std::copy(r.vn.begin(), r.vn.end(), normals.begin() + 1 + (std::ptrdiff_t)r.vn_base);
            });
            onAll([&](ObjRange& r) { r.error = buildObjFaces(r, vertices, normals); });
            for (const ObjRange& r : ranges) failed = failed || (bool)r.error;
            if (!failed) {
                size_t total = 0;
                std::vector<size_t> at((size_t)threads);
                for (int t = 0; t < threads; t++) {
                    at[(size_t)t] = mesh_.faces_.size() + total;
                    total += ranges[(size_t)t].built.size();
                }
                mesh_.faces_.resize(mesh_.faces_.size() + total);
                {
                    std::vector<std::thread> pool;
                    for (int t = 0; t < threads; t++)
                        pool.emplace_back([&, t] {
                            std::copy(ranges[(size_t)t].built.begin(), ranges[(size_t)t].built.end(),
                                      mesh_.faces_.begin() + (std::ptrdiff_t)at[(size_t)t]);
                        });
                    for (std::thread& th : pool) th.join();
                }
                for (const ObjRange& r : ranges)
                    for (const ObjWarning& w : r.warnings) ParseException::showWarning(w.msg, w.lineno);
                return;
            }
        }
    }
    // Serial: the same two passes over one range.  Pass 1 stops at its first error (line E1);
    // pass 2 then sees exactly the faces of the lines up to E1 and may fail earlier (E2 <= E1),
    // which is the error the reference's line-by-line loop would have hit first.  Warnings of
    // the lines before the error are printed, as the reference would have by then.
    ObjRange all;
    all.begin = p0;
    all.end = pend;
    parseObjRange(all);
    std::vector<Vec4> vertices(1), normals(1);   // 1-indexed
    vertices.insert(vertices.end(), all.v.begin(), all.v.end());
    normals.insert(normals.end(), all.vn.begin(), all.vn.end());
    std::exception_ptr error = all.error;
    const std::exception_ptr error2 = buildObjFaces(all, vertices, normals);
    if (error2 && (!error || lineOfError(error2) <= lineOfError(error))) error = error2;
    const int error_line = error ? lineOfError(error) : INT32_MAX;
    for (const ObjWarning& w : all.warnings)
        if (w.lineno < error_line) ParseException::showWarning(w.msg, w.lineno);
    if (error) std::rethrow_exception(error);
    mesh_.faces_.insert(mesh_.faces_.end(), all.built.begin(), all.built.end());
}

}  // namespace as2
