// parsers.cpp — see parsers.h.  Grammar reference: src/parsers.cpp:5-19 (statement
// table), :24-91 (tokens / numbers), :93-251 (.rti), :253-374 (.obj).
#include "parsers.h"

#include <libgen.h>

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

void OBJParser::parseFile(std::string filename) {
    std::vector<char> data = slurpFile(filename);
    std::vector<Vec4> vertices(1), normals(1);   // 1-indexed
    std::vector<Corner> corners;
    std::vector<std::pair<const char*, const char*>> spans;
    std::vector<double> nums;
    forEachLine(data, [&](const char* lb, const char* le, int lineno) {
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
            corners.clear();
            for (const auto& sp : spans) {
                Corner c = parseCorner(sp.first, sp.second, lineno);
                if (c.v == 0) throw ParseException("vertex index is required", lineno);
                if ((size_t)c.v >= vertices.size()) throw ParseException("vertex index out of range", lineno);
                if (c.vn != 0 && (size_t)c.vn >= normals.size())
                    throw ParseException("normal index out of range", lineno);
                corners.push_back(c);
            }
            const Corner& base = corners[0];
            for (size_t k = 1; k + 1 < corners.size(); k++) {
                const Corner* tri[3] = {&base, &corners[k], &corners[k + 1]};
                Vec4 e1 = vertices[tri[1]->v] - vertices[base.v];
                Vec4 e2 = vertices[tri[2]->v] - vertices[base.v];
                Vec4 fn = Vec4::dir(cross(e1.head(), e2.head()));
                if (fn.isZero()) {
                    ParseException::showWarning("degenerate face", lineno);
                    continue;
                }
                // in-place normalize(): Eigen 3.2 multiplies by the reciprocal of the norm
                fn = (1.0 / norm4(fn)) * fn;
                Mesh::Face face;
                for (int i = 0; i < 3; i++) {
                    face.points_[i] = vertices[tri[i]->v];
                    face.normals_[i] = tri[i]->vn ? normals[tri[i]->vn] : fn;
                }
                mesh_.faces_.push_back(face);
            }
        } else if (klen == 1 && tb[0] == 'v') {
            nums.clear();
            while (lex.next(tb, te)) nums.push_back(parseNumber(tb, te, lineno));
            if (nums.size() != 3 && nums.size() != 4) throw ParseException("v requires 3 or 4 parameters", lineno);
            Vec4 v(nums[0], nums[1], nums[2], nums.size() == 4 ? nums[3] : 1.0);
            if (v.w == 0) throw ParseException("v must be a point vector", lineno);
            vertices.push_back(v);
        } else if (klen == 2 && tb[0] == 'v' && tb[1] == 'n') {
            nums.clear();
            while (lex.next(tb, te)) nums.push_back(parseNumber(tb, te, lineno));
            if (nums.size() != 3) throw ParseException("vn requires 3 parameters", lineno);
            normals.push_back(Vec4(nums[0], nums[1], nums[2], 0.0));
        } else {
            ParseException::showWarning("unknown obj line type " + std::string(tb, te), lineno);
        }
    });
}

}  // namespace as2
