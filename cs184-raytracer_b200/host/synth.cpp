// synth.cpp — see synth.h.
#include "synth.h"

#include <cmath>
#include <cstdio>
#include <memory>
#include <vector>

namespace as2 {
namespace {

struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }   // [0,1)
    double range(double lo, double hi) { return lo + (hi - lo) * uniform(); }
};

struct MatSpec {
    double v[17];   // ka(3) kd(3) ks(3) sp kr(3) kt(3) ior — the `mat` statement's order
};
struct SphereSpec {
    double c[3];
    double r;
    MatSpec mat;
};
struct LightSpec {
    int directional;
    double v[3], color[3];
};
struct Description {
    int cells = 0;
    std::vector<double> vx, vy, vz, nx, ny, nz;   // (cells+1)^2 vertices, row-major in z then x
    MatSpec terrain;
    std::vector<SphereSpec> spheres;
    std::vector<LightSpec> lights;
    double ambient[3];
    double cam[15];
};

struct Terrain {
    double amp[4], freq[4], phx[4], phz[4];
    double height(double x, double z) const {
        double y = 0;
        for (int k = 0; k < 4; k++) y += amp[k] * std::sin(freq[k] * x + phx[k]) * std::sin(freq[k] * z + phz[k]);
        return y;
    }
    void gradient(double x, double z, double& dx, double& dz) const {
        dx = dz = 0;
        for (int k = 0; k < 4; k++) {
            dx += amp[k] * freq[k] * std::cos(freq[k] * x + phx[k]) * std::sin(freq[k] * z + phz[k]);
            dz += amp[k] * freq[k] * std::sin(freq[k] * x + phx[k]) * std::cos(freq[k] * z + phz[k]);
        }
    }
};

MatSpec makeMat(double ka, const double kd[3], double ks, double sp, double kr, double kt, double ior) {
    MatSpec m;
    for (int k = 0; k < 3; k++) {
        m.v[k] = ka * kd[k];
        m.v[3 + k] = kd[k];
        m.v[6 + k] = ks;
        m.v[10 + k] = kr;
        m.v[13 + k] = kt;
    }
    m.v[9] = sp;
    m.v[16] = ior;
    return m;
}

Description describe(int cells, int num_spheres, uint64_t seed) {
    if (cells < 1 || num_spheres < 0) throw ParseException("synthetic scene: bad size");
    Description d;
    d.cells = cells;
    SplitMix64 rng(seed);
    Terrain t;
    for (int k = 0; k < 4; k++) {
        t.amp[k] = 8.0 / (double)(1 << k);
        t.freq[k] = 0.05 * (double)(1 << k);
        t.phx[k] = rng.range(0.0, 2 * M_PI);
        t.phz[k] = rng.range(0.0, 2 * M_PI);
    }
    const int nv = cells + 1;
    d.vx.resize((size_t)nv * nv);
    d.vy = d.vz = d.nx = d.ny = d.nz = d.vx;
    for (int iz = 0; iz < nv; iz++)
        for (int ix = 0; ix < nv; ix++) {
            size_t i = (size_t)iz * nv + ix;
            double x = -100.0 + 200.0 * ix / cells, z = -100.0 + 200.0 * iz / cells;
            double gx, gz;
            t.gradient(x, z, gx, gz);
            double len = std::sqrt(gx * gx + 1.0 + gz * gz);
            d.vx[i] = x;
            d.vy[i] = t.height(x, z);
            d.vz[i] = z;
            d.nx[i] = -gx / len;
            d.ny[i] = 1.0 / len;
            d.nz[i] = -gz / len;
        }
    const double terrainKd[3] = {0.45, 0.55, 0.35};
    d.terrain = makeMat(0.1, terrainKd, 0.2, 20.0, 0.15, 0.0, 0.0);
    for (int s = 0; s < num_spheres; s++) {
        SphereSpec sp;
        sp.c[0] = rng.range(-90.0, 90.0);
        sp.c[2] = rng.range(-90.0, 90.0);
        sp.c[1] = t.height(sp.c[0], sp.c[2]) + rng.range(1.0, 40.0);
        sp.r = rng.range(0.5, 3.0);
        double kind = rng.uniform();
        double kd[3] = {rng.range(0.2, 1.0), rng.range(0.2, 1.0), rng.range(0.2, 1.0)};
        if (kind < 0.7)
            sp.mat = makeMat(0.1, kd, 0.6, 30.0, 0.0, 0.0, 0.0);
        else if (kind < 0.9)
            sp.mat = makeMat(0.05, kd, 0.8, 60.0, 0.5, 0.0, 0.0);
        else {
            const double glassKd[3] = {0.05, 0.05, 0.05};
            sp.mat = makeMat(0.0, glassKd, 1.0, 120.0, 0.1, 0.9, 1.5);
        }
        d.spheres.push_back(sp);
    }
    for (int k = 0; k < 6; k++) {
        LightSpec l;
        l.directional = 0;
        double a = 2 * M_PI * k / 6.0;
        l.v[0] = 150.0 * std::cos(a);
        l.v[1] = 120.0;
        l.v[2] = 150.0 * std::sin(a);
        l.color[0] = 0.22; l.color[1] = 0.2; l.color[2] = 0.18;
        d.lights.push_back(l);
    }
    const double dirs[2][3] = {{-0.3, -1.0, -0.2}, {0.4, -1.0, 0.3}};
    for (int k = 0; k < 2; k++) {
        LightSpec l;
        l.directional = 1;
        for (int j = 0; j < 3; j++) { l.v[j] = dirs[k][j]; l.color[j] = 0.15; }
        d.lights.push_back(l);
    }
    d.ambient[0] = d.ambient[1] = d.ambient[2] = 0.5;
    // 16:9 camera looking down at the terrain centre
    const double eye[3] = {0.0, 95.0, 185.0}, target[3] = {0.0, 0.0, 10.0};
    double f[3] = {target[0] - eye[0], target[1] - eye[1], target[2] - eye[2]};
    double fl = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (double& c : f) c /= fl;
    double r[3] = {-f[2], 0.0, f[0]};   // f x up(0,1,0)
    double rl = std::sqrt(r[0] * r[0] + r[2] * r[2]);
    for (double& c : r) c /= rl;
    double u[3] = {r[1] * f[2] - r[2] * f[1], r[2] * f[0] - r[0] * f[2], r[0] * f[1] - r[1] * f[0]};
    const double hw = 0.62, hh = hw * 9.0 / 16.0;
    for (int k = 0; k < 3; k++) {
        double c = eye[k] + f[k];
        d.cam[k] = eye[k];
        d.cam[3 + k] = c - hw * r[k] - hh * u[k];    // LL
        d.cam[6 + k] = c + hw * r[k] - hh * u[k];    // LR
        d.cam[9 + k] = c - hw * r[k] + hh * u[k];    // UL
        d.cam[12 + k] = c + hw * r[k] + hh * u[k];   // UR
    }
    return d;
}

Material toMaterial(const MatSpec& m) {
    Material mat;
    for (int k = 0; k < 3; k++) {
        mat.ambientColor_[k] = m.v[k];
        mat.diffuseColor_[k] = m.v[3 + k];
        mat.specularColor_[k] = m.v[6 + k];
        mat.reflectiveColor_[k] = m.v[10 + k];
        mat.translucencyColor_[k] = m.v[13 + k];
    }
    mat.specularCoefficient_ = m.v[9];
    mat.indexOfRefractivity_ = m.v[16];
    return mat;
}

void printMat(FILE* f, const MatSpec& m) {
    std::fprintf(f, "mat");
    for (int k = 0; k < 17; k++) std::fprintf(f, " %.17g", m.v[k]);
    std::fprintf(f, "\n");
}

}  // namespace

void buildSyntheticScene(Scene& scene, int grid_cells, int num_spheres, uint64_t seed) {
    Description d = describe(grid_cells, num_spheres, seed);
    const Affine I = Affine::identity();
    Camera cam;
    cam.forwardTransform(I);
    auto P = [&](const double* p) { return Vec4(p[0], p[1], p[2], 1.0); };
    cam.eyePoint(P(d.cam));
    cam.lowerLeftPoint(P(d.cam + 3));
    cam.lowerRightPoint(P(d.cam + 6));
    cam.upperLeftPoint(P(d.cam + 9));
    cam.upperRightPoint(P(d.cam + 12));
    scene.camera(cam);
    for (const LightSpec& l : d.lights) {
        if (l.directional) {
            std::unique_ptr<DirectionalLight> dl(new DirectionalLight());
            dl->forwardTransform(I);
            dl->direction(Vec4::dir(normalized3(Vec3(l.v[0], l.v[1], l.v[2]))));
            dl->color_ = {{l.color[0], l.color[1], l.color[2]}};
            scene.addLight(std::move(dl));
        } else {
            std::unique_ptr<PointLight> pl(new PointLight());
            pl->forwardTransform(I);
            pl->point(P(l.v));
            pl->color_ = {{l.color[0], l.color[1], l.color[2]}};
            pl->falloffExponent_ = 0.0;
            scene.addLight(std::move(pl));
        }
    }
    {
        std::unique_ptr<AmbientLight> al(new AmbientLight());
        al->forwardTransform(I);
        al->color_ = {{d.ambient[0], d.ambient[1], d.ambient[2]}};
        scene.addLight(std::move(al));
    }
    {
        std::unique_ptr<Mesh> mesh(new Mesh());
        mesh->forwardTransform(I);
        mesh->material_ = toMaterial(d.terrain);
        const int nv = d.cells + 1;
        mesh->faces_.reserve((size_t)d.cells * d.cells * 2);
        auto vert = [&](size_t i) { return Vec4(d.vx[i], d.vy[i], d.vz[i], 1.0); };
        auto nrm = [&](size_t i) { return Vec4(d.nx[i], d.ny[i], d.nz[i], 0.0); };
        for (int iz = 0; iz < d.cells; iz++)
            for (int ix = 0; ix < d.cells; ix++) {
                size_t i00 = (size_t)iz * nv + ix, i10 = i00 + 1, i01 = i00 + nv, i11 = i01 + 1;
                const size_t tris[2][3] = {{i00, i11, i10}, {i00, i01, i11}};
                for (auto& t : tris) {
                    Mesh::Face face;
                    for (int k = 0; k < 3; k++) {
                        face.points_[k] = vert(t[k]);
                        face.normals_[k] = nrm(t[k]);
                    }
                    mesh->faces_.push_back(face);
                }
            }
        mesh->updateBoundingBox();
        scene.addGeometry(std::move(mesh));
    }
    for (const SphereSpec& s : d.spheres) {
        std::unique_ptr<Sphere> sp(new Sphere());
        sp->forwardTransform(I);
        sp->material_ = toMaterial(s.mat);
        sp->center_ = P(s.c);
        sp->radius_ = (float)s.r;
        scene.addGeometry(std::move(sp));
    }
}

void writeSyntheticScene(const std::string& rti_path, const std::string& obj_path, int grid_cells,
                         int num_spheres, uint64_t seed) {
    Description d = describe(grid_cells, num_spheres, seed);
    FILE* o = std::fopen(obj_path.c_str(), "w");
    if (!o) throw WriteException("cannot open " + obj_path);
    const int nv = d.cells + 1;
    for (size_t i = 0; i < d.vx.size(); i++) std::fprintf(o, "v %.17g %.17g %.17g\n", d.vx[i], d.vy[i], d.vz[i]);
    for (size_t i = 0; i < d.vx.size(); i++) std::fprintf(o, "vn %.17g %.17g %.17g\n", d.nx[i], d.ny[i], d.nz[i]);
    for (int iz = 0; iz < d.cells; iz++)
        for (int ix = 0; ix < d.cells; ix++) {
            size_t i00 = (size_t)iz * nv + ix + 1, i10 = i00 + 1, i01 = i00 + nv, i11 = i01 + 1;   // 1-based
            std::fprintf(o, "f %zu//%zu %zu//%zu %zu//%zu\n", i00, i00, i11, i11, i10, i10);
            std::fprintf(o, "f %zu//%zu %zu//%zu %zu//%zu\n", i00, i00, i01, i01, i11, i11);
        }
    std::fclose(o);
    FILE* f = std::fopen(rti_path.c_str(), "w");
    if (!f) throw WriteException("cannot open " + rti_path);
    std::fprintf(f, "# synthetic scene: %d x %d cells, %d spheres, seed %llu\n", d.cells, d.cells, num_spheres,
                 (unsigned long long)seed);
    std::fprintf(f, "cam");
    for (int k = 0; k < 15; k++) std::fprintf(f, " %.17g", d.cam[k]);
    std::fprintf(f, "\n");
    for (const LightSpec& l : d.lights)
        std::fprintf(f, "%s %.17g %.17g %.17g %.17g %.17g %.17g\n", l.directional ? "ltd" : "ltp", l.v[0], l.v[1],
                     l.v[2], l.color[0], l.color[1], l.color[2]);
    std::fprintf(f, "lta %.17g %.17g %.17g\n", d.ambient[0], d.ambient[1], d.ambient[2]);
    printMat(f, d.terrain);
    std::string objname = obj_path;
    size_t slash = objname.find_last_of('/');
    if (slash != std::string::npos) objname = objname.substr(slash + 1);
    std::fprintf(f, "obj \"%s\"\n", objname.c_str());
    for (const SphereSpec& s : d.spheres) {
        printMat(f, s.mat);
        std::fprintf(f, "sph %.17g %.17g %.17g %.17g\n", s.c[0], s.c[1], s.c[2], s.r);
    }
    std::fclose(f);
}

}  // namespace as2
