// rt_device.cuh — device-side data layout and the exact FP64 arithmetic of the trace loop.
//
// Everything that DECIDES a hit (object-space transforms, sphere quadratic, Cramer's rule,
// shading-normal cull, closest-distance compare) is evaluated in FP64 with exactly the
// association order of the reference build (g++ -O2, SSE2 Eigen 3.2.2, no FMA), so hit
// ids and ray counts match the reference bit for bit; this translation unit MUST be
// compiled with -fmad=false.  Only the LBVH boxes (culling, conservative) are FP32.
//
// Reference map:
//   xf_point/xf_dir      Transform4d * Vector4d           eigen/.../Transform.h:1244-1267
//   dot4/norm4           Vector4d dot / norm (SSE2 redux)  eigen/.../Redux.h:131-136,299-305
//   det3                 Matrix3d::determinant            eigen/.../LU/Determinant.h:18-23,61-69
//   sphere_object        Sphere::calculateIntNormInObjSpace  src/geometry.cpp:47-67
//   face_object          Mesh::calculateIntNormInObjSpace    src/geometry.cpp:78-123 (loop body)
//   mesh_bbox            hitsBoundingBox                     src/geometry.cpp:5-29
//   to_world             Geometry::calculateIntersectionNormal src/geometry.cpp:39-43
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rt_b200.h"

namespace rt {

struct d3 {
    double x, y, z;
};

__device__ __forceinline__ d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator-(d3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ d3 operator*(double s, d3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ d3 vdiv(d3 a, double s) { return mk3(a.x / s, a.y / s, a.z / s); }

// two 2-wide packets added lane-wise, then horizontally: (x0+x2)+(x1+x3); w product is +0
__device__ __forceinline__ double sum4(double a, double b, double c, double d) { return (a + c) + (b + d); }
__device__ __forceinline__ double dot4(d3 a, d3 b) { return sum4(a.x * b.x, a.y * b.y, a.z * b.z, 0.0); }
__device__ __forceinline__ double norm4(d3 a) { return sqrt(dot4(a, a)); }
// Ray constructor: direction / direction.norm()  (true division; src/rtbase.h:23)
__device__ __forceinline__ d3 ray_normalize(d3 a) { return vdiv(a, norm4(a)); }
// Vector4d::normalize(): multiply by the reciprocal (Eigen 3.2 operator/=)
__device__ __forceinline__ d3 inplace_normalize(d3 a) { return (1.0 / norm4(a)) * a; }

// rows 0..2 of a row-major 3x4 affine matrix
__device__ __forceinline__ d3 xf_point(const double* __restrict__ m, d3 v) {
    return mk3(((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3] * 1.0,
               ((m[4] * v.x + m[5] * v.y) + m[6] * v.z) + m[7] * 1.0,
               ((m[8] * v.x + m[9] * v.y) + m[10] * v.z) + m[11] * 1.0);
}
__device__ __forceinline__ d3 xf_dir(const double* __restrict__ m, d3 v) {
    return mk3(((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3] * 0.0,
               ((m[4] * v.x + m[5] * v.y) + m[6] * v.z) + m[7] * 0.0,
               ((m[8] * v.x + m[9] * v.y) + m[10] * v.z) + m[11] * 0.0);
}
// inverse.matrix().transpose() * N, w := 0
__device__ __forceinline__ d3 xf_normal(const double* __restrict__ inv, d3 n) {
    return mk3(sum4(inv[0] * n.x, inv[4] * n.y, inv[8] * n.z, 0.0),
               sum4(inv[1] * n.x, inv[5] * n.y, inv[9] * n.z, 0.0),
               sum4(inv[2] * n.x, inv[6] * n.y, inv[10] * n.z, 0.0));
}

__device__ __forceinline__ double det3(d3 c0, d3 c1, d3 c2) {
    double h012 = c0.x * (c1.y * c2.z - c2.y * c1.z);
    double h102 = c1.x * (c0.y * c2.z - c2.y * c0.z);
    double h201 = c2.x * (c0.y * c1.z - c1.y * c0.z);
    return h012 - h102 + h201;
}

// ---- device scene ------------------------------------------------------------
struct alignas(16) DGeom {
    double inv[12];
    double fwd[12];
    double det;
    double center[3];
    double radius2;
    double bbmin[3];
    double bbmax[3];
    int type;
    int mat;
    int first_face;
    int num_faces;
    int use_bbox;
    int pad_[3];
};

struct alignas(16) DMat {
    double ka[3], kd[3], ks[3], kr[3];
    double sp, ior;
    int has_kt;        // translucencyColor_ != 0  (src/scene.cpp:115)
    int has_kr;        // reflectiveColor_ != 0    (src/scene.cpp:130)
    int pad_[2];
};

struct alignas(16) DLight {
    double v[3];
    double color[3];
    double falloff;
    int type;
    int pad_;
};

struct DCamera {
    double eye[3], ll[3], lr[3], ul[3], ur[3];
};

// One face = 5 x double2 = 80 B: p0.xy | p0.z va.x | va.yz | vb.xy | vb.z aux
// where va = p1 - p0, vb = p2 - p0 (the subtractions src/geometry.cpp:80-81 does per ray)
// and aux packs (geometry index, face index within the geometry) as two int32.
// Normals: n0.xy | n0.z n1.x | n1.yz | n2.xy | n2.z pad.
#define RT_FACE_D2 5

// 4-wide LBVH node, 128 B = 8 x float4: the four child boxes as SoA (FP32, padded outward)
// + four child references.  ref >= 0: wide-node index; ref < 0: leaf, prim code = ~ref;
// empty slot: ref = BVH_DONE with a NaN box (never hit).  A wide node is a binary
// Karras node at even depth with its grandchildren pulled up, which halves the number of
// dependent fetches per ray (the traversal kernels are latency-bound).
struct alignas(16) BvhNode {
    float4 lox, loy, loz;
    float4 hix, hiy, hiz;
    int4 ref;
    int4 pad_;
};
#define BVH_DONE ((int)0x80000000)     // not a valid ref (prim codes stay below 3<<29)

// prim code: kind in the top 2 bits of a 31-bit value
#define PRIM_KIND_SHIFT 29
#define PRIM_FACE 0      // index = global face index (mesh face)
#define PRIM_SPHERE 1    // index = geometry index
#define PRIM_TRI 2       // index = geometry index (tests faces first_face, first_face+1)
#define PRIM_INDEX_MASK ((1 << PRIM_KIND_SHIFT) - 1)

struct DScene {
    DCamera cam;
    const DGeom* geoms;
    const DMat* mats;
    const DLight* slights;     // non-ambient lights, insertion order (each casts a shadow ray)
    const DLight* alights;     // ambient lights, insertion order
    const double2* face_pts;
    const double2* face_nrm;
    const BvhNode* nodes;      // null when fewer than two primitives are in the BVH
    const int* flat;           // prim codes tested by every ray, in geometry order
    const int* all_prims;      // every prim code in geometry/face order (brute force)
    const float4* sph_bound;   // per geometry: world-space bounding sphere (xyz, radius) of a SPHERE, FP32
    int num_geoms;
    int num_slights;
    int num_alights;
    int num_flat;
    int num_all;
    int num_bvh_prims;
    int single_leaf;           // BVH with exactly one primitive: its prim code
    int shadow_mode;           // bit 0: light-major thread mapping of k_shadow
};

// ---- hit bookkeeping ------------------------------------------------------------
struct Best {
    int geom;        // -1: none
    int face;
    double dobj;     // object-space distance t*|d'| (mesh) — compared within a geometry
    double wd;       // world distance |P - o| (src/scene.cpp:153) — compared across geometries
    d3 P, N;         // world-space point, un-normalised world normal (after det flip)
};

// object-space ray of the geometry currently being tested (cached across BVH leaves)
struct ObjRay {
    int geom;        // -1: nothing cached
    int box_ok;      // hitsBoundingBox verdict for MESH geometries with use_bbox
    d3 o, d, nd;     // origin, unit direction, -direction
    double dn;       // Vector3d norm of d: sqrt(dx^2 + (dy^2 + dz^2))
};

struct WorkCounters {
    unsigned long long nodes, tris, spheres;
};

__device__ __forceinline__ bool mesh_bbox(d3 o, d3 d, const double* __restrict__ bbmin,
                                          const double* __restrict__ bbmax) {
    const double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
#pragma unroll
    for (int axis = 0; axis < 3; axis++) {
#pragma unroll
        for (int bn = 0; bn < 2; bn++) {
            double mag = dd[axis];
            if (mag == 0.0) continue;
            double t = ((bn ? bbmax[axis] : bbmin[axis]) - oo[axis]) / mag;
            if (t < 0) continue;
            bool ok = true;
#pragma unroll
            for (int a2 = 0; a2 < 3; a2++) {
                if (a2 == axis) continue;
                double ip = oo[a2] + t * dd[a2];
                if (ip < bbmin[a2] || ip > bbmax[a2]) ok = false;
            }
            if (ok) return true;
        }
    }
    return false;
}

__device__ __forceinline__ void load_objray(const DScene& S, int gi, d3 o, d3 d, ObjRay& R) {
    const DGeom* g = S.geoms + gi;
    R.geom = gi;
    R.o = xf_point(g->inv, o);
    R.d = ray_normalize(xf_dir(g->inv, d));
    R.nd = -R.d;
    R.dn = sqrt(R.d.x * R.d.x + (R.d.y * R.d.y + R.d.z * R.d.z));
    R.box_ok = 1;
    if (g->type == RT_GEOM_MESH && g->use_bbox) R.box_ok = mesh_bbox(R.o, R.d, g->bbmin, g->bbmax) ? 1 : 0;
}

// Offer a candidate (geometry gi, face f) to the running closest hit.  Within one
// geometry the reference keeps the accepted face of smallest object-space distance, first
// face on ties (src/geometry.cpp:108-110); across geometries the smallest world distance,
// first geometry on ties (src/scene.cpp:153-155).
__device__ __forceinline__ bool beats(const Best& best, int gi, int f, double dobj, double wd) {
    if (best.geom < 0) return true;
    if (gi == best.geom) return dobj < best.dobj || (dobj == best.dobj && f < best.face);
    return wd < best.wd || (wd == best.wd && gi < best.geom);
}

// World-space completion of an object-space hit (src/geometry.cpp:39-43) + world distance.
__device__ __forceinline__ void to_world(const DGeom* g, d3 Pobj, d3 Nobj, d3 o, d3& P, d3& N, double& wd) {
    P = xf_point(g->fwd, Pobj);
    N = xf_normal(g->inv, Nobj);
    if (g->det < 0) N = -N;
    wd = norm4(P - o);
}

// ANYHIT = true: shadow query — report as soon as an accepted hit has wd <= limit.
template <bool ANYHIT>
__device__ __forceinline__ bool test_sphere(const DScene& S, int gi, d3 o, d3 d, bool reverse, double limit,
                                            Best& best) {
    const DGeom* g = S.geoms + gi;
    d3 oo = xf_point(g->inv, o);
    d3 dd = ray_normalize(xf_dir(g->inv, d));
    d3 c = mk3(g->center[0], g->center[1], g->center[2]);
    d3 ocd = oo - c;
    double a = dot4(dd, dd);
    double b = 2 * dot4(dd, ocd);
    double cc = dot4(ocd, ocd) - g->radius2;
    double disc = b * b - 4 * a * cc;
    if (disc < 0) return false;
    double res = reverse ? (-b + sqrt(disc)) / (2 * a) : (-b - sqrt(disc)) / (2 * a);
    if (res < 0) return false;
    d3 Pobj = oo + res * dd;
    d3 Nobj = Pobj - c;
    d3 P, N;
    double wd;
    to_world(g, Pobj, Nobj, o, P, N, wd);
    if (ANYHIT) return wd <= limit;
    if (beats(best, gi, 0, 0.0, wd)) {
        best.geom = gi; best.face = 0; best.dobj = 0.0; best.wd = wd; best.P = P; best.N = N;
    }
    return false;
}

// One face of a TRI/MESH geometry; R must hold the object-space ray of geometry gi.
template <bool ANYHIT>
__device__ __forceinline__ bool test_face(const DScene& S, int gi, int gface, int lface, const ObjRay& R, d3 o,
                                          bool reverse, double limit, Best& best) {
    const double2* __restrict__ fp = S.face_pts + (size_t)gface * RT_FACE_D2;
    double2 q0 = __ldg(fp + 0), q1 = __ldg(fp + 1), q2 = __ldg(fp + 2), q3 = __ldg(fp + 3), q4 = __ldg(fp + 4);
    d3 p0 = mk3(q0.x, q0.y, q1.x), va = mk3(q1.y, q2.x, q2.y), vb = mk3(q3.x, q3.y, q4.x);
    d3 rhs = R.o - p0;
    double dlower = det3(va, vb, R.nd);
    if (dlower == 0) return false;
    double a = det3(rhs, vb, R.nd) / dlower;
    if (a < 0 || a > 1) return false;
    double b = det3(va, rhs, R.nd) / dlower;
    if (b < 0 || a + b > 1) return false;
    double t = det3(va, vb, rhs) / dlower;
    if (t < 0) return false;
    double dobj = t * R.dn;
    if (!ANYHIT && best.geom == gi && (dobj > best.dobj || (dobj == best.dobj && lface > best.face))) return false;
    const double2* __restrict__ fn = S.face_nrm + (size_t)gface * RT_FACE_D2;
    double2 m0 = __ldg(fn + 0), m1 = __ldg(fn + 1), m2 = __ldg(fn + 2), m3 = __ldg(fn + 3), m4 = __ldg(fn + 4);
    d3 n0 = mk3(m0.x, m0.y, m1.x), n1 = mk3(m1.y, m2.x, m2.y), n2 = mk3(m3.x, m3.y, m4.x);
    double w0 = 1.0 - a - b;
    d3 tn = (w0 * n0 + a * n1) + b * n2;
    bool front = dot4(tn, R.d) < 0;
    if ((!front) != reverse) return false;      // (!hitsFront) ^ reverseNormals -> skip
    d3 Pobj = p0 + (a * va + b * vb);
    d3 P, N;
    double wd;
    to_world(S.geoms + gi, Pobj, tn, o, P, N, wd);
    if (ANYHIT) return wd <= limit;
    if (beats(best, gi, lface, dobj, wd)) {
        best.geom = gi; best.face = lface; best.dobj = dobj; best.wd = wd; best.P = P; best.N = N;
    }
    return false;
}

// FP32 pre-test against the sphere's world-space BOUNDING sphere (centre, radius >= the
// ellipsoid's largest semi-axis, both rounded outward at upload).  It may only say "missed"
// when the exact test would: the margins are ~250x the FP32 rounding of the expression
// (relative 1e-4 on |c-o|^2, 1e-3 on the radius).  Skips about half of the exact tests
// (a sphere fills 52 % of its box' projected area).
__device__ __forceinline__ bool sphere_certainly_missed(float4 bs, d3 o, d3 d) {
    float ocx = bs.x - (float)o.x, ocy = bs.y - (float)o.y, ocz = bs.z - (float)o.z;
    float dx = (float)d.x, dy = (float)d.y, dz = (float)d.z;
    float c2 = ocx * ocx + ocy * ocy + ocz * ocz;
    float b = ocx * dx + ocy * dy + ocz * dz;
    float r = bs.w * 1.001f;
    float r2 = r * r + 1e-4f * c2 + 1e-30f;
    if (c2 - b * b > r2) return true;                       // the line passes outside the sphere
    return c2 > r2 && b < -(r + 1e-3f * sqrtf(c2));        // origin outside and the sphere entirely behind it
}

// Dispatch one prim code.  Returns true only for ANYHIT occlusion.
template <bool ANYHIT, bool COUNT>
__device__ __forceinline__ bool test_prim(const DScene& S, int code, d3 o, d3 d, bool reverse, double limit,
                                          ObjRay& R, Best& best, WorkCounters& wc) {
    const int kind = code >> PRIM_KIND_SHIFT, idx = code & PRIM_INDEX_MASK;
    if (kind == PRIM_SPHERE) {
        if (sphere_certainly_missed(__ldg(S.sph_bound + idx), o, d)) return false;
        if (COUNT) wc.spheres++;
        return test_sphere<ANYHIT>(S, idx, o, d, reverse, limit, best);
    }
    if (kind == PRIM_TRI) {
        const DGeom* g = S.geoms + idx;
        if (R.geom != idx) load_objray(S, idx, o, d, R);
        if (COUNT) wc.tris += 2;
        if (test_face<ANYHIT>(S, idx, g->first_face, 0, R, o, reverse, limit, best)) return true;
        return test_face<ANYHIT>(S, idx, g->first_face + 1, 1, R, o, reverse, limit, best);
    }
    // mesh face: geometry/local face index ride in the face record's aux slot
    double2 q4 = __ldg(S.face_pts + (size_t)idx * RT_FACE_D2 + 4);
    int gi = __double2loint(q4.y), lf = __double2hiint(q4.y);
    if (R.geom != gi) load_objray(S, gi, o, d, R);
    if (!R.box_ok) return false;
    if (COUNT) wc.tris++;
    return test_face<ANYHIT>(S, gi, idx, lf, R, o, reverse, limit, best);
}

// ---- FP32 conservative slab test -------------------------------------------------
struct FRay {
    float ix, iy, iz;      // 1/d (clamped away from 0)
    float bx, by, bz;      // -o/d, so that t = plane * (1/d) + (-o/d) is one FMA
};
__device__ __forceinline__ FRay make_fray(d3 o, d3 d) {
    FRay r;
    float ox = (float)o.x, oy = (float)o.y, oz = (float)o.z;
    float dx = (float)d.x, dy = (float)d.y, dz = (float)d.z;
    const float tiny = 1e-30f;
    if (fabsf(dx) < tiny) dx = copysignf(tiny, dx);
    if (fabsf(dy) < tiny) dy = copysignf(tiny, dy);
    if (fabsf(dz) < tiny) dz = copysignf(tiny, dz);
    r.ix = 1.0f / dx; r.iy = 1.0f / dy; r.iz = 1.0f / dz;
    r.bx = -ox * r.ix; r.by = -oy * r.iy; r.bz = -oz * r.iz;
    return r;
}
// Entry distance of the (already padded) box, or +inf when missed.  Every FP32 rounding in
// here (conversion of o and d, 1/d, the FMA) moves a plane distance by the equivalent of at
// most a few 1e-7 x the largest scene coordinate in position, which the box padding
// (2e-6 x that coordinate, rt_bvh.cu) covers, and tlimit is already rounded up with slack
// (prune_limit), so rounding can only let MORE boxes through, never fewer.
// (Explicit __fmaf_rn: this file is compiled with -fmad=false.)
__device__ __forceinline__ float slab(const FRay& r, float lx, float ly, float lz, float hx, float hy, float hz,
                                      float tlimit) {
    float t0x = __fmaf_rn(lx, r.ix, r.bx), t1x = __fmaf_rn(hx, r.ix, r.bx);
    float t0y = __fmaf_rn(ly, r.iy, r.by), t1y = __fmaf_rn(hy, r.iy, r.by);
    float t0z = __fmaf_rn(lz, r.iz, r.bz), t1z = __fmaf_rn(hz, r.iz, r.bz);
    float tmin = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.f));
    float tmax = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tlimit));
    return tmin <= tmax ? tmin : __int_as_float(0x7f800000);
}

#define RT_STACK 96
#ifndef RT_ANYHIT_UNSORTED
#define RT_ANYHIT_UNSORTED 1
#endif

__device__ __forceinline__ float prune_limit(double x) {
    // FP32 prune limit, rounded up with slack
    if (!(x < 3.0e38)) return __int_as_float(0x7f800000);
    return __double2float_ru(x) * 1.00001f + 1e-30f;
}

// Traversal state of one ray over the 4-wide LBVH.  Leaves are deferred through the same
// stack as internal nodes ("while-while"): descend() only walks internal nodes (cheap FP32
// slab tests) and returns at a leaf, so the threads of a warp reconverge before the long
// exact FP64 primitive test instead of diverging into it.
template <bool ANYHIT>
struct Trav {
    int stack[RT_STACK];
    float tstack[ANYHIT ? 1 : RT_STACK];   // closest hit: entry distance of deferred subtrees
    int sp;
    int cur;

    __device__ __forceinline__ void start() { sp = 0; cur = 0; }   // node 0 is the root
    __device__ __forceinline__ int pop(float tlim) {
        while (sp) {
            --sp;
            if (ANYHIT || tstack[sp] <= tlim) return stack[sp];   // skip subtrees a closer hit made obsolete
        }
        return BVH_DONE;
    }
    __device__ __forceinline__ void push(int ref, float t) {
        if (sp < RT_STACK) {               // 3 pushes per wide level, depth <= 31 wide levels
            stack[sp] = ref;
            if (!ANYHIT) tstack[sp] = t;
            sp++;
        }
    }
    // Walk internal nodes until `cur` is a leaf reference (< 0) or BVH_DONE.
    template <bool COUNT>
    __device__ __forceinline__ void descend(const DScene& S, const FRay& fr, float tlim, WorkCounters& wc) {
        while (cur >= 0) {
            const float4* __restrict__ np = reinterpret_cast<const float4*>(S.nodes + cur);
            const float4 lox = __ldg(np + 0), loy = __ldg(np + 1), loz = __ldg(np + 2);
            const float4 hix = __ldg(np + 3), hiy = __ldg(np + 4), hiz = __ldg(np + 5);
            const int4 ref = __ldg(reinterpret_cast<const int4*>(np + 6));
            if (COUNT) wc.nodes += 4;
            const float t0 = slab(fr, lox.x, loy.x, loz.x, hix.x, hiy.x, hiz.x, tlim);
            const float t1 = slab(fr, lox.y, loy.y, loz.y, hix.y, hiy.y, hiz.y, tlim);
            // slots 2 and 3 may be empty (ref == BVH_DONE, NaN box: min/max would ignore the NaNs)
            const float INFT = __int_as_float(0x7f800000);
            const float t2 = ref.z != BVH_DONE ? slab(fr, lox.z, loy.z, loz.z, hix.z, hiy.z, hiz.z, tlim) : INFT;
            const float t3 = ref.w != BVH_DONE ? slab(fr, lox.w, loy.w, loz.w, hix.w, hiy.w, hiz.w, tlim) : INFT;
            if (ANYHIT && RT_ANYHIT_UNSORTED) {
                // occlusion query: any order will do; take the first hit slot, defer the others
                const float INF = __int_as_float(0x7f800000);
                int next = BVH_DONE;
                if (t3 < INF) next = ref.w;
                if (t2 < INF) { if (next != BVH_DONE) push(next, 0.f); next = ref.z; }
                if (t1 < INF) { if (next != BVH_DONE) push(next, 0.f); next = ref.y; }
                if (t0 < INF) { if (next != BVH_DONE) push(next, 0.f); next = ref.x; }
                cur = next != BVH_DONE ? next : pop(tlim);
                continue;
            }
            // sort the four (distance, slot) pairs: slot index rides in the two low mantissa
            // bits (t >= 0, so the float bit patterns order like unsigned ints; +inf = miss)
            unsigned k0 = (__float_as_uint(t0) & ~3u) | 0u, k1 = (__float_as_uint(t1) & ~3u) | 1u;
            unsigned k2 = (__float_as_uint(t2) & ~3u) | 2u, k3 = (__float_as_uint(t3) & ~3u) | 3u;
            unsigned a, b;
            a = min(k0, k1); b = max(k0, k1); k0 = a; k1 = b;
            a = min(k2, k3); b = max(k2, k3); k2 = a; k3 = b;
            a = min(k0, k2); b = max(k0, k2); k0 = a; k2 = b;
            a = min(k1, k3); b = max(k1, k3); k1 = a; k3 = b;
            a = min(k1, k2); b = max(k1, k2); k1 = a; k2 = b;
            const unsigned MISS = 0x7f800000u;
            auto ref_of = [&](unsigned k) -> int {
                unsigned i = k & 3u;
                return i == 0 ? ref.x : (i == 1 ? ref.y : (i == 2 ? ref.z : ref.w));
            };
            if (k0 < MISS) {
                if (k1 < MISS) {
                    if (k2 < MISS) {
                        if (k3 < MISS) push(ref_of(k3), __uint_as_float(k3 & ~3u));
                        push(ref_of(k2), __uint_as_float(k2 & ~3u));
                    }
                    push(ref_of(k1), __uint_as_float(k1 & ~3u));
                }
                cur = ref_of(k0);
            } else {
                cur = pop(tlim);
            }
        }
    }
};

// The closest-hit / any-hit query == Scene::castRay (src/scene.cpp:142-167).
//   ANYHIT: returns true when an accepted hit with world distance <= limit exists.
//   BRUTE : ignore the LBVH and test every primitive (debug / parity aid).
template <bool ANYHIT, bool BRUTE, bool COUNT>
__device__ __forceinline__ bool cast_ray(const DScene& S, d3 o, d3 d, bool reverse, double limit, Best& best,
                                         WorkCounters& wc) {
    best.geom = -1; best.face = -1; best.dobj = 0.0; best.wd = 0.0;
    ObjRay R;
    R.geom = -1; R.box_ok = 1;
    if (BRUTE) {
        for (int i = 0; i < S.num_all; i++)
            if (test_prim<ANYHIT, COUNT>(S, __ldg(S.all_prims + i), o, d, reverse, limit, R, best, wc)) return true;
        return false;
    }
    for (int i = 0; i < S.num_flat; i++)
        if (test_prim<ANYHIT, COUNT>(S, __ldg(S.flat + i), o, d, reverse, limit, R, best, wc)) return true;
    if (S.num_bvh_prims == 0) return false;
    if (S.num_bvh_prims == 1)
        return test_prim<ANYHIT, COUNT>(S, S.single_leaf, o, d, reverse, limit, R, best, wc);

    const FRay fr = make_fray(o, d);
    float tlim = ANYHIT ? prune_limit(limit) : __int_as_float(0x7f800000);   // shrinks as closer hits are found
    Trav<ANYHIT> T;
    T.start();
    while (true) {
        T.template descend<COUNT>(S, fr, tlim, wc);
        if (T.cur == BVH_DONE) break;
        if (test_prim<ANYHIT, COUNT>(S, ~T.cur, o, d, reverse, limit, R, best, wc)) return true;
        if (!ANYHIT && best.geom >= 0) tlim = prune_limit(best.wd);
        T.cur = T.pop(tlim);
        if (T.cur == BVH_DONE) break;
    }
    return false;
}

}  // namespace rt
