// rt_bvh.cu — GPU LBVH build: primitive bounds -> 30-bit Morton codes -> hand-written
// LSD radix sort (no CUB) -> Karras hierarchy (pure integer, deterministic) -> bottom-up
// refit -> surface-area-guided collapse into 128-byte 4-wide nodes.  It replaces the reference's only acceleration structure,
// one object-space AABB per .obj mesh (Mesh::updateBoundingBox src/geometry.cpp:145-162,
// hitsBoundingBox src/geometry.cpp:5-29).  The boxes only CULL (FP32, padded outward);
// every hit decision is still the exact FP64 test in rt_device.cuh.
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "rt_bvh.h"
#include "rt_sort.cuh"

namespace rt {

// ---- float <-> ordered int so atomicMin/atomicMax work on floats -----------------
__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__host__ __device__ __forceinline__ float ord2f(int i) {
    int j = i >= 0 ? i : i ^ 0x7fffffff;
#ifdef __CUDA_ARCH__
    return __int_as_float(j);
#else
    float f;
    memcpy(&f, &j, 4);
    return f;
#endif
}

// World-space bounds of one primitive (FP64 math, rounded outward to FP32).
__global__ void k_prim_bounds(DScene S, const int* __restrict__ codes, int n, float* __restrict__ lo,
                              float* __restrict__ hi, int* __restrict__ gbounds /*[7]: cmin xyz, cmax xyz, maxabs*/) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int code = codes[i];
    int kind = code >> PRIM_KIND_SHIFT, idx = code & PRIM_INDEX_MASK;
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    auto grow = [&](d3 p) {
        mn[0] = fmin(mn[0], p.x); mn[1] = fmin(mn[1], p.y); mn[2] = fmin(mn[2], p.z);
        mx[0] = fmax(mx[0], p.x); mx[1] = fmax(mx[1], p.y); mx[2] = fmax(mx[2], p.z);
    };
    auto face_bounds = [&](int gface, const DGeom* g) {
        const double2* fp = S.face_pts + (size_t)gface * RT_FACE_D2;
        double2 q0 = fp[0], q1 = fp[1], q2 = fp[2], q3 = fp[3], q4 = fp[4];
        d3 p0 = mk3(q0.x, q0.y, q1.x), va = mk3(q1.y, q2.x, q2.y), vb = mk3(q3.x, q3.y, q4.x);
        grow(xf_point(g->fwd, p0));
        grow(xf_point(g->fwd, p0 + va));
        grow(xf_point(g->fwd, p0 + vb));
    };
    if (kind == PRIM_SPHERE) {
        const DGeom* g = S.geoms + idx;
        d3 c = xf_point(g->fwd, mk3(g->center[0], g->center[1], g->center[2]));
        double r = sqrt(g->radius2) * (1.0 + 1e-7);
        double cc[3] = {c.x, c.y, c.z};
        for (int a = 0; a < 3; a++) {
            const double* row = g->fwd + 4 * a;
            double e = r * sqrt(row[0] * row[0] + row[1] * row[1] + row[2] * row[2]);
            mn[a] = cc[a] - e;
            mx[a] = cc[a] + e;
        }
    } else if (kind == PRIM_TRI) {
        const DGeom* g = S.geoms + idx;
        face_bounds(g->first_face, g);
        face_bounds(g->first_face + 1, g);
    } else {
        double2 q4 = S.face_pts[(size_t)idx * RT_FACE_D2 + 4];
        face_bounds(idx, S.geoms + __double2loint(q4.y));
    }
    float maxabs = 0.f;
    for (int a = 0; a < 3; a++) {
        float l = __double2float_rd(mn[a]), h = __double2float_ru(mx[a]);
        lo[3 * (size_t)i + a] = l;
        hi[3 * (size_t)i + a] = h;
        float c = 0.5f * l + 0.5f * h;
        atomicMin(&gbounds[a], f2ord(c));
        atomicMax(&gbounds[3 + a], f2ord(c));
        maxabs = fmaxf(maxabs, fmaxf(fabsf(l), fabsf(h)));
    }
    atomicMax(&gbounds[6], f2ord(maxabs));
}

__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void k_morton(const float* __restrict__ lo, const float* __restrict__ hi, int n,
                         const int* __restrict__ gbounds, uint32_t* __restrict__ keys, int* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // Quantise all three axes with the SAME cell size (the largest centroid extent): cubic
    // Morton cells.  Normalising each axis separately would spend as many bits on the thin
    // axis of a 2.5-D scene (a height field) as on the long ones and cut it into "hills" and
    // "valleys" whose boxes overlap everywhere.
    uint32_t q[3];
    float ext = 0.f;
    for (int a = 0; a < 3; a++) ext = fmaxf(ext, ord2f(gbounds[3 + a]) - ord2f(gbounds[a]));
    for (int a = 0; a < 3; a++) {
        float cmin = ord2f(gbounds[a]);
        float c = 0.5f * lo[3 * (size_t)i + a] + 0.5f * hi[3 * (size_t)i + a];
        float u = ext > 0.f ? (c - cmin) / ext : 0.f;
        int v = (int)(u * 1024.f);
        q[a] = (uint32_t)min(max(v, 0), 1023);
    }
    keys[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    vals[i] = i;
}

// ---- Karras 2012 ---------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint32_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = keys[i], b = keys[j];
    if (a == b) return 32 + __clz(i ^ j);
    return __clz(a ^ b);
}

// child encoding inside the build: >= 0 internal node, < 0 leaf ~index (sorted position)
__global__ void k_karras(const uint32_t* __restrict__ keys, int n, int* __restrict__ left, int* __restrict__ right,
                         int* __restrict__ parent_int, int* __restrict__ parent_leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lc, rc;
    if (lo == gamma) { lc = ~gamma; parent_leaf[gamma] = i; } else { lc = gamma; parent_int[gamma] = i; }
    if (hi == gamma + 1) { rc = ~(gamma + 1); parent_leaf[gamma + 1] = i; } else { rc = gamma + 1; parent_int[gamma + 1] = i; }
    left[i] = lc;
    right[i] = rc;
    if (i == 0) parent_int[0] = -1;
}

__global__ void k_refit(int n, const int* __restrict__ vals, const float* __restrict__ plo,
                        const float* __restrict__ phi, const int* __restrict__ left, const int* __restrict__ right,
                        const int* __restrict__ parent_int, const int* __restrict__ parent_leaf,
                        float* __restrict__ ilo, float* __restrict__ ihi, int* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cur = parent_leaf[i];
    while (cur >= 0) {
        if (atomicAdd(&flags[cur], 1) == 0) return;     // first arrival waits for the sibling
        __threadfence();
        float lo[3], hi[3];
        for (int a = 0; a < 3; a++) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; }
        int ch[2] = {left[cur], right[cur]};
        for (int k = 0; k < 2; k++) {
            const float *bl, *bh;
            if (ch[k] < 0) {
                int p = vals[~ch[k]];
                bl = plo + 3 * (size_t)p; bh = phi + 3 * (size_t)p;
            } else {
                bl = ilo + 3 * (size_t)ch[k]; bh = ihi + 3 * (size_t)ch[k];
            }
            for (int a = 0; a < 3; a++) {
                lo[a] = fminf(lo[a], __ldcg(bl + a));
                hi[a] = fmaxf(hi[a], __ldcg(bh + a));
            }
        }
        for (int a = 0; a < 3; a++) { ilo[3 * (size_t)cur + a] = lo[a]; ihi[3 * (size_t)cur + a] = hi[a]; }
        __threadfence();
        cur = parent_int[cur];
    }
}

// depth of every internal node of one tree (root = 0) by walking the parent chain
__global__ void k_depth(int n, const int* __restrict__ parent_int, int* __restrict__ depth, int* __restrict__ max_depth) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = 0;
    for (int p = parent_int[i]; p >= 0; p = parent_int[p]) d++;
    depth[i] = d;
    // deepest internal node of the whole forest (bounds the traversal stack, see build_lbvh)
    const unsigned act = __activemask();
    const int m = __reduce_max_sync(act, d);
    if ((int)(threadIdx.x & 31) == __ffs(act) - 1) atomicMax(max_depth, m);
}

// Surface-area-guided choice of the children of the 4-wide nodes (one launch per level of the wide tree, top down):
// a binary node that is the root of a wide node in wave w starts from its two children and twice replaces the
// internal child of LARGEST surface area by that child's two children — the child a ray is most likely to enter
// anyway is the one whose box test is worth saving — and the internal children that remain become the roots of wave
// w + 1.  wave_of[]: wave of every root (-1: not a root); slots4[]: the four chosen children (BVH_DONE: none).
// Against pulling up the grandchildren of every even-depth node: 3.3 % fewer node visits per ray on the 1 M-triangle
// scene (19.27 -> 18.64 closest hit, 13.28 -> 12.83 any hit).
__global__ void k_choose_children(int n, int wave, int* __restrict__ wave_of, int4* __restrict__ slots4,
                                  const int* __restrict__ left, const int* __restrict__ right,
                                  const float* __restrict__ ilo, const float* __restrict__ ihi, int* __restrict__ wave_any) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1 || wave_of[i] != wave) return;
    int slots[4] = {left[i], right[i], BVH_DONE, BVH_DONE};
    int ns = 2;
    for (int rep = 0; rep < 2; rep++) {
        int best = -1;
        float best_area = -1.f;
        for (int k = 0; k < ns; k++) {
            if (slots[k] < 0) continue;
            const float* bl = ilo + 3 * (size_t)slots[k];
            const float* bh = ihi + 3 * (size_t)slots[k];
            const float ex = bh[0] - bl[0], ey = bh[1] - bl[1], ez = bh[2] - bl[2];
            const float area = ex * ey + ey * ez + ez * ex;
            if (area > best_area) { best_area = area; best = k; }
        }
        if (best < 0) break;
        const int c = slots[best];
        slots[best] = left[c];
        slots[ns++] = right[c];
    }
    for (int k = 0; k < ns; k++)
        if (slots[k] >= 0) wave_of[slots[k]] = wave + 1;
    slots4[i] = make_int4(slots[0], slots[1], slots[2], slots[3]);
    wave_any[wave] = 1;          // same value from every root of the wave: the deepest marked wave is the wide tree's depth
}

// Binary Karras nodes at EVEN depth become 4-wide nodes: each internal child (odd depth) is
// replaced by its own two children.  Wide nodes keep the binary node's index (+ offset), so
// no compaction pass is needed; odd-depth slots of the array stay unused.
//
// slots4 != null: the children were chosen by k_choose_children (wave_of[i] >= 0 marks the wide nodes) instead.
__global__ void k_pack_wide(int n, const int* __restrict__ vals, const int* __restrict__ codes,
                            const float* __restrict__ plo, const float* __restrict__ phi,
                            const int* __restrict__ left, const int* __restrict__ right,
                            const float* __restrict__ ilo, const float* __restrict__ ihi,
                            const int* __restrict__ depth, const int* __restrict__ gbounds,
                            BvhNode* __restrict__ nodes, int node_offset, const int* __restrict__ wave_of,
                            const int4* __restrict__ slots4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    if (slots4 ? wave_of[i] < 0 : (depth[i] & 1) != 0) return;
    // pad: a few FP32 ulps at the scene's largest coordinate (covers the slab test's rounding)
    const float pad = 2e-6f * fmaxf(ord2f(gbounds[6]), 1e-30f);
    int slots[4];
    int ns = 0;
    int ch[2] = {left[i], right[i]};
    if (!slots4) {
        for (int k = 0; k < 2; k++) {
            if (ch[k] < 0) slots[ns++] = ch[k];
            else { slots[ns++] = left[ch[k]]; slots[ns++] = right[ch[k]]; }
        }
    } else {
        const int4 c4 = slots4[i];
        const int c[4] = {c4.x, c4.y, c4.z, c4.w};
        for (int k = 0; k < 4; k++)
            if (c[k] != BVH_DONE) slots[ns++] = c[k];
    }
    float lo[3][4], hi[3][4];
    int ref[4];
    // empty slots get inverted boxes, which the sign-ordered slab test never hits (rt_device.cuh)
    for (int k = 0; k < 4; k++) {
        if (k >= ns) {
            for (int a = 0; a < 3; a++) { lo[a][k] = BVH_EMPTY_LO; hi[a][k] = BVH_EMPTY_HI; }
            ref[k] = BVH_DONE;
            continue;
        }
        const float *bl, *bh;
        if (slots[k] < 0) {
            int p = vals[~slots[k]];
            bl = plo + 3 * (size_t)p; bh = phi + 3 * (size_t)p;
            ref[k] = ~codes[p];
        } else {
            bl = ilo + 3 * (size_t)slots[k]; bh = ihi + 3 * (size_t)slots[k];
            ref[k] = slots[k] + node_offset;
        }
        for (int a = 0; a < 3; a++) {
            float l = bl[a], h = bh[a];
            lo[a][k] = l - pad - fabsf(l) * 2e-7f;
            hi[a][k] = h + pad + fabsf(h) * 2e-7f;
        }
    }
    BvhNode nd;
    nd.lox = make_float4(lo[0][0], lo[0][1], lo[0][2], lo[0][3]);
    nd.loy = make_float4(lo[1][0], lo[1][1], lo[1][2], lo[1][3]);
    nd.loz = make_float4(lo[2][0], lo[2][1], lo[2][2], lo[2][3]);
    nd.hix = make_float4(hi[0][0], hi[0][1], hi[0][2], hi[0][3]);
    nd.hiy = make_float4(hi[1][0], hi[1][1], hi[1][2], hi[1][3]);
    nd.hiz = make_float4(hi[2][0], hi[2][1], hi[2][2], hi[2][3]);
    nd.ref = make_int4(ref[0], ref[1], ref[2], ref[3]);
    nd.pad_ = make_int4(0, 0, 0, 0);
    nodes[node_offset + i] = nd;
}

#define RT_MAX_WAVES 256      // levels of the wide tree k_choose_children may walk (the traversal stack allows far fewer)
#define TAKE(var, type, count)                                                         \
    do {                                                                               \
        var = arena.take<type>((size_t)(count));                                       \
        if (!var) {                                                                    \
            snprintf(err, errlen, "LBVH scratch arena exhausted at %s", #var);         \
            goto fail;                                                                 \
        }                                                                              \
    } while (0)

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            snprintf(err, errlen, "%s failed: %s", #x, cudaGetErrorString(e_));        \
            goto fail;                                                                 \
        }                                                                              \
    } while (0)

// One LBVH per primitive GROUP (e.g. spheres/`tri`s vs mesh faces), joined under a short
// chain of super nodes: floating spheres mixed into a height-field's Morton order would
// otherwise stretch the leaf-level boxes of the mesh over the whole air space.
int build_lbvh(const DScene& S, const int* d_codes, const int* group_first_code, int n, const int* group_sizes, int ngroups,
               float extra_abs, cudaStream_t stream, DeviceArena& arena, BvhNode* nodes, size_t* out_count,
               int* launches, char* err, int errlen, float* centroid_bounds, int* out_max_depth, float* out_pad_scale) {
    *out_count = 0;
    float *plo = nullptr, *phi = nullptr, *ilo = nullptr, *ihi = nullptr;
    int *gb = nullptr, *vals0 = nullptr, *vals1 = nullptr, *hist = nullptr;
    int4* slots4 = nullptr;
    int* wave_any = nullptr;
    std::vector<int> h_wave_any(RT_MAX_WAVES, 0);
    int *left = nullptr, *right = nullptr, *pint = nullptr, *pleaf = nullptr, *flags = nullptr, *depth = nullptr;
    uint32_t *keys0 = nullptr, *keys1 = nullptr;
    const int T = 256;
    const int nblk = (n + T - 1) / T;
    int hb[7];
    int max_bin_depth = 0, max_wide_wave = 0;
    // RT_LBVH_COLLAPSE=even: round 1's fixed collapse (every even-depth binary node pulls up its grandchildren), for A/B
    const bool sah_collapse = !(getenv("RT_LBVH_COLLAPSE") && !strcmp(getenv("RT_LBVH_COLLAPSE"), "even"));
    float pad_scale = 0.f;
    struct Group { int start, n, root_ref, first_code; float box[6]; };
    Group groups[4];
    int K = 0;
    size_t total_nodes = 0;
    {
        if (n < 2 || ngroups > 4) return RT_OK;    // nothing to build (caller tests a single primitive directly)
        int acc = 0;
        for (int g = 0; g < ngroups; g++) {
            if (group_sizes[g] > 0) { groups[K].start = acc; groups[K].n = group_sizes[g]; groups[K].first_code = group_first_code[g]; K++; }
            acc += group_sizes[g];
        }
        const int nsuper = K > 1 ? 1 : 0;      // one 4-wide super node joins up to four trees
        total_nodes = (size_t)nsuper;
        for (int g = 0; g < K; g++) total_nodes += (size_t)(groups[g].n > 1 ? groups[g].n - 1 : 0);
        if (total_nodes > ((size_t)1 << 25)) {      // descend() addresses nodes with a 32-bit byte offset
            snprintf(err, errlen, "%zu LBVH nodes exceed the limit of 2^25 (4 GB node array)", total_nodes);
            return RT_ERR_LIMIT;
        }
        TAKE(plo, float, 3 * (size_t)n);
        TAKE(phi, float, 3 * (size_t)n);
        TAKE(gb, int, 16);
        int init[9] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000,
                       (int)0x80000000, 0, 0};
        CK(cudaMemcpyAsync(gb, init, sizeof(init), cudaMemcpyHostToDevice, stream));
        k_prim_bounds<<<nblk, T, 0, stream>>>(S, d_codes, n, plo, phi, gb);
        (*launches)++;
        CK(cudaMemcpyAsync(hb, gb, sizeof(hb), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        {   // fold the camera eye into the padding scale (positive float: ordered int == bits)
            float m = ord2f(hb[6]);
            if (extra_abs > m) m = extra_abs;
            pad_scale = m;
            int enc;
            memcpy(&enc, &m, 4);
            CK(cudaMemcpyAsync(gb + 6, &enc, sizeof(int), cudaMemcpyHostToDevice, stream));
        }
        TAKE(keys0, uint32_t, (size_t)n);
        TAKE(keys1, uint32_t, (size_t)n);
        TAKE(vals0, int, (size_t)n);
        TAKE(vals1, int, (size_t)n);
        TAKE(hist, int, 256 * (SORT_MAX_BLOCKS + 1));
        TAKE(left, int, (size_t)n);
        TAKE(right, int, (size_t)n);
        TAKE(pint, int, (size_t)n);
        TAKE(pleaf, int, (size_t)n);
        TAKE(flags, int, (size_t)n);
        TAKE(depth, int, (size_t)n);
        TAKE(ilo, float, 3 * (size_t)n);
        TAKE(ihi, float, 3 * (size_t)n);
        TAKE(slots4, int4, (size_t)n);
        TAKE(wave_any, int, RT_MAX_WAVES);
        CK(cudaMemsetAsync(wave_any, 0, sizeof(int) * RT_MAX_WAVES, stream));
        CK(cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)n, stream));
        k_morton<<<nblk, T, 0, stream>>>(plo, phi, n, gb, keys0, vals0);
        (*launches)++;
        size_t node_cursor = (size_t)nsuper;
        for (int g = 0; g < K; g++) {
            const int gs = groups[g].start, gn = groups[g].n;
            if (gn == 1) {
                groups[g].root_ref = ~groups[g].first_code;
                CK(cudaMemcpyAsync(groups[g].box, plo + 3 * (size_t)gs, 12, cudaMemcpyDeviceToHost, stream));
                CK(cudaMemcpyAsync(groups[g].box + 3, phi + 3 * (size_t)gs, 12, cudaMemcpyDeviceToHost, stream));
                continue;
            }
            const int gblk = (gn + T - 1) / T;
            // radix sort of this group's (key, prim index) pairs: 4 passes x 8 bits
            uint32_t *kin = keys0 + gs, *kout = keys1 + gs;
            int *vin = vals0 + gs, *vout = vals1 + gs;
            sort_pairs(stream, kin, kout, vin, vout, hist, gn, nullptr, 32, launches);
            // kin/vin now point at the sorted data
            k_karras<<<gblk, T, 0, stream>>>(kin, gn, left + gs, right + gs, pint + gs, pleaf + gs);
            k_refit<<<gblk, T, 0, stream>>>(gn, vin, plo, phi, left + gs, right + gs, pint + gs, pleaf + gs,
                                            ilo + 3 * (size_t)gs, ihi + 3 * (size_t)gs, flags + gs);
            k_depth<<<gblk, T, 0, stream>>>(gn, pint + gs, depth + gs, gb + 7);
            if (!sah_collapse) {
                k_pack_wide<<<gblk, T, 0, stream>>>(gn, vin, d_codes, plo, phi, left + gs, right + gs, ilo + 3 * (size_t)gs,
                                                    ihi + 3 * (size_t)gs, depth + gs, gb, nodes, (int)node_cursor, nullptr, nullptr);
            } else {
                // one light launch per level of the wide tree; the wide tree is at most as deep as the binary one
                int gdepth = 0;
                CK(cudaMemcpyAsync(&gdepth, gb + 7, sizeof(int), cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
                int* wave_of = flags + gs;            // k_refit is done with its arrival counters
                CK(cudaMemsetAsync(wave_of, 0xff, sizeof(int) * (size_t)gn, stream));
                CK(cudaMemsetAsync(wave_of, 0, sizeof(int), stream));                          // the group's root: wave 0
                if (gdepth + 1 > RT_MAX_WAVES) gdepth = RT_MAX_WAVES - 1;      // deeper trees fail the stack check below anyway
                for (int w = 0; w <= gdepth; w++) {
                    k_choose_children<<<gblk, T, 0, stream>>>(gn, w, wave_of, slots4 + gs, left + gs, right + gs, ilo + 3 * (size_t)gs,
                                                              ihi + 3 * (size_t)gs, wave_any);
                    (*launches)++;
                }
                k_pack_wide<<<gblk, T, 0, stream>>>(gn, vin, d_codes, plo, phi, left + gs, right + gs, ilo + 3 * (size_t)gs,
                                                    ihi + 3 * (size_t)gs, depth + gs, gb, nodes, (int)node_cursor, wave_of, slots4 + gs);
            }
            (*launches) += 4;
            groups[g].root_ref = (int)node_cursor;
            CK(cudaMemcpyAsync(groups[g].box, ilo + 3 * (size_t)gs, 12, cudaMemcpyDeviceToHost, stream));
            CK(cudaMemcpyAsync(groups[g].box + 3, ihi + 3 * (size_t)gs, 12, cudaMemcpyDeviceToHost, stream));
            node_cursor += (size_t)(gn - 1);
        }
        CK(cudaMemcpyAsync(&max_bin_depth, gb + 7, sizeof(int), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(h_wave_any.data(), wave_any, sizeof(int) * RT_MAX_WAVES, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        for (int w = 0; w < RT_MAX_WAVES; w++)
            if (h_wave_any[(size_t)w]) max_wide_wave = w;
        {   // The traversal defers at most 3 siblings per wide level.  Wide levels on a root-to-leaf path: one per
            // two binary levels (even-depth collapse) + the super node.  The stack (RT_SH_STACK shared + RT_STACK
            // local entries) must hold them all; Karras trees over 30-bit codes with index-split duplicates stay
            // far below this, so a scene that does not is rejected instead of silently dropping subtrees.
            const int wide_levels = (sah_collapse ? max_wide_wave + 1 : max_bin_depth / 2 + 1) + nsuper;
            if (3 * wide_levels > RT_STACK + RT_SH_STACK) {
                snprintf(err, errlen, "LBVH is %d binary levels deep: %d deferred entries exceed the traversal stack of %d",
                         max_bin_depth + 1, 3 * wide_levels, RT_STACK + RT_SH_STACK);
                return RT_ERR_LIMIT;
            }
            if (out_max_depth) *out_max_depth = max_bin_depth;
        }
        if (nsuper > 0) {
            const float pad = 2e-6f * fmaxf(pad_scale, 1e-30f);
            float lo[3][4], hi[3][4];
            int ref[4];
            for (int k = 0; k < 4; k++) {
                for (int a = 0; a < 3; a++) { lo[a][k] = BVH_EMPTY_LO; hi[a][k] = BVH_EMPTY_HI; }
                ref[k] = BVH_DONE;
                if (k >= K) continue;
                ref[k] = groups[k].root_ref;
                for (int a = 0; a < 3; a++) {
                    float l = groups[k].box[a], h = groups[k].box[3 + a];
                    lo[a][k] = l - pad - fabsf(l) * 2e-7f;
                    hi[a][k] = h + pad + fabsf(h) * 2e-7f;
                }
            }
            BvhNode nd;
            nd.lox = make_float4(lo[0][0], lo[0][1], lo[0][2], lo[0][3]);
            nd.loy = make_float4(lo[1][0], lo[1][1], lo[1][2], lo[1][3]);
            nd.loz = make_float4(lo[2][0], lo[2][1], lo[2][2], lo[2][3]);
            nd.hix = make_float4(hi[0][0], hi[0][1], hi[0][2], hi[0][3]);
            nd.hiy = make_float4(hi[1][0], hi[1][1], hi[1][2], hi[1][3]);
            nd.hiz = make_float4(hi[2][0], hi[2][1], hi[2][2], hi[2][3]);
            nd.ref = make_int4(ref[0], ref[1], ref[2], ref[3]);
            nd.pad_ = make_int4(0, 0, 0, 0);
            CK(cudaMemcpy(nodes, &nd, sizeof(BvhNode), cudaMemcpyHostToDevice));
        }
    }
    *out_count = total_nodes;
    if (out_pad_scale) *out_pad_scale = pad_scale;
    if (centroid_bounds)
        for (int a = 0; a < 6; a++) centroid_bounds[a] = ord2f(hb[a]);
    return RT_OK;
fail:
    return RT_ERR_CUDA;
}

}  // namespace rt
