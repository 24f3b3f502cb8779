/*
 * rr_oracle.c -- CPU restatement of the RadeonRays 4.1 Vulkan compute kernels on the HLBVH
 * build / refit / TLAS / traversal path.  TEST INFRASTRUCTURE ONLY (see rr_oracle.h): the product
 * library never links or calls this file.
 *
 * Every function cites the reference shader lines it restates (paths relative to
 * /root/reference/src/core/src/vlk/kernels unless noted).  Compile with
 *     gcc -O2 -std=c11 -ffp-contract=off -fopenmp -fPIC -shared
 * so that no multiply-add is fused except the explicit fmaf() calls that mirror GLSL fma().
 *
 * Parity status: "parity unpinned" at the hit-value level by the reference's own tests (they hold
 * no golden vectors); pinned structurally -- see rr_oracle.h and tests/test_oracle_*.py.
 */
#include "rr_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * small vector helpers with the evaluation order of GLSL dot()/cross()
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float x, y, z; } v3;

static inline v3    v3_make(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3    v3_ld(const float* p) { return v3_make(p[0], p[1], p[2]); }
static inline void  v3_st(float* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
static inline v3    v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
/* min/max with the zero ordering of IEEE 754-2019 minimum/maximum and of the GPU's FMNMX (-0.0 < +0.0) and
 * the NaN rule of C fminf/fmaxf (return the other operand).  GLSL leaves both cases undefined; fixing them
 * makes box bits reproducible (the Cornell box has both -0.0 and +0.0 y coordinates). */
static inline float rr_min(float a, float b)
{
    if (a != a) return b;
    if (b != b) return a;
    if (a < b) return a;
    if (b < a) return b;
    return signbit(a) ? a : b;
}
static inline float rr_max(float a, float b)
{
    if (a != a) return b;
    if (b != b) return a;
    if (a > b) return a;
    if (b > a) return b;
    return signbit(a) ? b : a;
}
static inline v3    v3_min(v3 a, v3 b) { return v3_make(rr_min(a.x, b.x), rr_min(a.y, b.y), rr_min(a.z, b.z)); }
static inline v3    v3_max(v3 a, v3 b) { return v3_make(rr_max(a.x, b.x), rr_max(a.y, b.y), rr_max(a.z, b.z)); }
static inline float v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* GLSL spec: cross(x,y) = (x1*y2 - y1*x2, x2*y0 - y2*x0, x0*y1 - y0*x1) */
static inline v3 v3_cross(v3 a, v3 b)
{
    return v3_make(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float    u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* common.h:68-90 -- order-preserving float<->uint map used for the atomicMin/Max scene box. */
static inline uint32_t float_to_ordered(float f)
{
    uint32_t v = f2u(f);
    v ^= (1u + ~(v >> 31)) | 0x80000000u;
    return v;
}
static inline float ordered_to_float(uint32_t v)
{
    v ^= ((v >> 31) - 1u) | 0x80000000u;
    return u2f(v);
}

int rro_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * B1 scene AABB -- lbvh_init_mesh.comp:61-77, lbvh_calc_mesh_aabb.comp:121-178.
 * The reference reduces with atomicMin/Max on the ordered-uint encoding; a min/max is order
 * independent so a serial sweep in the same encoding is bit-exact (-0.0 orders below +0.0).
 * ---------------------------------------------------------------------------------------------- */
void rro_scene_aabb(const float* verts, uint32_t sf, const uint32_t* idx, uint32_t n, float smin[3], float smax[3])
{
    uint32_t lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = float_to_ordered(FLT_MAX); hi[a] = float_to_ordered(-FLT_MAX); }
    for (uint32_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k)
        {
            const float* p = verts + (size_t)idx[3 * (size_t)i + k] * sf;
            for (int a = 0; a < 3; ++a)
            {
                uint32_t e = float_to_ordered(p[a]);
                if (e < lo[a]) lo[a] = e;
                if (e > hi[a]) hi[a] = e;
            }
        }
    for (int a = 0; a < 3; ++a) { smin[a] = ordered_to_float(lo[a]); smax[a] = ordered_to_float(hi[a]); }
}

/* common.h:251-258 */
static inline uint32_t expand_bits(uint32_t r)
{
    r = (r * 0x00010001u) & 0xFF0000FFu;
    r = (r * 0x00000101u) & 0x0F00F00Fu;
    r = (r * 0x00000011u) & 0xC30C30C3u;
    r = (r * 0x00000005u) & 0x49249249u;
    return r;
}
/* common.h:261-268.  clamp(x,lo,hi)=min(max(x,lo),hi); NaN -> 0 (fmaxf semantics). */
static inline uint32_t morton_code(v3 p)
{
    float x = rr_min(rr_max(p.x * 1024.0f, 0.0f), 1023.0f);
    float y = rr_min(rr_max(p.y * 1024.0f, 0.0f), 1023.0f);
    float z = rr_min(rr_max(p.z * 1024.0f, 0.0f), 1023.0f);
    return (expand_bits((uint32_t)x) << 2) | (expand_bits((uint32_t)y) << 1) | expand_bits((uint32_t)z);
}
static inline uint32_t morton_of_box(v3 bmin, v3 bmax, v3 smin, v3 smax)
{
    /* lbvh_calc_morton_codes_mesh.comp:114-125 */
    v3 ext = v3_sub(smax, smin);
    v3 c   = v3_make(0.5f * (bmin.x + bmax.x), 0.5f * (bmin.y + bmax.y), 0.5f * (bmin.z + bmax.z));
    v3 p   = v3_sub(c, smin);
    p      = v3_make(p.x / ext.x, p.y / ext.y, p.z / ext.z);
    return morton_code(p);
}

/* B2 -- lbvh_calc_morton_codes_mesh.comp:78-130 */
void rro_morton_codes(const float* verts, uint32_t sf, const uint32_t* idx, uint32_t n, const float smin[3],
                      const float smax[3], uint32_t* codes)
{
    v3 lo = v3_ld(smin), hi = v3_ld(smax);
    for (uint32_t i = 0; i < n; ++i)
    {
        v3 v0 = v3_ld(verts + (size_t)idx[3 * (size_t)i + 0] * sf);
        v3 v1 = v3_ld(verts + (size_t)idx[3 * (size_t)i + 1] * sf);
        v3 v2 = v3_ld(verts + (size_t)idx[3 * (size_t)i + 2] * sf);
        /* calculate_aabb_for_triangle, common.h:241-248 */
        v3 bmin = v3_min(v3_min(v0, v1), v2);
        v3 bmax = v3_max(v3_max(v0, v1), v2);
        codes[i] = morton_of_box(bmin, bmax, lo, hi);
    }
}

/* B3 -- vlk/radix_sort.cpp:202-215: 8 LSD passes of 4 bits, each a stable counting sort. */
void rro_sort_pairs(const uint32_t* keys, const uint32_t* vals, uint32_t n, uint32_t* out_keys, uint32_t* out_vals)
{
    uint32_t* k0 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
    uint32_t* v0 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
    uint32_t* k1 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
    uint32_t* v1 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
    memcpy(k0, keys, sizeof(uint32_t) * (size_t)n);
    memcpy(v0, vals, sizeof(uint32_t) * (size_t)n);
    for (uint32_t shift = 0; shift < 32; shift += 4)
    {
        size_t hist[17] = {0};
        for (uint32_t i = 0; i < n; ++i) hist[((k0[i] >> shift) & 0xf) + 1]++;
        for (int b = 0; b < 16; ++b) hist[b + 1] += hist[b];
        for (uint32_t i = 0; i < n; ++i)
        {
            size_t d = hist[(k0[i] >> shift) & 0xf]++;
            k1[d] = k0[i];
            v1[d] = v0[i];
        }
        uint32_t* t;
        t = k0; k0 = k1; k1 = t;
        t = v0; v0 = v1; v1 = t;
    }
    memcpy(out_keys, k0, sizeof(uint32_t) * (size_t)n);
    memcpy(out_vals, v0, sizeof(uint32_t) * (size_t)n);
    free(k0); free(v0); free(k1); free(v1);
}

/* ------------------------------------------------------------------------------------------------
 * B4 topology -- lbvh_emit_hierarchy_mesh.comp:78-217
 * ---------------------------------------------------------------------------------------------- */
static inline int ref_clz(uint32_t v) /* :78-81  32 - findMSB(v); findMSB(0) == -1 */
{
    int msb = -1;
    for (int b = 31; b >= 0; --b)
        if (v >> b) { msb = b; break; }
    return 32 - msb;
}
/* `codes` is uint32_t[] (the reference's 30-bit codes) or, for the 63-bit extension, uint64_t[] -- then the prefix is the one
 * dx/kernels/build_hlbvh_fallback.hlsl:95-108 sketches behind its compiled-out !USE_30BIT_MORTON_CODE branch:
 * clz64(l ^ r), ties broken by 64 + clz64 of the indices.  Selected by the sign of n_wide (see emit_hierarchy). */
typedef struct { const void* p; int wide; } code_array;
static inline int common_prefix(code_array codes, int64_t n, int64_t i1, int64_t i2) /* :85-103 */
{
    int64_t l = i1 < i2 ? i1 : i2, r = i1 < i2 ? i2 : i1;
    if (l < 0 || r >= n) return 0;
    if (codes.wide)
    {
        uint64_t lc = ((const uint64_t*)codes.p)[l], rc = ((const uint64_t*)codes.p)[r];
        return lc != rc ? 1 + __builtin_clzll(lc ^ rc) : 1 + 64 + __builtin_clzll((uint64_t)(l ^ r));
    }
    uint32_t lc = ((const uint32_t*)codes.p)[l], rc = ((const uint32_t*)codes.p)[r];
    return lc != rc ? ref_clz(lc ^ rc) : 32 + ref_clz((uint32_t)(l ^ r));
}
static void find_span(code_array codes, int64_t n, int64_t i, int64_t* sx, int64_t* sy) /* :105-140 */
{
    int diff = common_prefix(codes, n, i, i + 1) - common_prefix(codes, n, i, i - 1);
    int64_t d = (diff > 0) - (diff < 0);
    int dmin = common_prefix(codes, n, i, i - d);
    int64_t lmax = 2;
    while (common_prefix(codes, n, i, i + lmax * d) > dmin) lmax *= 2;
    int64_t l = 0, t = lmax;
    do
    {
        t /= 2;
        if (common_prefix(codes, n, i, i + (l + t) * d) > dmin) l += t;
    } while (t > 1);
    int64_t a = i, b = i + l * d;
    int64_t lo = a < b ? a : b, hi = a < b ? b : a;
    *sx = lo < 0 ? 0 : lo;
    *sy = hi > n - 1 ? n - 1 : hi;
}
static int64_t find_split(code_array codes, int64_t n, int64_t sx, int64_t sy) /* :142-168 */
{
    int64_t left = sx, right = sy;
    int ident = common_prefix(codes, n, left, right);
    do
    {
        int64_t m = (right + left) / 2;
        if (common_prefix(codes, n, left, m) > ident) left = m; else right = m;
    } while (right > left + 1);
    return left;
}

static void emit_hierarchy(code_array codes, const uint32_t* refs, uint32_t n, rro_node* nodes);
void rro_emit_hierarchy(const uint32_t* codes, const uint32_t* refs, uint32_t n, rro_node* nodes) /* :170-217 */
{
    code_array c = {codes, 0};
    emit_hierarchy(c, refs, n, nodes);
}
void rro_emit_hierarchy64(const uint64_t* codes, const uint32_t* refs, uint32_t n, rro_node* nodes)
{
    code_array c = {codes, 1};
    emit_hierarchy(c, refs, n, nodes);
}
static void emit_hierarchy(code_array codes, const uint32_t* refs, uint32_t n, rro_node* nodes)
{
    const uint32_t leaf0 = n - 1;
    memset(nodes, 0, sizeof(rro_node) * (size_t)(2 * (size_t)n - 1));
    for (uint32_t j = 0; j < n; ++j)
    {
        nodes[leaf0 + j].child0 = RRO_INVALID;
        nodes[leaf0 + j].child1 = refs[j];
        nodes[leaf0 + j].update = 0;
    }
    if (n == 1) { nodes[0].parent = RRO_INVALID; return; } /* reference leaves it unwritten */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n - 1; ++i)
    {
        int64_t sx, sy;
        find_span(codes, n, i, &sx, &sy);
        int64_t split = find_split(codes, n, sx, sy);
        uint32_t l = (split == sx) ? leaf0 + (uint32_t)split : (uint32_t)split;
        uint32_t r = (split + 1 == sy) ? leaf0 + (uint32_t)split + 1 : (uint32_t)split + 1;
        nodes[i].child0 = l;
        nodes[i].child1 = r;
        nodes[i].update = 0;
        nodes[l].parent = (uint32_t)i;
        nodes[r].parent = (uint32_t)i;
        if (i == 0) nodes[0].parent = RRO_INVALID;
    }
}

/* ------------------------------------------------------------------------------------------------
 * B5 fit / refit -- lbvh_fit_aabb_mesh.comp:94-205 (refit: same kernel with UPDATE_KERNEL,
 * vlk/update_hlbvh.cpp:118-185).  The shader climbs with atomicExchange on `update`; min/max is
 * order independent, so a serial climb with an arrival counter yields identical boxes.  `update`
 * itself is scratch state (the reference leaves 1 in internal nodes); we leave 0 and exclude it
 * from parity.
 * ---------------------------------------------------------------------------------------------- */
static inline int is_internal(const rro_node* nd) { return nd->child0 != RRO_INVALID; }

static void node_box(const rro_node* nd, v3* bmin, v3* bmax) /* calculate_aabb_for_node :94-118 */
{
    if (is_internal(nd))
    {
        *bmin = v3_min(v3_ld(nd->aabb0_min_or_v0), v3_ld(nd->aabb1_min_or_v2));
        *bmax = v3_max(v3_ld(nd->aabb0_max_or_v1), v3_ld(nd->aabb1_max_or_v3));
    }
    else
    {
        v3 a = v3_ld(nd->aabb0_min_or_v0), b = v3_ld(nd->aabb0_max_or_v1), c = v3_ld(nd->aabb1_min_or_v2);
        *bmin = v3_min(v3_min(a, b), c);
        *bmax = v3_max(v3_max(a, b), c);
    }
}
static void refresh_internal(rro_node* nodes, uint32_t addr) /* :172-190 */
{
    v3 lo, hi;
    node_box(&nodes[nodes[addr].child0], &lo, &hi);
    v3_st(nodes[addr].aabb0_min_or_v0, lo);
    v3_st(nodes[addr].aabb0_max_or_v1, hi);
    node_box(&nodes[nodes[addr].child1], &lo, &hi);
    v3_st(nodes[addr].aabb1_min_or_v2, lo);
    v3_st(nodes[addr].aabb1_max_or_v3, hi);
}

void rro_fit_mesh(rro_node* nodes, uint32_t n, const float* verts, uint32_t sf, const uint32_t* idx)
{
    const uint32_t leaf0 = n - 1;
    uint8_t* arrived = (uint8_t*)calloc(n ? n : 1, 1);
    for (uint32_t j = 0; j < n; ++j)
    {
        rro_node* lf = &nodes[leaf0 + j];
        uint32_t tri = lf->child1;
        v3_st(lf->aabb0_min_or_v0, v3_ld(verts + (size_t)idx[3 * (size_t)tri + 0] * sf));
        v3_st(lf->aabb0_max_or_v1, v3_ld(verts + (size_t)idx[3 * (size_t)tri + 1] * sf));
        v3_st(lf->aabb1_min_or_v2, v3_ld(verts + (size_t)idx[3 * (size_t)tri + 2] * sf));
        lf->aabb1_max_or_v3[0] = lf->aabb1_max_or_v3[1] = lf->aabb1_max_or_v3[2] = 0.0f; /* unwritten in ref */
        uint32_t addr = lf->parent;
        while (addr != RRO_INVALID)
        {
            if (!arrived[addr]) { arrived[addr] = 1; break; }
            refresh_internal(nodes, addr);
            addr = nodes[addr].parent;
        }
    }
    free(arrived);
}

/* ------------------------------------------------------------------------------------------------
 * B6 treelet restructuring -- vlk/restructure_hlbvh.cpp:151-277, init_primitive_count.comp:63-80,
 * find_treelet_roots.comp:63-98, restructure_bvh.comp:71-542.  One 64-thread work-group per
 * treelet root in the shader; here each "thread" section is executed literally in lane order.
 * ---------------------------------------------------------------------------------------------- */
#define TL 7
#define C_INT 1.2f

static void treelet_node_box(const rro_node* nodes, uint32_t addr, uint32_t n, v3* lo, v3* hi) /* GetNodeAabb :71-88 */
{
    const rro_node* nd = &nodes[addr];
    v3 mn = v3_make(FLT_MAX, FLT_MAX, FLT_MAX), mx = v3_make(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    v3 p;
    p = v3_ld(nd->aabb0_min_or_v0); mn = v3_min(mn, p); mx = v3_max(mx, p);
    p = v3_ld(nd->aabb0_max_or_v1); mn = v3_min(mn, p); mx = v3_max(mx, p);
    p = v3_ld(nd->aabb1_min_or_v2); mn = v3_min(mn, p); mx = v3_max(mx, p);
    if (addr < n - 1) { p = v3_ld(nd->aabb1_max_or_v3); mn = v3_min(mn, p); mx = v3_max(mx, p); }
    *lo = mn; *hi = mx;
}
static inline float box_area(v3 lo, v3 hi) /* GetAabbSurfaceArea :90-94: 2*dot(e, e.zxy) */
{
    v3 e = v3_sub(hi, lo);
    return 2.0f * v3_dot(e, v3_make(e.z, e.x, e.y));
}
static float treelet_node_area(const rro_node* nodes, uint32_t addr, uint32_t n)
{
    v3 lo, hi;
    treelet_node_box(nodes, addr, n, &lo, &hi);
    return box_area(lo, hi);
}
static inline int popc(uint32_t v) { return __builtin_popcount(v); }

static const uint32_t k_perm6[8] = {0x3f, 0x5f, 0x6f, 0x77, 0x7b, 0x7d, 0x7e, 0x00}; /* s_bit_permutations row 4 */

static void restructure_treelet(rro_node* nodes, uint32_t n, uint32_t root)
{
    uint32_t leaves[TL], internal[TL - 1];
    float    areas7[TL];
    /* ---- form treelet (thread 0), :170-217 ---- */
    internal[0] = root;
    leaves[0]   = nodes[root].child0;
    leaves[1]   = nodes[root].child1;
    areas7[0]   = treelet_node_area(nodes, leaves[0], n);
    areas7[1]   = treelet_node_area(nodes, leaves[1], n);
    uint32_t size = 2;
    while (size < TL)
    {
        float    largest = 0.0f;
        uint32_t pick = 0, pick_slot = 0;
        for (uint32_t i = 0; i < size; ++i)
        {
            if (leaves[i] < n - 1) /* IS_INTERNAL_NODE */
            {
                float a = areas7[i];
                if (largest == 0.0f || a > largest) { largest = a; pick = leaves[i]; pick_slot = i; }
            }
        }
        internal[size - 1] = pick;
        uint32_t c0 = nodes[pick].child0, c1 = nodes[pick].child1;
        leaves[pick_slot] = c0;
        leaves[size]      = c1;
        areas7[pick_slot] = treelet_node_area(nodes, c0, n);
        areas7[size]      = treelet_node_area(nodes, c1, n);
        ++size;
    }

    /* ---- subset areas, :223-240 ---- */
    float    area[128], cost[128];
    uint32_t part[128];
    v3 lo7[TL], hi7[TL];
    for (int i = 0; i < TL; ++i) treelet_node_box(nodes, leaves[i], n, &lo7[i], &hi7[i]);
    for (uint32_t m = 0; m < 128; ++m)
    {
        v3 lo = v3_make(FLT_MAX, FLT_MAX, FLT_MAX), hi = v3_make(-FLT_MAX, -FLT_MAX, -FLT_MAX);
        for (int i = 0; i < TL; ++i)
            if (m & (1u << i)) { lo = v3_min(lo, lo7[i]); hi = v3_max(hi, hi7[i]); }
        area[m] = box_area(lo, hi);
        cost[m] = 0.0f;
        part[m] = 0;
    }
    /* singletons :249-252: C_INT * area / 1 */
    for (int i = 0; i < TL; ++i) cost[1u << i] = C_INT * area[1u << i] / 1.0f;

    for (int bits = 2; bits <= TL; ++bits)
    {
        if (bits == TL - 1)
        {
            /* :264-338: 7 masks x 8 lanes, 4 partitions per lane, then an 8-wide tree min with '>' */
            for (int mi = 0; mi < 7; ++mi)
            {
                uint32_t mask = k_perm6[mi];
                float    lane_cost[8];
                uint32_t lane_part[8];
                uint32_t delta = (mask - 1u) & mask;
                uint32_t p0    = (0u - delta) & mask;
                for (uint32_t t = 0; t < 8; ++t)
                {
                    uint32_t p = (p0 - delta * t * 4u) & mask;
                    float    lowest = FLT_MAX;
                    uint32_t best = 0;
                    int counter = 0;
                    do
                    {
                        float c = cost[p] + cost[mask ^ p];
                        if (best == 0 || c < lowest) { lowest = c; best = p; }
                        p = (p - delta) & mask;
                        ++counter;
                    } while (p != 0 && counter < 4);
                    lane_cost[t] = lowest;
                    lane_part[t] = best;
                }
                for (uint32_t l = 0; l < 3; ++l)
                    for (uint32_t t = 0; t < 8; ++t)
                        if ((t & ((1u << (l + 1)) - 1u)) == 0)
                            if (lane_cost[t] > lane_cost[t + (1u << l)])
                            {
                                lane_cost[t] = lane_cost[t + (1u << l)];
                                lane_part[t] = lane_part[t + (1u << l)];
                            }
                cost[mask] = C_INT * area[mask] + lane_cost[0];
                part[mask] = lane_part[0];
            }
        }
        else if (bits == TL)
        {
            /* :339-383: 63 candidate partitions (lidx+1)<<1; the tree min never looks at index 62 */
            uint32_t mask = 0x7f;
            float    pc[64];
            uint32_t pm[64];
            for (uint32_t t = 0; t < 63; ++t)
            {
                uint32_t p = (t + 1u) << 1;
                pc[t] = cost[p] + cost[mask ^ p];
                pm[t] = p;
            }
            const uint32_t unsorted = 63;
            for (uint32_t l = 0; l < 6; ++l) /* ceil(log2(63)) */
                for (uint32_t t = 0; t < 64; ++t)
                    if ((t & ((1u << (l + 1)) - 1u)) == 0 && (t + (1u << l)) < unsorted - 1)
                        if (pc[t] > pc[t + (1u << l)])
                        {
                            pc[t] = pc[t + (1u << l)];
                            pm[t] = pm[t + (1u << l)];
                        }
            cost[mask] = C_INT * area[mask] + pc[0];
            part[mask] = pm[0];
        }
        else
        {
            /* :385-413 one mask per lane, Karras & Aila 2013 Algorithm 3 enumeration */
            for (uint32_t mask = 1; mask < 128; ++mask)
            {
                if (popc(mask) != bits) continue;
                float    lowest = FLT_MAX;
                uint32_t best = 0;
                uint32_t delta = (mask - 1u) & mask;
                uint32_t p     = (0u - delta) & mask;
                do
                {
                    float c = cost[p] + cost[mask ^ p];
                    if (best == 0 || c < lowest) { lowest = c; best = p; }
                    p = (p - delta) & mask;
                } while (p != 0);
                cost[mask] = C_INT * area[mask] + lowest;
                part[mask] = (popc(mask) & 1) ? best : (mask ^ best); /* :412 */
            }
        }
    }

    /* ---- rebuild topology, :417-475 ---- */
    struct { uint32_t mask, node; } stack[TL];
    uint32_t allocated = 1, sp = 1;
    stack[0].mask = 0x7f;
    stack[0].node = internal[0];
    while (sp > 0)
    {
        uint32_t pmask = stack[sp - 1].mask, pnode = stack[sp - 1].node;
        --sp;
        uint32_t lmask = part[pmask], lnode, rmask = pmask ^ lmask, rnode;
        if (popc(lmask) > 1) { lnode = internal[allocated++]; stack[sp].mask = lmask; stack[sp].node = lnode; ++sp; }
        else lnode = leaves[31 - __builtin_clz(lmask)];
        if (popc(rmask) > 1) { rnode = internal[allocated++]; stack[sp].mask = rmask; stack[sp].node = rnode; ++sp; }
        else rnode = leaves[31 - __builtin_clz(rmask)];
        nodes[pnode].child0 = lnode;
        nodes[pnode].child1 = rnode;
        nodes[lnode].parent = pnode;
        nodes[rnode].parent = pnode;
    }
    /* ---- refit the 6 internal nodes bottom-up, :480-494 ---- */
    for (int j = TL - 2; j >= 0; --j)
    {
        uint32_t in = internal[j];
        v3 lo, hi;
        treelet_node_box(nodes, nodes[in].child0, n, &lo, &hi);
        v3_st(nodes[in].aabb0_min_or_v0, lo);
        v3_st(nodes[in].aabb0_max_or_v1, hi);
        treelet_node_box(nodes, nodes[in].child1, n, &lo, &hi);
        v3_st(nodes[in].aabb1_min_or_v2, lo);
        v3_st(nodes[in].aabb1_max_or_v3, hi);
    }
}

void rro_restructure_round(rro_node* nodes, uint32_t n, uint32_t min_prims)
{
    if (n < 2) return;
    const uint32_t leaf0 = n - 1;
    uint32_t* counts = (uint32_t*)calloc(2 * (size_t)n - 1, sizeof(uint32_t));
    uint32_t* roots  = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    uint32_t  nroots = 0;
    for (uint32_t j = 0; j < n; ++j) counts[leaf0 + j] = 1; /* init_primitive_count.comp:66-76 */
    /* find_treelet_roots.comp:63-98 (atomicExchange emulated in leaf order; the result set is order independent) */
    for (uint32_t j = 0; j < n; ++j)
    {
        uint32_t prim = 1, index = nodes[leaf0 + j].parent;
        while (index != RRO_INVALID)
        {
            uint32_t old = counts[index];
            counts[index] = prim;
            prim += old;
            if (old > 0)
            {
                if (prim >= min_prims) { roots[nroots++] = index; break; }
            }
            else break;
            index = nodes[index].parent;
        }
    }
    /* restructure_bvh.comp:155-541: per root, optimise then climb while the atomicAdd gate is open */
    for (uint32_t r = 0; r < nroots; ++r)
    {
        uint32_t node = roots[r];
        for (;;)
        {
            restructure_treelet(nodes, n, node);
            node = nodes[node].parent;
            if (node == RRO_INVALID) break;
            uint32_t old = counts[node]++;
            if (old == 0) break;
            /* :513-521 recompute the gate node's child boxes (a no-op numerically: same leaf set) */
            v3 lo, hi;
            treelet_node_box(nodes, nodes[node].child0, n, &lo, &hi);
            v3_st(nodes[node].aabb0_min_or_v0, lo);
            v3_st(nodes[node].aabb0_max_or_v1, hi);
            treelet_node_box(nodes, nodes[node].child1, n, &lo, &hi);
            v3_st(nodes[node].aabb1_min_or_v2, lo);
            v3_st(nodes[node].aabb1_max_or_v3, hi);
        }
    }
    free(counts);
    free(roots);
}

void rro_restructure(rro_node* nodes, uint32_t n) /* restructure_hlbvh.cpp:151-277 */
{
    rro_restructure_round(nodes, n, 64);
    rro_restructure_round(nodes, n, 128);
    rro_restructure_round(nodes, n, 256);
}

void rro_build_blas(const float* verts, uint32_t sf, const uint32_t* idx, uint32_t n, int restructure, rro_node* nodes,
                    uint32_t* sorted_codes, uint32_t* sorted_refs)
{
    float smin[3], smax[3];
    uint32_t* codes = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    uint32_t* refs  = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    uint32_t* sc    = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    uint32_t* sr    = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    rro_scene_aabb(verts, sf, idx, n, smin, smax);
    rro_morton_codes(verts, sf, idx, n, smin, smax, codes);
    for (uint32_t i = 0; i < n; ++i) refs[i] = i;
    rro_sort_pairs(codes, refs, n, sc, sr);
    rro_emit_hierarchy(sc, sr, n, nodes);
    rro_fit_mesh(nodes, n, verts, sf, idx);
    if (restructure) rro_restructure(nodes, n);
    if (sorted_codes) memcpy(sorted_codes, sc, sizeof(uint32_t) * (size_t)n);
    if (sorted_refs) memcpy(sorted_refs, sr, sizeof(uint32_t) * (size_t)n);
    free(codes); free(refs); free(sc); free(sr);
}

/* ---- 63-bit Morton extension (BASELINE north_star "30/63-bit Morton codes") --------------------------------------------
 * The reference ships 30-bit codes only; its DX fallback kernel keeps a compiled-out 64-bit branch of delta()
 * (dx/kernels/build_hlbvh_fallback.hlsl:16,95-108) and nothing else, so there is no reference output to be exact against.
 * Defined here as the same pipeline with 21 bits per axis: p = (centroid - sceneMin) / extent as in the 30-bit path,
 * clamp(p * 2^21, 0, 2^21 - 1) -> uint -> 3-dilate to 63 bits, code = x << 2 | y << 1 | z; stable sort on the 64-bit key;
 * Karras topology with the 64-bit prefix above; same fit / restructure. */
static inline uint64_t expand_bits21(uint32_t v)
{
    uint64_t x = v & 0x1FFFFFu;
    x = (x | (x << 32)) & 0x001F00000000FFFFull;
    x = (x | (x << 16)) & 0x001F0000FF0000FFull;
    x = (x | (x << 8)) & 0x100F00F00F00F00Full;
    x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}
void rro_morton_codes63(const float* verts, uint32_t sf, const uint32_t* idx, uint32_t n, const float smin[3],
                        const float smax[3], uint64_t* codes)
{
    v3 lo = v3_ld(smin), hi = v3_ld(smax);
    v3 ext = v3_sub(hi, lo);
    for (uint32_t i = 0; i < n; ++i)
    {
        v3 v0 = v3_ld(verts + (size_t)idx[3 * (size_t)i + 0] * sf);
        v3 v1 = v3_ld(verts + (size_t)idx[3 * (size_t)i + 1] * sf);
        v3 v2 = v3_ld(verts + (size_t)idx[3 * (size_t)i + 2] * sf);
        v3 bmin = v3_min(v3_min(v0, v1), v2);
        v3 bmax = v3_max(v3_max(v0, v1), v2);
        v3 c = v3_make(0.5f * (bmin.x + bmax.x), 0.5f * (bmin.y + bmax.y), 0.5f * (bmin.z + bmax.z));
        v3 p = v3_sub(c, lo);
        p    = v3_make(p.x / ext.x, p.y / ext.y, p.z / ext.z);
        float x = rr_min(rr_max(p.x * 2097152.0f, 0.0f), 2097151.0f);
        float y = rr_min(rr_max(p.y * 2097152.0f, 0.0f), 2097151.0f);
        float z = rr_min(rr_max(p.z * 2097152.0f, 0.0f), 2097151.0f);
        codes[i] = (expand_bits21((uint32_t)x) << 2) | (expand_bits21((uint32_t)y) << 1) | expand_bits21((uint32_t)z);
    }
}
void rro_build_blas63(const float* verts, uint32_t sf, const uint32_t* idx, uint32_t n, int restructure, rro_node* nodes,
                      uint64_t* sorted_codes, uint32_t* sorted_refs)
{
    float smin[3], smax[3];
    uint64_t* k0 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
    uint64_t* k1 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
    uint32_t* v0 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    uint32_t* v1 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n);
    rro_scene_aabb(verts, sf, idx, n, smin, smax);
    rro_morton_codes63(verts, sf, idx, n, smin, smax, k0);
    for (uint32_t i = 0; i < n; ++i) v0[i] = i;
    for (uint32_t shift = 0; shift < 64; shift += 8) /* stable LSD counting sort */
    {
        size_t hist[257] = {0};
        for (uint32_t i = 0; i < n; ++i) hist[((k0[i] >> shift) & 0xff) + 1]++;
        for (int b = 0; b < 256; ++b) hist[b + 1] += hist[b];
        for (uint32_t i = 0; i < n; ++i)
        {
            size_t d = hist[(k0[i] >> shift) & 0xff]++;
            k1[d] = k0[i];
            v1[d] = v0[i];
        }
        uint64_t* tk = k0; k0 = k1; k1 = tk;
        uint32_t* tv = v0; v0 = v1; v1 = tv;
    }
    rro_emit_hierarchy64(k0, v0, n, nodes);
    rro_fit_mesh(nodes, n, verts, sf, idx);
    if (restructure) rro_restructure(nodes, n);
    if (sorted_codes) memcpy(sorted_codes, k0, sizeof(uint64_t) * (size_t)n);
    if (sorted_refs) memcpy(sorted_refs, v0, sizeof(uint32_t) * (size_t)n);
    free(k0); free(k1); free(v0); free(v1);
}

/* ------------------------------------------------------------------------------------------------
 * B7 TLAS -- lbvh_{init,calc_scene_aabb,calc_morton_codes,emit_hierarchy,fit_aabb}_scene.comp,
 * common.h:270-342, host side vlk/intersector.cpp:205-264.
 * ---------------------------------------------------------------------------------------------- */
static inline v3 xform_point(const float* m, v3 p) /* transform_point common.h:270-278 */
{
    return v3_make(v3_dot(v3_make(m[0], m[1], m[2]), p) + m[3], v3_dot(v3_make(m[4], m[5], m[6]), p) + m[7],
                   v3_dot(v3_make(m[8], m[9], m[10]), p) + m[11]);
}
static void xform_aabb(v3* lo, v3* hi, const float* m, int quirk) /* transform_aabb common.h:280-308 */
{
    v3 a = *lo, b = *hi, c[8];
    c[0] = a;
    c[1] = v3_make(a.x, a.y, b.z);
    c[2] = v3_make(a.x, b.y, a.z);
    c[3] = v3_make(a.x, b.y, b.z);
    c[4] = v3_make(b.x, a.y, b.z);
    c[5] = v3_make(b.x, b.y, a.z);
    c[6] = quirk ? b : v3_make(b.x, a.y, a.z); /* reference: p6 == p7 == pmax, (max,min,min) missing */
    c[7] = b;
    v3 p = xform_point(m, c[0]);
    v3 mn = p, mx = p;
    for (int i = 1; i < 8; ++i) { p = xform_point(m, c[i]); mn = v3_min(mn, p); mx = v3_max(mx, p); }
    *lo = mn; *hi = mx;
}
/* Affine inverse.  GLSL inverse(mat4) (common.h:326-342) is implementation defined; we fix the
 * adjugate/determinant form below and the CUDA kernel evaluates the same expression tree. */
static void affine_inverse(const float* m, float* r)
{
    float a00 = m[0], a01 = m[1], a02 = m[2], tx = m[3];
    float a10 = m[4], a11 = m[5], a12 = m[6], ty = m[7];
    float a20 = m[8], a21 = m[9], a22 = m[10], tz = m[11];
    float c00 = a11 * a22 - a12 * a21, c01 = a02 * a21 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    float c10 = a12 * a20 - a10 * a22, c11 = a00 * a22 - a02 * a20, c12 = a02 * a10 - a00 * a12;
    float c20 = a10 * a21 - a11 * a20, c21 = a01 * a20 - a00 * a21, c22 = a00 * a11 - a01 * a10;
    float det = (a00 * c00 + a01 * c10) + a02 * c20;
    float inv = 1.0f / det;
    float r00 = c00 * inv, r01 = c01 * inv, r02 = c02 * inv;
    float r10 = c10 * inv, r11 = c11 * inv, r12 = c12 * inv;
    float r20 = c20 * inv, r21 = c21 * inv, r22 = c22 * inv;
    r[0] = r00; r[1] = r01; r[2] = r02;  r[3]  = -((r00 * tx + r01 * ty) + r02 * tz);
    r[4] = r10; r[5] = r11; r[6] = r12;  r[7]  = -((r10 * tx + r11 * ty) + r12 * tz);
    r[8] = r20; r[9] = r21; r[10] = r22; r[11] = -((r20 * tx + r21 * ty) + r22 * tz);
}
static void instance_box(const rro_node* root, uint32_t tris, const float* m, int quirk, v3* lo, v3* hi)
{
    if (tris == 1)
    { /* the reference reads a leaf root as if it held two boxes; we bound its 3 vertices */
        v3 a = v3_ld(root->aabb0_min_or_v0), b = v3_ld(root->aabb0_max_or_v1), c = v3_ld(root->aabb1_min_or_v2);
        *lo = v3_min(v3_min(a, b), c);
        *hi = v3_max(v3_max(a, b), c);
    }
    else
    { /* lbvh_fit_aabb_scene.comp:119-123 */
        *lo = v3_min(v3_ld(root->aabb0_min_or_v0), v3_ld(root->aabb1_min_or_v2));
        *hi = v3_max(v3_ld(root->aabb0_max_or_v1), v3_ld(root->aabb1_max_or_v3));
    }
    xform_aabb(lo, hi, m, quirk);
}

void rro_build_tlas(const rro_node* const* blas_roots, const uint32_t* blas_tris, const float* transforms, uint32_t n,
                    int quirk, rro_node* nodes, float* out_transforms)
{
    v3* lo = (v3*)malloc(sizeof(v3) * n);
    v3* hi = (v3*)malloc(sizeof(v3) * n);
    uint32_t* codes = (uint32_t*)malloc(4 * (size_t)n);
    uint32_t* refs  = (uint32_t*)malloc(4 * (size_t)n);
    uint32_t* sc    = (uint32_t*)malloc(4 * (size_t)n);
    uint32_t* sr    = (uint32_t*)malloc(4 * (size_t)n);
    uint32_t slo[3], shi[3];
    for (int a = 0; a < 3; ++a) { slo[a] = float_to_ordered(FLT_MAX); shi[a] = float_to_ordered(-FLT_MAX); }
    for (uint32_t i = 0; i < n; ++i)
    { /* lbvh_calc_scene_aabb.comp:131-163 */
        instance_box(blas_roots[i], blas_tris[i], transforms + 12 * (size_t)i, quirk, &lo[i], &hi[i]);
        const float l[3] = {lo[i].x, lo[i].y, lo[i].z}, h[3] = {hi[i].x, hi[i].y, hi[i].z};
        for (int a = 0; a < 3; ++a)
        {
            uint32_t e0 = float_to_ordered(l[a]), e1 = float_to_ordered(h[a]);
            if (e0 < slo[a]) slo[a] = e0;
            if (e1 > shi[a]) shi[a] = e1;
        }
    }
    v3 smin = v3_make(ordered_to_float(slo[0]), ordered_to_float(slo[1]), ordered_to_float(slo[2]));
    v3 smax = v3_make(ordered_to_float(shi[0]), ordered_to_float(shi[1]), ordered_to_float(shi[2]));
    for (uint32_t i = 0; i < n; ++i)
    { /* lbvh_calc_morton_codes_scene.comp:84-112 */
        codes[i] = morton_of_box(lo[i], hi[i], smin, smax);
        refs[i]  = i;
    }
    rro_sort_pairs(codes, refs, n, sc, sr);
    rro_emit_hierarchy(sc, sr, n, nodes);
    /* lbvh_fit_aabb_scene.comp:98-167 */
    const uint32_t leaf0 = n - 1;
    uint8_t* arrived = (uint8_t*)calloc(n ? n : 1, 1);
    for (uint32_t j = 0; j < n; ++j)
    {
        rro_node* lf = &nodes[leaf0 + j];
        uint32_t inst = lf->child1;
        affine_inverse(transforms + 12 * (size_t)inst, out_transforms + 24 * (size_t)inst);
        memcpy(out_transforms + 24 * (size_t)inst + 12, transforms + 12 * (size_t)inst, 48);
        v3_st(lf->aabb0_min_or_v0, lo[inst]);
        v3_st(lf->aabb0_max_or_v1, hi[inst]);
        v3_st(lf->aabb1_min_or_v2, lo[inst]);
        v3_st(lf->aabb1_max_or_v3, hi[inst]);
        uint32_t addr = lf->parent;
        while (addr != RRO_INVALID)
        {
            if (!arrived[addr]) { arrived[addr] = 1; break; }
            /* internal TLAS nodes: children are always read as two-box nodes (:50-63) */
            for (int c = 0; c < 2; ++c)
            {
                const rro_node* ch = &nodes[c ? nodes[addr].child1 : nodes[addr].child0];
                v3 l = v3_min(v3_ld(ch->aabb0_min_or_v0), v3_ld(ch->aabb1_min_or_v2));
                v3 h = v3_max(v3_ld(ch->aabb0_max_or_v1), v3_ld(ch->aabb1_max_or_v3));
                v3_st(c ? nodes[addr].aabb1_min_or_v2 : nodes[addr].aabb0_min_or_v0, l);
                v3_st(c ? nodes[addr].aabb1_max_or_v3 : nodes[addr].aabb0_max_or_v1, h);
            }
            addr = nodes[addr].parent;
        }
    }
    free(arrived); free(lo); free(hi); free(codes); free(refs); free(sc); free(sr);
}

/* ------------------------------------------------------------------------------------------------
 * validation helpers
 * ---------------------------------------------------------------------------------------------- */
static void any_node_box(const rro_node* nodes, uint32_t addr, v3* lo, v3* hi) { node_box(&nodes[addr], lo, hi); }

int rro_check_consistency(const rro_node* nodes, uint32_t n) /* test/test_vk/hlbvh_test.h:69-100 */
{
    size_t total = 2 * (size_t)n - 1;
    uint32_t* queue  = (uint32_t*)malloc(sizeof(uint32_t) * total);
    uint32_t* qpar   = (uint32_t*)malloc(sizeof(uint32_t) * total);
    uint8_t*  seen   = (uint8_t*)calloc(total, 1);
    uint8_t*  prim   = (uint8_t*)calloc(n, 1);
    size_t head = 0, tail = 0;
    int ok = 1;
    queue[tail] = 0; qpar[tail] = RRO_INVALID; ++tail;
    while (head < tail && ok)
    {
        uint32_t addr = queue[head], parent = qpar[head];
        ++head;
        if (addr >= total || seen[addr]) { ok = 0; break; }
        seen[addr] = 1;
        const rro_node* nd = &nodes[addr];
        if (nd->parent != parent) { ok = 0; break; }
        if (is_internal(nd))
        {
            if (nd->child0 >= total || nd->child1 >= total || tail + 2 > total) { ok = 0; break; }
            v3 lo, hi, l0, h0, l1, h1;
            any_node_box(nodes, addr, &lo, &hi);
            any_node_box(nodes, nd->child0, &l0, &h0);
            any_node_box(nodes, nd->child1, &l1, &h1);
            const float e = -1e-8f; /* common.h:195-201 Includes() */
            if (l0.x - lo.x < e || l0.y - lo.y < e || l0.z - lo.z < e || hi.x - h0.x < e || hi.y - h0.y < e || hi.z - h0.z < e) ok = 0;
            if (l1.x - lo.x < e || l1.y - lo.y < e || l1.z - lo.z < e || hi.x - h1.x < e || hi.y - h1.y < e || hi.z - h1.z < e) ok = 0;
            /* stronger than the reference: stored child boxes must equal the child's own box exactly */
            if (memcmp(&l0, nd->aabb0_min_or_v0, 12) || memcmp(&h0, nd->aabb0_max_or_v1, 12)) ok = 0;
            if (memcmp(&l1, nd->aabb1_min_or_v2, 12) || memcmp(&h1, nd->aabb1_max_or_v3, 12)) ok = 0;
            queue[tail] = nd->child0; qpar[tail] = addr; ++tail;
            queue[tail] = nd->child1; qpar[tail] = addr; ++tail;
        }
        else
        {
            if (nd->child1 >= n || prim[nd->child1]) { ok = 0; break; }
            prim[nd->child1] = 1;
        }
    }
    if (ok && tail != total) ok = 0;
    free(queue); free(qpar); free(seen); free(prim);
    return ok;
}

float rro_sah(const rro_node* nodes, uint32_t n) /* bvh_analyzer/bvh.h:187-223 shape: sum(area)/root area */
{
    if (n < 2) return 1.0f;
    v3 lo, hi;
    node_box(&nodes[0], &lo, &hi);
    v3 e = v3_sub(hi, lo);
    double root = 2.0 * ((double)e.x * e.y + (double)e.x * e.z + (double)e.y * e.z);
    double acc = 0.0;
    for (size_t i = 0; i < 2 * (size_t)n - 1; ++i)
    {
        node_box(&nodes[i], &lo, &hi);
        e = v3_sub(hi, lo);
        acc += 2.0 * ((double)e.x * e.y + (double)e.x * e.z + (double)e.y * e.z);
    }
    return (float)(acc / root);
}

uint32_t rro_depth(const rro_node* nodes, uint32_t n)
{
    uint32_t best = 0;
    for (uint32_t j = 0; j < n; ++j)
    {
        uint32_t d = 1, a = nodes[n - 1 + j].parent;
        while (a != RRO_INVALID) { ++d; a = nodes[a].parent; }
        if (d > best) best = d;
    }
    return best;
}

/* ------------------------------------------------------------------------------------------------
 * B8/B9 traversal -- common.h:103-209, isect.comp:88-246, isect_2l.comp:105-323
 * ---------------------------------------------------------------------------------------------- */
static inline v3 safe_invdir(v3 d) /* common.h:166-183 */
{
    const float e = 1e-5f;
    v3 r;
    r.x = 1.0f / (fabsf(d.x) > e ? d.x : (d.x < 0.0f ? -e : e));
    r.y = 1.0f / (fabsf(d.y) > e ? d.y : (d.y < 0.0f ? -e : e));
    r.z = 1.0f / (fabsf(d.z) > e ? d.z : (d.z < 0.0f ? -e : e));
    return r;
}
/* fast_intersect_aabb common.h:150-164 (fma is explicit in the shader) */
static inline void slab(const float* pmin, const float* pmax, v3 inv, v3 oxinv, float t_max, float t_min, float* t0,
                        float* t1)
{
    float fx = fmaf(pmax[0], inv.x, oxinv.x), fy = fmaf(pmax[1], inv.y, oxinv.y), fz = fmaf(pmax[2], inv.z, oxinv.z);
    float nx = fmaf(pmin[0], inv.x, oxinv.x), ny = fmaf(pmin[1], inv.y, oxinv.y), nz = fmaf(pmin[2], inv.z, oxinv.z);
    float ax = rr_max(fx, nx), ay = rr_max(fy, ny), az = rr_max(fz, nz);
    float ix = rr_min(fx, nx), iy = rr_min(fy, ny), iz = rr_min(fz, nz);
    *t1 = rr_min(rr_min(az, rr_min(ax, ay)), t_max); /* mymin3(a,b,c)=min(c,min(a,b)) */
    *t0 = rr_max(rr_max(iz, rr_max(ix, iy)), t_min);
}
/* fast_intersect_triangle common.h:103-137; returns 1 and *t on acceptance by the shader's bounds test */
static inline int tri_test(v3 o, v3 d, float min_t, const float* pv0, const float* pv1, const float* pv2, float t_max,
                           float* t)
{
    v3 v0 = v3_ld(pv0), e1 = v3_sub(v3_ld(pv1), v0), e2 = v3_sub(v3_ld(pv2), v0);
    v3 s1 = v3_cross(d, e2);
    float denom = v3_dot(s1, e1);
    if (denom == 0.0f) return 0;
    float invd = 1.0f / denom;
    v3 r = v3_sub(o, v0);
    float b1 = v3_dot(r, s1) * invd;
    v3 s2 = v3_cross(r, e1);
    float b2 = v3_dot(d, s2) * invd;
    float tt = v3_dot(e2, s2) * invd;
    if (b1 < 0.0f || b1 > 1.0f || b2 < 0.0f || b1 + b2 > 1.0f || tt < min_t || tt > t_max) return 0;
    *t = tt;
    return 1;
}
/* calculate_barycentrics common.h:187-209 */
static inline void barycentrics(v3 p, const float* pv0, const float* pv1, const float* pv2, float uv[2])
{
    v3 v0 = v3_ld(pv0), e1 = v3_sub(v3_ld(pv1), v0), e2 = v3_sub(v3_ld(pv2), v0), e = v3_sub(p, v0);
    float d00 = v3_dot(e1, e1), d01 = v3_dot(e1, e2), d11 = v3_dot(e2, e2), d20 = v3_dot(e, e1), d21 = v3_dot(e, e2);
    float denom = d00 * d11 - d01 * d01;
    if (denom == 0.0f) { uv[0] = 0.0f; uv[1] = 0.0f; return; }
    float inv = 1.0f / (d00 * d11 - d01 * d01);
    uv[0] = (d11 * d20 - d01 * d21) * inv;
    uv[1] = (d00 * d21 - d01 * d20) * inv;
}
static inline v3 ray_point(v3 o, v3 d, float t) { return v3_make(o.x + t * d.x, o.y + t * d.y, o.z + t * d.z); }

#define RRO_STACK 512

static void trace_one(const rro_node* bvh, const rro_ray* ray, int query, int output, int tie, void* hits, uint32_t i,
                      rro_ray_stats* st)
{
    v3 o = v3_ld(ray->o), d = v3_ld(ray->d);
    v3 inv = safe_invdir(d);
    v3 oxinv = v3_make(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
    float closest = ray->max_t;
    uint32_t closest_addr = RRO_INVALID, closest_prim = RRO_INVALID;
    uint32_t stack[RRO_STACK];
    uint32_t sp = 0, addr = 0, visited = 0, tested = 0, maxsp = 0;
    stack[sp++] = RRO_INVALID;
    while (addr != RRO_INVALID)
    {
        const rro_node* nd = &bvh[addr];
        ++visited;
        if (is_internal(nd))
        {
            float a0, a1, b0, b1;
            slab(nd->aabb0_min_or_v0, nd->aabb0_max_or_v1, inv, oxinv, closest, ray->min_t, &a0, &a1);
            slab(nd->aabb1_min_or_v2, nd->aabb1_max_or_v3, inv, oxinv, closest, ray->min_t, &b0, &b1);
            int t0 = a0 <= a1, t1 = b0 <= b1, c1first = t1 && (a0 > b0);
            if (t0 || t1)
            {
                uint32_t deferred;
                if (c1first || !t0) { addr = nd->child1; deferred = nd->child0; }
                else { addr = nd->child0; deferred = nd->child1; }
                if (t0 && t1)
                {
                    if (sp < RRO_STACK) stack[sp++] = deferred;
                    if (sp > maxsp) maxsp = sp;
                }
                continue;
            }
        }
        else
        {
            float t;
            ++tested;
            if (tri_test(o, d, ray->min_t, nd->aabb0_min_or_v0, nd->aabb0_max_or_v1, nd->aabb1_min_or_v2, closest, &t))
            {
                int accept = t < closest;
                if (!accept && tie == RRO_TIE_LOWEST_ID && query == RRO_QUERY_CLOSEST && t == closest &&
                    closest_addr != RRO_INVALID && nd->child1 < closest_prim)
                    accept = 1;
                if (accept)
                {
                    if (query == RRO_QUERY_ANY)
                    {
                        if (output == RRO_OUTPUT_FULL_HIT)
                        {
                            rro_hit* h = (rro_hit*)hits + i;
                            barycentrics(ray_point(o, d, t), nd->aabb0_min_or_v0, nd->aabb0_max_or_v1,
                                         nd->aabb1_min_or_v2, h->uv);
                            h->prim_id = nd->child1;
                            h->inst_id = 0;
                        }
                        else ((uint32_t*)hits)[i] = nd->child1; /* prim id, not 0 */
                        if (st) { st->nodes_visited = visited; st->triangles_tested = tested; st->max_stack = maxsp; st->t = t; }
                        return;
                    }
                    closest = t;
                    closest_addr = addr;
                    closest_prim = nd->child1;
                }
            }
        }
        addr = stack[--sp];
    }
    if (st) { st->nodes_visited = visited; st->triangles_tested = tested; st->max_stack = maxsp; st->t = closest; }
    if (closest_addr != RRO_INVALID)
    {
        const rro_node* nd = &bvh[closest_addr];
        if (output == RRO_OUTPUT_FULL_HIT)
        {
            rro_hit* h = (rro_hit*)hits + i;
            barycentrics(ray_point(o, d, closest), nd->aabb0_min_or_v0, nd->aabb0_max_or_v1, nd->aabb1_min_or_v2, h->uv);
            h->prim_id = nd->child1;
            h->inst_id = 0;
        }
        else ((uint32_t*)hits)[i] = nd->child1;
    }
    else
    {
        if (output == RRO_OUTPUT_FULL_HIT) ((rro_hit*)hits)[i].inst_id = RRO_INVALID; /* isect.comp:238-245 */
        else ((uint32_t*)hits)[i] = RRO_INVALID;
    }
}

void rro_trace(const rro_node* bvh, const rro_ray* rays, uint32_t count, int query, int output, int tie, void* hits,
               rro_ray_stats* stats)
{
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < (int64_t)count; ++i)
        trace_one(bvh, &rays[i], query, output, tie, hits, (uint32_t)i, stats ? &stats[i] : NULL);
}

static inline void xform_ray(const float* m, v3* o, v3* d) /* transform_ray common.h:310-324 */
{
    v3 oo = *o, dd = *d;
    o->x = v3_dot(v3_make(m[0], m[1], m[2]), oo) + m[3];
    o->y = v3_dot(v3_make(m[4], m[5], m[6]), oo) + m[7];
    o->z = v3_dot(v3_make(m[8], m[9], m[10]), oo) + m[11];
    d->x = v3_dot(v3_make(m[0], m[1], m[2]), dd);
    d->y = v3_dot(v3_make(m[4], m[5], m[6]), dd);
    d->z = v3_dot(v3_make(m[8], m[9], m[10]), dd);
}

static void trace_one_2l(const rro_node* tlas, const float* xf, const rro_node* const* blas, const rro_ray* ray,
                         int query, int output, int tie, void* hits, uint32_t i, rro_ray_stats* st)
{
    v3 o = v3_ld(ray->o), d = v3_ld(ray->d);
    v3 inv = safe_invdir(d);
    v3 oxinv = v3_make(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
    float closest = ray->max_t;
    uint32_t closest_addr = RRO_INVALID, closest_inst = RRO_INVALID, closest_prim = RRO_INVALID;
    uint32_t cur_inst = RRO_INVALID;
    uint32_t stack[RRO_STACK];
    uint32_t sp = 0, addr = 0, visited = 0, tested = 0, maxsp = 0;
    stack[sp++] = RRO_INVALID;
    while (addr != RRO_INVALID)
    {
        const rro_node* nd = (cur_inst == RRO_INVALID) ? &tlas[addr] : &blas[cur_inst][addr];
        ++visited;
        if (is_internal(nd))
        {
            float a0, a1, b0, b1;
            slab(nd->aabb0_min_or_v0, nd->aabb0_max_or_v1, inv, oxinv, closest, ray->min_t, &a0, &a1);
            slab(nd->aabb1_min_or_v2, nd->aabb1_max_or_v3, inv, oxinv, closest, ray->min_t, &b0, &b1);
            int t0 = a0 <= a1, t1 = b0 <= b1, c1first = t1 && (a0 > b0);
            if (t0 || t1)
            {
                uint32_t deferred;
                if (c1first || !t0) { addr = nd->child1; deferred = nd->child0; }
                else { addr = nd->child0; deferred = nd->child1; }
                if (t0 && t1)
                {
                    if (sp < RRO_STACK) stack[sp++] = deferred;
                    if (sp > maxsp) maxsp = sp;
                }
                continue;
            }
        }
        else if (cur_inst == RRO_INVALID)
        { /* TLAS leaf: enter the instance, isect_2l.comp:231-245 */
            cur_inst = nd->child1;
            xform_ray(xf + 24 * (size_t)cur_inst, &o, &d);
            inv = safe_invdir(d);
            oxinv = v3_make(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
            if (sp < RRO_STACK) stack[sp++] = RRO_SENTINEL;
            if (sp > maxsp) maxsp = sp;
            addr = 0;
            continue;
        }
        else
        {
            float t;
            ++tested;
            if (tri_test(o, d, ray->min_t, nd->aabb0_min_or_v0, nd->aabb0_max_or_v1, nd->aabb1_min_or_v2, closest, &t))
            {
                int accept = t < closest;
                if (!accept && tie == RRO_TIE_LOWEST_ID && query == RRO_QUERY_CLOSEST && t == closest &&
                    closest_addr != RRO_INVALID &&
                    (cur_inst < closest_inst || (cur_inst == closest_inst && nd->child1 < closest_prim)))
                    accept = 1;
                if (accept)
                {
                    if (query == RRO_QUERY_ANY)
                    {
                        if (output == RRO_OUTPUT_FULL_HIT)
                        {
                            rro_hit* h = (rro_hit*)hits + i;
                            barycentrics(ray_point(o, d, t), nd->aabb0_min_or_v0, nd->aabb0_max_or_v1,
                                         nd->aabb1_min_or_v2, h->uv);
                            h->prim_id = nd->child1;
                            h->inst_id = cur_inst;
                        }
                        else ((uint32_t*)hits)[i] = cur_inst;
                        if (st) { st->nodes_visited = visited; st->triangles_tested = tested; st->max_stack = maxsp; st->t = t; }
                        return;
                    }
                    closest = t;
                    closest_addr = addr;
                    closest_prim = nd->child1;
                    closest_inst = cur_inst;
                }
            }
        }
        addr = stack[--sp];
        if (addr == RRO_SENTINEL)
        { /* back to the top level: restore the original ray, isect_2l.comp:279-287 */
            cur_inst = RRO_INVALID;
            o = v3_ld(ray->o); d = v3_ld(ray->d);
            inv = safe_invdir(d);
            oxinv = v3_make(-o.x * inv.x, -o.y * inv.y, -o.z * inv.z);
            addr = stack[--sp];
        }
    }
    if (st) { st->nodes_visited = visited; st->triangles_tested = tested; st->max_stack = maxsp; st->t = closest; }
    if (closest_addr != RRO_INVALID)
    {
        if (output == RRO_OUTPUT_FULL_HIT)
        {
            const rro_node* nd = &blas[closest_inst][closest_addr];
            o = v3_ld(ray->o); d = v3_ld(ray->d);
            xform_ray(xf + 24 * (size_t)closest_inst, &o, &d);
            rro_hit* h = (rro_hit*)hits + i;
            barycentrics(ray_point(o, d, closest), nd->aabb0_min_or_v0, nd->aabb0_max_or_v1, nd->aabb1_min_or_v2, h->uv);
            h->prim_id = closest_prim;
            h->inst_id = closest_inst;
        }
        else ((uint32_t*)hits)[i] = closest_inst;
    }
    else
    {
        if (output == RRO_OUTPUT_FULL_HIT) ((rro_hit*)hits)[i].inst_id = RRO_INVALID;
        else ((uint32_t*)hits)[i] = RRO_INVALID;
    }
}

void rro_trace_2l(const rro_node* tlas, const float* xf, const rro_node* const* blas, const rro_ray* rays, uint32_t count,
                  int query, int output, int tie, void* hits, rro_ray_stats* stats)
{
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t i = 0; i < (int64_t)count; ++i)
        trace_one_2l(tlas, xf, blas, &rays[i], query, output, tie, hits, (uint32_t)i, stats ? &stats[i] : NULL);
}

void rro_brute_force(const float* verts, uint32_t sf, const uint32_t* idx, uint32_t n, const rro_ray* rays, uint32_t count,
                     rro_hit* hits, float* out_t)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)count; ++i)
    {
        const rro_ray* ray = &rays[i];
        v3 o = v3_ld(ray->o), d = v3_ld(ray->d);
        float closest = ray->max_t;
        uint32_t best = RRO_INVALID;
        for (uint32_t k = 0; k < n; ++k)
        {
            const float* a = verts + (size_t)idx[3 * (size_t)k + 0] * sf;
            const float* b = verts + (size_t)idx[3 * (size_t)k + 1] * sf;
            const float* c = verts + (size_t)idx[3 * (size_t)k + 2] * sf;
            float t;
            /* same acceptance as the traversal: hits at exactly max_t are never accepted first */
            if (tri_test(o, d, ray->min_t, a, b, c, closest, &t) && (t < closest)) { closest = t; best = k; }
        }
        if (best != RRO_INVALID)
        {
            const float* a = verts + (size_t)idx[3 * (size_t)best + 0] * sf;
            const float* b = verts + (size_t)idx[3 * (size_t)best + 1] * sf;
            const float* c = verts + (size_t)idx[3 * (size_t)best + 2] * sf;
            barycentrics(ray_point(o, d, closest), a, b, c, hits[i].uv);
            hits[i].prim_id = best;
            hits[i].inst_id = 0;
        }
        else { hits[i].uv[0] = hits[i].uv[1] = 0.0f; hits[i].prim_id = RRO_INVALID; hits[i].inst_id = RRO_INVALID; }
        if (out_t) out_t[i] = closest;
    }
}
