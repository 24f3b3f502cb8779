/*
 * rr_oracle.h -- CPU restatement of the RadeonRays 4.1 Vulkan kernels for the HLBVH build /
 * refit / TLAS / closest+any-hit traversal path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the shipped library
 * (radeonrays_sdk_b200/csrc) never links, calls or falls back to anything in oracle/.
 *
 * Parity status: PINNED at value level to reference code run here, for everything the reference's one compilable
 * component (bvh_analyzer, built from /root/reference into oracle/_ref by oracle/Makefile) computes: the triangle test
 * (triangle.h:34-70: closest t bit for bit, hit/miss masks, primitive ids on C1 at 1024 x 1024 and on C2 samples), its
 * BVH2 tracer (bvh.h:226-319: ids on single-intersection and clipped rays, uv), its validator and SAH (bvh.h:130-223)
 * and the stock binary on the 7-line config -- tests/test_reference_pin_cpu.py, tests/test_gpu_reference_pin.py.
 * The reference's tests hold NO golden vectors for the rest: Morton codes, sorted order, Karras
 * topology, treelet output and TLAS are pinned by the structural assertions the reference's tests make (BVH
 * consistency, sorted == std::sort, SAH_after <= SAH_before), by bvh_analyzer accepting the dump, and otherwise by the
 * shader text each function cites; those rows stay "unpinned by reference values" because the Vulkan library cannot
 * be compiled here (no Vulkan SDK, glslang or spdlog).
 *
 * Arithmetic contract: IEEE-754 binary32, round to nearest even, no FMA
 * contraction except where the GLSL writes fma() (compile with -ffp-contract=off).
 */
#ifndef RR_ORACLE_H
#define RR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RRO_INVALID 0xFFFFFFFFu
#define RRO_SENTINEL 0xFFFFFFFEu /* isect_2l.comp: RR_TOP_LEVEL_SENTINEL */

/* 64-byte BVH2 node, vlk/kernels/bvh2.h:25-35 (host mirror bvh_analyzer/transform.h:31-41). */
typedef struct
{
    float    aabb0_min_or_v0[3];
    uint32_t child0;
    float    aabb0_max_or_v1[3];
    uint32_t child1;
    float    aabb1_min_or_v2[3];
    uint32_t parent;
    float    aabb1_max_or_v3[3];
    uint32_t update;
} rro_node;

typedef struct { float o[3]; float min_t; float d[3]; float max_t; } rro_ray; /* radeonrays.h:147-153 */
typedef struct { float uv[2]; uint32_t inst_id; uint32_t prim_id; } rro_hit;  /* radeonrays.h:157-162 */

enum { RRO_QUERY_CLOSEST = 0, RRO_QUERY_ANY = 1 };
enum { RRO_OUTPUT_FULL_HIT = 0, RRO_OUTPUT_INSTANCE_ID = 1 };
enum { RRO_TIE_LOWEST_ID = 0, RRO_TIE_FIRST_FOUND = 1 };

/* per-ray statistics (optional) */
typedef struct { uint32_t nodes_visited; uint32_t triangles_tested; uint32_t max_stack; float t; } rro_ray_stats;

/* ---- build ------------------------------------------------------------------------------------ */
void rro_scene_aabb(const float* verts, uint32_t stride_floats, const uint32_t* idx, uint32_t n,
                    float smin[3], float smax[3]);
void rro_morton_codes(const float* verts, uint32_t stride_floats, const uint32_t* idx, uint32_t n,
                      const float smin[3], const float smax[3], uint32_t* codes);
void rro_sort_pairs(const uint32_t* keys, const uint32_t* vals, uint32_t n, uint32_t* out_keys, uint32_t* out_vals);
void rro_emit_hierarchy(const uint32_t* sorted_codes, const uint32_t* sorted_refs, uint32_t n, rro_node* nodes);
void rro_fit_mesh(rro_node* nodes, uint32_t n, const float* verts, uint32_t stride_floats, const uint32_t* idx);
/* 3 rounds of 7-leaf treelet restructuring, min_prims 64/128/256 (restructure_hlbvh.cpp:151-277). */
void rro_restructure(rro_node* nodes, uint32_t n);
void rro_restructure_round(rro_node* nodes, uint32_t n, uint32_t min_prims);
/* whole BLAS build; codes/refs (sorted) are optional outputs (may be NULL). */
void rro_build_blas(const float* verts, uint32_t stride_floats, const uint32_t* idx, uint32_t n, int restructure,
                    rro_node* nodes, uint32_t* sorted_codes, uint32_t* sorted_refs);

/* 63-bit Morton extension (21 bits per axis; dx/kernels/build_hlbvh_fallback.hlsl:95-108 is the reference's only trace of it). */
void rro_morton_codes63(const float* verts, uint32_t stride_floats, const uint32_t* idx, uint32_t n,
                        const float smin[3], const float smax[3], uint64_t* codes);
void rro_emit_hierarchy64(const uint64_t* sorted_codes, const uint32_t* sorted_refs, uint32_t n, rro_node* nodes);
void rro_build_blas63(const float* verts, uint32_t stride_floats, const uint32_t* idx, uint32_t n, int restructure,
                      rro_node* nodes, uint64_t* sorted_codes, uint32_t* sorted_refs);

/* TLAS over instances. blas_roots[i] -> node 0 of instance i's BLAS, blas_tris[i] = its triangle count;
 * transforms = 12 floats per instance (row-major 3x4).  out_transforms = 2n x 12 floats
 * ([2i]=inverse, [2i+1]=forward).  reference_corner_quirk=1 reproduces the reference's
 * transform_aabb corner set (common.h:282-289: one corner duplicated, one missing). */
void rro_build_tlas(const rro_node* const* blas_roots, const uint32_t* blas_tris, const float* transforms,
                    uint32_t n, int reference_corner_quirk, rro_node* nodes, float* out_transforms);

/* ---- validation helpers ----------------------------------------------------------------------- */
/* BFS parent/child/box-containment check (test/test_vk/hlbvh_test.h:69-100); returns 1 if consistent.
 * Also verifies every leaf/internal node is reached exactly once. */
int   rro_check_consistency(const rro_node* nodes, uint32_t n);
float rro_sah(const rro_node* nodes, uint32_t n);
uint32_t rro_depth(const rro_node* nodes, uint32_t n);

/* ---- trace ------------------------------------------------------------------------------------ */
/* One-level traversal (isect.comp:88-246).  hits: rro_hit[count] (FULL_HIT) or uint32_t[count].
 * Miss writes only inst_id (FULL_HIT) -- other fields keep their previous contents. */
void rro_trace(const rro_node* bvh, const rro_ray* rays, uint32_t count, int query, int output, int tie_rule,
               void* hits, rro_ray_stats* stats);
/* Two-level traversal (isect_2l.comp:136-323). transforms = out_transforms of rro_build_tlas. */
void rro_trace_2l(const rro_node* tlas, const float* transforms, const rro_node* const* blas_roots,
                  const rro_ray* rays, uint32_t count, int query, int output, int tie_rule, void* hits,
                  rro_ray_stats* stats);
/* Order-independent reference: test every triangle (same triangle arithmetic), closest by (t, prim). */
void rro_brute_force(const float* verts, uint32_t stride_floats, const uint32_t* idx, uint32_t n,
                     const rro_ray* rays, uint32_t count, rro_hit* hits, float* out_t);

int rro_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
