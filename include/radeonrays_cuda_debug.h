/*
 * radeonrays_cuda_debug.h -- test / benchmark hooks of the CUDA backend.  No reference counterpart in
 * the public headers: RadeonRays' own tests reach the same internals by compiling library sources into
 * the test binary (test/test_vk/CMakeLists.txt:21-47: algos_test.h drives RadixSortKeyValue directly,
 * hlbvh_test.h drives BuildHlBvh/RestructureHlBvh/UpdateHlBvh).  Everything here launches the same CUDA
 * kernels the rr* entry points use.
 */
#ifndef RADEONRAYS_CUDA_DEBUG_H
#define RADEONRAYS_CUDA_DEBUG_H

#include "radeonrays_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Byte offsets, inside the temporary build buffer of a geometry with `triangle_count` triangles, of the
 * intermediate arrays the build leaves behind: scene AABB (uint[8], ordered encoding, reference
 * lbvh_init_mesh.comp:61-77), unsorted Morton codes, sorted codes (u32[N] each; with RR_CUDA_OPTION_MORTON_BITS = 63 the sorted codes
 * are u64[N] and morton_codes_offset holds their unsorted low words).  The sorted primitive refs (u32[N]) are kept in
 * the GEOMETRY buffer, after the node array, where RR_BUILD_OPERATION_UPDATE reads them: sorted_refs_offset is an offset
 * into the geometry buffer. */
typedef struct
{
    size_t scene_aabb_offset;
    size_t morton_codes_offset;
    size_t sorted_codes_offset;
    size_t sorted_refs_offset;     /* into the geometry buffer (not the temporary buffer) */
    size_t sort_tmp_values_offset; /* ping-pong value buffer of the sort, dead once the sort has finished */
} RRCudaBuildScratchLayout;
RR_API RRError rrCudaDebugGetBuildScratchLayout(RRContext context, uint32_t triangle_count, RRCudaBuildScratchLayout* layout);

/* Stable key-value radix sort (the analogue of algos_test.h:322-390 SortTest).  All pointers are device
 * addresses; values_in may be NULL (values become 0..n-1).  Executes immediately on the context stream and
 * returns after it completes. */
RR_API RRError rrCudaDebugSortPairs(RRContext context, void* keys_in, void* values_in, void* keys_out, void* values_out,
                                    uint32_t count);

/* Treelet restructuring alone on an already built geometry (hlbvh_test.h:350-420 RestructureTest). */
RR_API RRError rrCudaDebugRestructure(RRContext context, RRDevicePtr geometry, uint32_t triangle_count, RRDevicePtr temporary_buffer);

/* Byte offsets inside a scene buffer with `instance_count` instances: nodes, 64-byte instance records
 * (inverse transform rows + BLAS address), forward transforms (3 x float4 per instance). */
typedef struct
{
    size_t nodes_offset;
    size_t records_offset;
    size_t forward_transforms_offset;
} RRCudaSceneLayout;
RR_API RRError rrCudaDebugGetSceneLayout(RRContext context, uint32_t instance_count, RRCudaSceneLayout* layout);

#ifdef __cplusplus
}
#endif
#endif
