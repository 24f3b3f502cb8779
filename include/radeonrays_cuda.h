/*
 * radeonrays_cuda.h -- CUDA interop for the rr* API (the analogue of the reference's
 * src/core/include/radeonrays_vlk.h:38-96 and radeonrays_dx.h:41-66).
 *
 * Signatures use only plain C types: a CUDA stream is passed as void* (cudaStream_t), device
 * memory as void* (a CUDA device address).  Calling any of these on a context that was not
 * created for RR_API_CUDA returns RR_ERROR_UNSUPPORTED_INTEROP (reference radeonrays.cpp:615-619).
 */
#ifndef RADEONRAYS_CUDA_H
#define RADEONRAYS_CUDA_H

#include "radeonrays.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Create a context on an existing device / stream (replaces rrCreateContextVk, radeonrays_vlk.h:38-43).
 * cuda_stream may be NULL: the library then owns a non-blocking stream on `device_ordinal`. */
RR_API RRError rrCreateContextCuda(uint32_t api_version, int device_ordinal, void* cuda_stream, RRContext* context);

/* Wrap client-owned device memory (replaces rrGetDevicePtrFromVkBuffer, radeonrays_vlk.h:53-56).
 * The client keeps ownership of the allocation; rrReleaseDevicePtr frees only the wrapper. */
RR_API RRError rrGetDevicePtrFromCudaPtr(RRContext context, void* device_memory, size_t offset, RRDevicePtr* device_ptr);

/* Wrap a client-owned cudaStream_t as a command stream (replaces rrGetCommandStreamFromVkCommandBuffer,
 * radeonrays_vlk.h:64-66).  Release with rrReleaseExternalCommandStream. */
RR_API RRError rrGetCommandStreamFromCudaStream(RRContext context, void* cuda_stream, RRCommandStream* command_stream);

/* Library-owned device buffers + host mapping (same names and meaning as radeonrays_vlk.h:74-96;
 * used by the reference's internal_resources_test.h:51-236).  Map returns a pinned host shadow that
 * holds the current device contents; Unmap writes the shadow back to the device and frees it. */
RR_API RRError rrAllocateDeviceBuffer(RRContext context, size_t size, RRDevicePtr* device_ptr);
RR_API RRError rrMapDevicePtr(RRContext context, RRDevicePtr device_ptr, void** mapping_ptr);
RR_API RRError rrUnmapDevicePtr(RRContext context, RRDevicePtr device_ptr, void** mapping_ptr);

/* ---- extensions without a reference counterpart (documented in DESIGN.md) -------------------- */

/* Raw device address behind an RRDevicePtr (base + offset). */
RR_API RRError rrGetCudaPtrFromDevicePtr(RRContext context, RRDevicePtr device_ptr, void** device_memory);

/* Backend options.
 * RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND: 0 (default) = equal-t closest hits resolve to the lowest
 *   (inst_id, prim_id), the rule BASELINE.json's north_star states; 1 = keep the first hit found in traversal
 *   order, which is what the reference shader does (vlk/kernels/isect.comp:179, `t < closest_t`).
 * RR_CUDA_OPTION_REFERENCE_TRANSFORM_AABB_QUIRK: 0 (default) = instance boxes use all 8 corners of the BLAS box;
 *   1 = reproduce the reference's corner set (vlk/kernels/common.h:282-289: pmax twice, (max,min,min) missing),
 *   which can under-estimate the box of a rotated instance. */
typedef enum
{
    RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND   = 1,
    RR_CUDA_OPTION_REFERENCE_TRANSFORM_AABB_QUIRK = 2,
    /* 1 = every rrCmdIntersect recorded afterwards first bins its rays on the device (direction octant, origin cell, coarse
     * direction: one key pass + a 3-pass radix sort) and traces them in that order; hits are bit-identical and stay in the
     * client's order.  For INCOHERENT batches (diffuse bounces); coherent batches are faster without.  Set it before
     * rrGetTraceMemoryRequirements: the scratch buffer grows by ~18 bytes per ray. */
    RR_CUDA_OPTION_SORT_RAYS                      = 4,
    /* 30 (default, the reference's codes) or 63: geometry builds recorded afterwards sort the triangles by 63-bit Morton codes
     * (21 bits per axis) -- far fewer equal codes on large or unevenly tessellated meshes.  An extension: the reference only keeps
     * a compiled-out 64-bit delta() (dx/kernels/build_hlbvh_fallback.hlsl:16,95-108); defined by oracle/rr_oracle.c
     * rro_build_blas63.  Set it before rrGetGeometryBuildMemoryRequirements (the temporary build buffer grows). */
    RR_CUDA_OPTION_MORTON_BITS                    = 5,
    /* Closest-hit packets are 8 x 8 tiles of the ray grid instead of 64 consecutive rays when the batch is an image in row
     * order (camera rays generated `for y, for x`, as the reference's tests do): 0 (default) = every rrCmdIntersect looks for
     * the row length on the device (one small kernel: where does the step between consecutive rays turn back?), 1 = never,
     * W >= 64 = the rows are W rays long, starting at ray 0.  Only the grouping of rays changes, never a hit: every ray keeps its
     * own (t, prim) minimum and writes it at its own index. */
    RR_CUDA_OPTION_RAY_GRID_WIDTH                 = 6,
    /* test hook: caps the hand-over lists of the staged refit (0 = library default) so that tests can drive them into overflow */
    RR_CUDA_OPTION_DEBUG_REFIT_LIST_CAPACITY      = 3
} RRCudaOption;
RR_API RRError rrCudaSetOption(RRContext context, RRCudaOption option, int value);

/* A scene buffer is self-describing (a header at its start holds a magic word, the instance count and the offsets of its parts;
 * rrCmdIntersect tells a scene from a geometry on the device), so it may be traced through any context of the same device, copied
 * with a plain device-to-device copy, or broadcast to another GPU.  Its instance records hold the device addresses of the
 * geometries they were built over: after copying a scene next to COPIES of its geometries (another GPU, another arena), record one
 * rebind per geometry -- every instance that pointed at `old_geometry_address` (the address the geometry had when the scene was
 * built, on whatever device) then points at `new_geometry`.  No reference counterpart (RadeonRays is single-GPU and keeps the
 * scene description in host-side state keyed by the buffer, vlk/intersector.cpp:86,263,289-290). */
RR_API RRError rrCudaCmdRebindSceneGeometry(RRContext context, RRDevicePtr scene_buffer, void* old_geometry_address,
                                            RRDevicePtr new_geometry, RRCommandStream command_stream);

/* Multi-GPU plumbing, one process per GPU (no reference counterpart, RadeonRays drives one device).
 * rrCudaExportDeviceMemory writes an opaque RR_CUDA_IPC_HANDLE_SIZE-byte handle for the allocation behind `device_ptr` and the
 * byte offset of `device_ptr` inside it; ship both to another process (any transport), whose rrCudaImportDeviceMemory maps the
 * allocation into ITS address space over NVLink / PCIe peer access and returns an RRDevicePtr at the same offset.  That pointer can
 * be passed as the `hits` argument of rrCmdIntersect: the traversal kernels then store every hit straight into the remote buffer
 * while they trace (a fused compute + gather; no staging copy, no separate collective).  Release it with rrReleaseDevicePtr. */
#define RR_CUDA_IPC_HANDLE_SIZE 64
RR_API RRError rrCudaExportDeviceMemory(RRContext context, RRDevicePtr device_ptr, void* handle_out, size_t* offset_out);
RR_API RRError rrCudaImportDeviceMemory(RRContext context, const void* handle, size_t offset, RRDevicePtr* device_ptr);

/* Number of CUDA kernels this context has launched so far (bench.py's gpu_launches evidence). */
RR_API RRError rrCudaGetLaunchCount(RRContext context, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* RADEONRAYS_CUDA_H */
