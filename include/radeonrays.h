/*
 * radeonrays.h -- public C ABI of the B200-native ray-intersection engine.
 *
 * This header declares exactly the rr* entry points, enums and POD structs that RadeonRays 4.1
 * exposes (reference: src/core/include/radeonrays.h:40-473) so that a client compiled against the
 * reference header links and runs unchanged against libradeonrays_b200.so.  Values and struct
 * layouts are binary-compatible; the only addition is the RR_API_CUDA enumerator (the reference
 * has RR_API_DX=1 and RR_API_VK=2, radeonrays.h:70-74).  CUDA interop lives in radeonrays_cuda.h.
 *
 * Execution model (reference: radeonrays.cpp:242-299,504-529): rrCmd* calls only RECORD work into
 * a command stream; nothing executes until rrSumbitCommandStream (sic -- the typo is ABI).
 * Threading (reference radeonrays.h:267-270): different contexts may be used from different
 * threads; calls on one context must be serialised by the caller.
 */
#ifndef RADEONRAYS_H
#define RADEONRAYS_H

#include <stddef.h>
#include <stdint.h>

#define RR_API_MAJOR_VERSION 0x000000
#define RR_API_MINOR_VERSION 0x000001
#define RR_API_PATCH_VERSION 0x000001
#define RR_API_VERSION RR_API_MAJOR_VERSION * 1000000 + RR_API_MINOR_VERSION * 1000 + RR_API_PATCH_VERSION

#if defined(_WIN32)
#  if defined(RR_EXPORT_API)
#    define RR_API __declspec(dllexport)
#  else
#    define RR_API __declspec(dllimport)
#  endif
#else
#  define RR_API __attribute__((visibility("default")))
#endif

/* ---- opaque handles (reference radeonrays.h:40-50) ------------------------------------------ */
typedef struct _RRDevicePtr*     RRDevicePtr;
typedef struct _RRContext*       RRContext;
typedef struct _RREvent*         RREvent;
typedef struct _RRCommandStream* RRCommandStream;
typedef uint32_t                 RRBuildFlags;
typedef uint32_t                 RRRayMask;

/* Miss marker written to RRHit::inst_id / the packed id output (reference :52-55, spelling is ABI). */
enum { RR_INAVLID_VALUE = ~0u };

/* ---- enums ------------------------------------------------------------------------------------ */
typedef enum {                              /* reference :57-68 */
    RR_SUCCESS                    = 0,
    RR_ERROR_NOT_IMPLEMENTED      = 1,
    RR_ERROR_INTERNAL             = 2,
    RR_ERROR_OUT_OF_HOST_MEMORY   = 3,
    RR_ERROR_OUT_OF_DEVICE_MEMORY = 4,
    RR_ERROR_INVALID_API_VERSION  = 5,
    RR_ERROR_INVALID_PARAMETER    = 6,
    RR_ERROR_UNSUPPORTED_API      = 7,
    RR_ERROR_UNSUPPORTED_INTEROP  = 8
} RRError;

typedef enum {                              /* reference :70-74, plus the CUDA backend */
    RR_API_DX   = 1,
    RR_API_VK   = 2,
    RR_API_CUDA = 3
} RRApi;

typedef enum {                              /* reference :76-83 */
    RR_LOG_LEVEL_DEBUG = 1,
    RR_LOG_LEVEL_INFO  = 2,
    RR_LOG_LEVEL_WARN  = 3,
    RR_LOG_LEVEL_ERROR = 4,
    RR_LOG_LEVEL_OFF   = 5
} RRLogLevel;

typedef enum {                              /* reference :90-94 */
    RR_BUILD_OPERATION_BUILD  = 1,          /* full HLBVH build                                  */
    RR_BUILD_OPERATION_UPDATE = 2           /* refit: new vertex positions, topology untouched   */
} RRBuildOperation;

typedef enum {                              /* reference :101-105 */
    RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD = 1,   /* skip treelet restructuring                    */
    RR_BUILD_FLAG_BITS_ALLOW_UPDATE      = 2    /* accepted, never needed: update always allowed */
} RRBuildFlagBits;

typedef enum { RR_PRIMITIVE_TYPE_TRIANGLE_MESH, RR_PRIMITIVE_TYPE_AABB_LIST } RRPrimitiveType; /* :113-117 */
typedef enum { RR_INDEX_TYPE_UINT32, RR_INDEX_TYPE_UINT16 } RRIndexType;                       /* :122-126 */
typedef enum { RR_INTERSECT_QUERY_CLOSEST = 0, RR_INTERSECT_QUERY_ANY = 1 } RRIntersectQuery;  /* :131-135 */
typedef enum {                                                                                 /* :139-143 */
    RR_INTERSECT_QUERY_OUTPUT_FULL_HIT,     /* RRHit[ray_count]                                  */
    RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID   /* uint32_t[ray_count]                               */
} RRIntersectQueryOutput;

/* ---- POD structs (layouts are ABI) ------------------------------------------------------------ */
typedef struct { float origin[3]; float min_t; float direction[3]; float max_t; } RRRay;   /* 32 B, :147-153 */
typedef struct { float uv[2]; uint32_t inst_id; uint32_t prim_id; } RRHit;                 /* 16 B, :157-162 */

typedef struct {                            /* reference :166-173 */
    RRBuildFlags build_flags;
    void*        backend_specific_info;
} RRBuildOptions;

typedef struct {                            /* reference :182-192 */
    RRDevicePtr vertices;                   /* xyz float positions                               */
    uint32_t    vertex_count;
    uint32_t    vertex_stride;              /* bytes between vertices, multiple of 4             */
    RRDevicePtr triangle_indices;           /* 3 indices per triangle                            */
    uint32_t    triangle_count;
    RRIndexType index_type;
} RRTriangleMeshPrimitive;

typedef struct {                            /* reference :201-206 */
    RRDevicePtr aabbs;
    uint32_t    aabb_count;
    uint32_t    aabb_stride;
} RRAABBListPrimitive;

typedef struct {                            /* reference :213-226 */
    RRPrimitiveType primitive_type;
    uint32_t        primitive_count;
    union {
        RRTriangleMeshPrimitive* triangle_mesh_primitives;
        RRAABBListPrimitive*     aabb_primitives;
    };
} RRGeometryBuildInput;

typedef struct { RRDevicePtr geometry; float transform[3][4]; } RRInstance;                /* :228-232 */
typedef struct { const RRInstance* instances; uint32_t instance_count; } RRSceneBuildInput; /* :246-252 */

typedef struct {                            /* reference :254-259 */
    size_t temporary_build_buffer_size;
    size_t temporary_update_buffer_size;
    size_t result_buffer_size;
} RRMemoryRequirements;

#ifdef __cplusplus
extern "C" {
#endif

/* Context / logging (reference :276-304). rrCreateContext(v, RR_API_CUDA, &ctx) = device 0 and a
 * library-owned stream; other RRApi values return RR_ERROR_UNSUPPORTED_API. */
RR_API RRError rrCreateContext(uint32_t api_version, RRApi api, RRContext* context);
RR_API RRError rrDestroyContext(RRContext context);
RR_API RRError rrSetLogLevel(RRLogLevel log_level);
RR_API RRError rrSetLogFile(char const* filename);

/* Bottom-level build / refit (reference :322-341). */
RR_API RRError rrCmdBuildGeometry(RRContext context, RRBuildOperation build_operation,
                                  const RRGeometryBuildInput* build_input, const RRBuildOptions* build_options,
                                  RRDevicePtr temporary_buffer, RRDevicePtr geometry_buffer,
                                  RRCommandStream command_stream);
RR_API RRError rrGetGeometryBuildMemoryRequirements(RRContext context, const RRGeometryBuildInput* build_input,
                                                    const RRBuildOptions* build_options,
                                                    RRMemoryRequirements* memory_requirements);

/* Top-level (instance) build (reference :357-375). */
RR_API RRError rrCmdBuildScene(RRContext context, const RRSceneBuildInput* build_input,
                               const RRBuildOptions* build_options, RRDevicePtr temporary_buffer,
                               RRDevicePtr scene_buffer, RRCommandStream command_stream);
RR_API RRError rrGetSceneBuildMemoryRequirements(RRContext context, const RRSceneBuildInput* build_input,
                                                 const RRBuildOptions* build_options,
                                                 RRMemoryRequirements* memory_requirements);

/* Trace (reference :391-409). scene_buffer may be a geometry (one level) or a scene (two level). */
RR_API RRError rrCmdIntersect(RRContext context, RRDevicePtr scene_buffer, RRIntersectQuery query,
                              RRDevicePtr rays, uint32_t ray_count, RRDevicePtr indirect_ray_count,
                              RRIntersectQueryOutput query_output, RRDevicePtr hits, RRDevicePtr scratch,
                              RRCommandStream command_stream);
RR_API RRError rrGetTraceMemoryRequirements(RRContext context, uint32_t ray_count, size_t* scratch_size);

/* Command streams, events, device pointers (reference :418-473). */
RR_API RRError rrAllocateCommandStream(RRContext context, RRCommandStream* command_stream);
RR_API RRError rrReleaseCommandStream(RRContext context, RRCommandStream command_stream);
RR_API RRError rrSumbitCommandStream(RRContext context, RRCommandStream command_stream, RREvent wait_event,
                                     RREvent* out_event);
RR_API RRError rrReleaseEvent(RRContext context, RREvent event);
RR_API RRError rrWaitEvent(RRContext context, RREvent event);
RR_API RRError rrReleaseDevicePtr(RRContext context, RRDevicePtr ptr);
RR_API RRError rrReleaseExternalCommandStream(RRContext context, RRCommandStream command_stream);

#ifdef __cplusplus
}
#endif
#endif /* RADEONRAYS_H */
