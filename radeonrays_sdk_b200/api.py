"""ctypes binding of libradeonrays_b200.so: the rr* C ABI (include/radeonrays.h) + CUDA interop.

Names, argument order and error codes are the reference's (src/core/include/radeonrays.h:276-473); the
thin `Context` helper below only removes ctypes boilerplate so tests read like test/test_vk/basic_test.h.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libradeonrays_b200.so")

RR_API_VERSION = 0 * 1000000 + 1 * 1000 + 1

# RRError
RR_SUCCESS, RR_ERROR_NOT_IMPLEMENTED, RR_ERROR_INTERNAL, RR_ERROR_OUT_OF_HOST_MEMORY = 0, 1, 2, 3
RR_ERROR_OUT_OF_DEVICE_MEMORY, RR_ERROR_INVALID_API_VERSION, RR_ERROR_INVALID_PARAMETER = 4, 5, 6
RR_ERROR_UNSUPPORTED_API, RR_ERROR_UNSUPPORTED_INTEROP = 7, 8
# RRApi
RR_API_DX, RR_API_VK, RR_API_CUDA = 1, 2, 3
# RRBuildOperation / flags
RR_BUILD_OPERATION_BUILD, RR_BUILD_OPERATION_UPDATE = 1, 2
RR_BUILD_FLAG_BITS_PREFER_FAST_BUILD, RR_BUILD_FLAG_BITS_ALLOW_UPDATE = 1, 2
RR_PRIMITIVE_TYPE_TRIANGLE_MESH, RR_PRIMITIVE_TYPE_AABB_LIST = 0, 1
RR_INDEX_TYPE_UINT32, RR_INDEX_TYPE_UINT16 = 0, 1
RR_INTERSECT_QUERY_CLOSEST, RR_INTERSECT_QUERY_ANY = 0, 1
RR_INTERSECT_QUERY_OUTPUT_FULL_HIT, RR_INTERSECT_QUERY_OUTPUT_INSTANCE_ID = 0, 1
RR_INVALID_VALUE = 0xFFFFFFFF
RR_CUDA_OPTION_CLOSEST_HIT_KEEP_FIRST_FOUND, RR_CUDA_OPTION_REFERENCE_TRANSFORM_AABB_QUIRK = 1, 2
RR_CUDA_OPTION_DEBUG_REFIT_LIST_CAPACITY = 3
RR_CUDA_OPTION_SORT_RAYS = 4
RR_CUDA_OPTION_MORTON_BITS = 5
RR_CUDA_OPTION_RAY_GRID_WIDTH = 6

_vp = C.c_void_p


class RRBuildOptions(C.Structure):
    _fields_ = [("build_flags", C.c_uint32), ("backend_specific_info", _vp)]


class RRTriangleMeshPrimitive(C.Structure):
    _fields_ = [("vertices", _vp), ("vertex_count", C.c_uint32), ("vertex_stride", C.c_uint32),
                ("triangle_indices", _vp), ("triangle_count", C.c_uint32), ("index_type", C.c_int)]


class RRGeometryBuildInput(C.Structure):
    _fields_ = [("primitive_type", C.c_int), ("primitive_count", C.c_uint32),
                ("triangle_mesh_primitives", C.POINTER(RRTriangleMeshPrimitive))]


class RRInstance(C.Structure):
    _fields_ = [("geometry", _vp), ("transform", (C.c_float * 4) * 3)]


class RRSceneBuildInput(C.Structure):
    _fields_ = [("instances", C.POINTER(RRInstance)), ("instance_count", C.c_uint32)]


class RRMemoryRequirements(C.Structure):
    _fields_ = [("temporary_build_buffer_size", C.c_size_t), ("temporary_update_buffer_size", C.c_size_t),
                ("result_buffer_size", C.c_size_t)]


class RRCudaBuildScratchLayout(C.Structure):
    _fields_ = [("scene_aabb_offset", C.c_size_t), ("morton_codes_offset", C.c_size_t),
                ("sorted_codes_offset", C.c_size_t), ("sorted_refs_offset", C.c_size_t), ("sort_tmp_values_offset", C.c_size_t)]


class RRCudaSceneLayout(C.Structure):
    _fields_ = [("nodes_offset", C.c_size_t), ("records_offset", C.c_size_t), ("forward_transforms_offset", C.c_size_t)]


# every symbol include/*.h declares: name -> argtypes (restype is RRError = int)
SIGNATURES = {
    "rrCreateContext": [C.c_uint32, C.c_int, C.POINTER(_vp)],
    "rrDestroyContext": [_vp],
    "rrSetLogLevel": [C.c_int],
    "rrSetLogFile": [C.c_char_p],
    "rrCmdBuildGeometry": [_vp, C.c_int, C.POINTER(RRGeometryBuildInput), C.POINTER(RRBuildOptions), _vp, _vp, _vp],
    "rrGetGeometryBuildMemoryRequirements": [_vp, C.POINTER(RRGeometryBuildInput), C.POINTER(RRBuildOptions),
                                             C.POINTER(RRMemoryRequirements)],
    "rrCmdBuildScene": [_vp, C.POINTER(RRSceneBuildInput), C.POINTER(RRBuildOptions), _vp, _vp, _vp],
    "rrGetSceneBuildMemoryRequirements": [_vp, C.POINTER(RRSceneBuildInput), C.POINTER(RRBuildOptions),
                                          C.POINTER(RRMemoryRequirements)],
    "rrCmdIntersect": [_vp, _vp, C.c_int, _vp, C.c_uint32, _vp, C.c_int, _vp, _vp, _vp],
    "rrGetTraceMemoryRequirements": [_vp, C.c_uint32, C.POINTER(C.c_size_t)],
    "rrAllocateCommandStream": [_vp, C.POINTER(_vp)],
    "rrReleaseCommandStream": [_vp, _vp],
    "rrSumbitCommandStream": [_vp, _vp, _vp, C.POINTER(_vp)],
    "rrReleaseEvent": [_vp, _vp],
    "rrWaitEvent": [_vp, _vp],
    "rrReleaseDevicePtr": [_vp, _vp],
    "rrReleaseExternalCommandStream": [_vp, _vp],
    # radeonrays_cuda.h
    "rrCreateContextCuda": [C.c_uint32, C.c_int, _vp, C.POINTER(_vp)],
    "rrGetDevicePtrFromCudaPtr": [_vp, _vp, C.c_size_t, C.POINTER(_vp)],
    "rrGetCommandStreamFromCudaStream": [_vp, _vp, C.POINTER(_vp)],
    "rrAllocateDeviceBuffer": [_vp, C.c_size_t, C.POINTER(_vp)],
    "rrMapDevicePtr": [_vp, _vp, C.POINTER(_vp)],
    "rrUnmapDevicePtr": [_vp, _vp, C.POINTER(_vp)],
    "rrGetCudaPtrFromDevicePtr": [_vp, _vp, C.POINTER(_vp)],
    "rrCudaSetOption": [_vp, C.c_int, C.c_int],
    "rrCudaGetLaunchCount": [_vp, C.POINTER(C.c_uint64)],
    "rrCudaCmdRebindSceneGeometry": [_vp, _vp, _vp, _vp, _vp],
    "rrCudaExportDeviceMemory": [_vp, _vp, _vp, C.POINTER(C.c_size_t)],
    "rrCudaImportDeviceMemory": [_vp, _vp, C.c_size_t, C.POINTER(_vp)],
    # radeonrays_cuda_debug.h
    "rrCudaDebugGetBuildScratchLayout": [_vp, C.c_uint32, C.POINTER(RRCudaBuildScratchLayout)],
    "rrCudaDebugSortPairs": [_vp, _vp, _vp, _vp, _vp, C.c_uint32],
    "rrCudaDebugRestructure": [_vp, _vp, C.c_uint32, _vp],
    "rrCudaDebugGetSceneLayout": [_vp, C.c_uint32, C.POINTER(RRCudaSceneLayout)],
}

_lib = None


def load():
    """Load the shared library; raises (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = lib
    return _lib


class RRException(RuntimeError):
    def __init__(self, call, code):
        super().__init__(f"{call} returned RRError {code}")
        self.code = code


def check(code, call="rr call"):
    if code != RR_SUCCESS:
        raise RRException(call, code)


class Context:
    """Convenience wrapper over an RRContext.  Every method is a thin pass-through to one rr* call."""

    def __init__(self, device=0, cuda_stream=None, use_default_create=False):
        self.lib = load()
        self.handle = _vp()
        if use_default_create:
            check(self.lib.rrCreateContext(RR_API_VERSION, RR_API_CUDA, C.byref(self.handle)), "rrCreateContext")
        else:
            check(self.lib.rrCreateContextCuda(RR_API_VERSION, device, _vp(cuda_stream), C.byref(self.handle)),
                  "rrCreateContextCuda")
        self._keep = []

    def destroy(self):
        if self.handle:
            check(self.lib.rrDestroyContext(self.handle), "rrDestroyContext")
            self.handle = _vp()

    # ---- device pointers -------------------------------------------------------------------------------
    def device_ptr(self, address, offset=0):
        p = _vp()
        check(self.lib.rrGetDevicePtrFromCudaPtr(self.handle, _vp(int(address)), offset, C.byref(p)), "rrGetDevicePtrFromCudaPtr")
        return p

    def tensor_ptr(self, tensor, offset=0):
        """RRDevicePtr over a torch CUDA tensor (the tensor must outlive the pointer's use)."""
        self._keep.append(tensor)
        return self.device_ptr(tensor.data_ptr(), offset)

    def allocate(self, size):
        p = _vp()
        check(self.lib.rrAllocateDeviceBuffer(self.handle, size, C.byref(p)), "rrAllocateDeviceBuffer")
        return p

    def map(self, ptr, size, dtype=np.uint8):
        m = _vp()
        check(self.lib.rrMapDevicePtr(self.handle, ptr, C.byref(m)), "rrMapDevicePtr")
        buf = (C.c_uint8 * size).from_address(m.value)
        return np.frombuffer(buf, dtype=dtype), m

    def unmap(self, ptr, mapping):
        check(self.lib.rrUnmapDevicePtr(self.handle, ptr, C.byref(mapping)), "rrUnmapDevicePtr")

    def release_ptr(self, ptr):
        check(self.lib.rrReleaseDevicePtr(self.handle, ptr), "rrReleaseDevicePtr")

    def raw_address(self, ptr):
        a = _vp()
        check(self.lib.rrGetCudaPtrFromDevicePtr(self.handle, ptr, C.byref(a)), "rrGetCudaPtrFromDevicePtr")
        return a.value

    # ---- command streams ---------------------------------------------------------------------------------
    def allocate_command_stream(self):
        s = _vp()
        check(self.lib.rrAllocateCommandStream(self.handle, C.byref(s)), "rrAllocateCommandStream")
        return s

    def command_stream_from_cuda_stream(self, cuda_stream):
        s = _vp()
        check(self.lib.rrGetCommandStreamFromCudaStream(self.handle, _vp(cuda_stream), C.byref(s)), "rrGetCommandStreamFromCudaStream")
        return s

    def submit(self, stream, wait_event=None):
        e = _vp()
        check(self.lib.rrSumbitCommandStream(self.handle, stream, wait_event, C.byref(e)), "rrSumbitCommandStream")
        return e

    def wait(self, event):
        check(self.lib.rrWaitEvent(self.handle, event), "rrWaitEvent")

    def release_event(self, event):
        check(self.lib.rrReleaseEvent(self.handle, event), "rrReleaseEvent")

    def release_command_stream(self, stream):
        check(self.lib.rrReleaseCommandStream(self.handle, stream), "rrReleaseCommandStream")

    def run(self, record):
        """allocate stream -> record(stream) -> submit -> wait -> release: the reference tests' idiom."""
        s = self.allocate_command_stream()
        try:
            record(s)
            e = self.submit(s)
            self.wait(e)
            self.release_event(e)
        finally:
            self.release_command_stream(s)

    # ---- build ---------------------------------------------------------------------------------------------
    @staticmethod
    def geometry_input(vertices_ptr, vertex_count, vertex_stride, indices_ptr, triangle_count, index_type=RR_INDEX_TYPE_UINT32):
        mesh = RRTriangleMeshPrimitive(vertices_ptr, vertex_count, vertex_stride, indices_ptr, triangle_count, index_type)
        gi = RRGeometryBuildInput(RR_PRIMITIVE_TYPE_TRIANGLE_MESH, 1, C.pointer(mesh))
        gi._mesh = mesh
        return gi

    @staticmethod
    def geometry_input_multi(meshes):
        """meshes: list of (vertices_ptr, vertex_count, vertex_stride, indices_ptr, triangle_count, index_type) -- a geometry made
        of several triangle meshes (primitive_count > 1; prim_id = running triangle index over the meshes)."""
        arr = (RRTriangleMeshPrimitive * len(meshes))(*[RRTriangleMeshPrimitive(*m) for m in meshes])
        gi = RRGeometryBuildInput(RR_PRIMITIVE_TYPE_TRIANGLE_MESH, len(meshes), C.cast(arr, C.POINTER(RRTriangleMeshPrimitive)))
        gi._mesh = arr
        return gi

    def geometry_requirements(self, geometry_input, options=None):
        req = RRMemoryRequirements()
        check(self.lib.rrGetGeometryBuildMemoryRequirements(self.handle, C.byref(geometry_input),
                                                            C.byref(options) if options is not None else None, C.byref(req)),
              "rrGetGeometryBuildMemoryRequirements")
        return req

    def cmd_build_geometry(self, op, geometry_input, options, temp_ptr, geometry_ptr, stream):
        check(self.lib.rrCmdBuildGeometry(self.handle, op, C.byref(geometry_input),
                                          C.byref(options) if options is not None else None, temp_ptr, geometry_ptr, stream),
              "rrCmdBuildGeometry")

    @staticmethod
    def scene_input(geometry_ptrs, transforms):
        n = len(geometry_ptrs)
        arr = (RRInstance * n)()
        t = np.ascontiguousarray(transforms, dtype=np.float32).reshape(n, 3, 4)
        for i in range(n):
            arr[i].geometry = geometry_ptrs[i]
            for r in range(3):
                for c in range(4):
                    arr[i].transform[r][c] = float(t[i, r, c])
        si = RRSceneBuildInput(arr, n)
        si._arr = arr
        return si

    def scene_requirements(self, scene_input, options=None):
        req = RRMemoryRequirements()
        check(self.lib.rrGetSceneBuildMemoryRequirements(self.handle, C.byref(scene_input),
                                                         C.byref(options) if options is not None else None, C.byref(req)),
              "rrGetSceneBuildMemoryRequirements")
        return req

    def cmd_build_scene(self, scene_input, options, temp_ptr, scene_ptr, stream):
        check(self.lib.rrCmdBuildScene(self.handle, C.byref(scene_input), C.byref(options) if options is not None else None,
                                       temp_ptr, scene_ptr, stream), "rrCmdBuildScene")

    # ---- trace ---------------------------------------------------------------------------------------------
    def trace_requirements(self, ray_count):
        sz = C.c_size_t()
        check(self.lib.rrGetTraceMemoryRequirements(self.handle, ray_count, C.byref(sz)), "rrGetTraceMemoryRequirements")
        return sz.value

    def cmd_intersect(self, scene_ptr, query, rays_ptr, ray_count, indirect_ptr, output, hits_ptr, scratch_ptr, stream):
        check(self.lib.rrCmdIntersect(self.handle, scene_ptr, query, rays_ptr, ray_count, indirect_ptr, output, hits_ptr,
                                      scratch_ptr, stream), "rrCmdIntersect")

    def cmd_rebind_scene_geometry(self, scene_ptr, old_address, new_geometry_ptr, stream):
        check(self.lib.rrCudaCmdRebindSceneGeometry(self.handle, scene_ptr, _vp(int(old_address)), new_geometry_ptr, stream),
              "rrCudaCmdRebindSceneGeometry")

    def export_memory(self, ptr):
        """-> (64-byte handle, offset) naming the allocation behind `ptr` for another process (rrCudaExportDeviceMemory)."""
        h = (C.c_ubyte * 64)()
        off = C.c_size_t()
        check(self.lib.rrCudaExportDeviceMemory(self.handle, ptr, C.cast(h, _vp), C.byref(off)), "rrCudaExportDeviceMemory")
        return bytes(h), off.value

    def import_memory(self, handle, offset=0):
        """RRDevicePtr onto another process's allocation (rrCudaImportDeviceMemory); release with release_ptr."""
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = _vp()
        check(self.lib.rrCudaImportDeviceMemory(self.handle, C.cast(buf, _vp), offset, C.byref(p)), "rrCudaImportDeviceMemory")
        return p

    # ---- options / debug -------------------------------------------------------------------------------------
    def set_option(self, option, value):
        check(self.lib.rrCudaSetOption(self.handle, option, int(value)), "rrCudaSetOption")

    def launch_count(self):
        n = C.c_uint64()
        check(self.lib.rrCudaGetLaunchCount(self.handle, C.byref(n)), "rrCudaGetLaunchCount")
        return n.value

    def build_scratch_layout(self, triangle_count):
        L = RRCudaBuildScratchLayout()
        check(self.lib.rrCudaDebugGetBuildScratchLayout(self.handle, triangle_count, C.byref(L)), "rrCudaDebugGetBuildScratchLayout")
        return L

    def scene_layout(self, instance_count):
        L = RRCudaSceneLayout()
        check(self.lib.rrCudaDebugGetSceneLayout(self.handle, instance_count, C.byref(L)), "rrCudaDebugGetSceneLayout")
        return L
