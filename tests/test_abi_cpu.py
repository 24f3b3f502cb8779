"""CPU tests (-m "not gpu"): the C-ABI library loads, exports every symbol include/*.h declares, and applies the
reference's argument checks (src/core/src/radeonrays.cpp) without needing a device."""
import ctypes as C
import os
import re

import pytest

from radeonrays_sdk_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in ("radeonrays.h", "radeonrays_cuda.h", "radeonrays_cuda_debug.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        names |= set(re.findall(r"RR_API\s+RRError\s+(rr\w+)\s*\(", text))
    return names


def test_every_declared_symbol_is_exported_and_bound():
    lib = api.load()
    decl = declared_symbols()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert decl == set(api.SIGNATURES), "python binding and headers must list the same entry points"
    # the 17 core entry points of the reference header (radeonrays.h:276-473), typo included
    for name in ("rrCreateContext", "rrDestroyContext", "rrSetLogLevel", "rrSetLogFile", "rrCmdBuildGeometry",
                 "rrGetGeometryBuildMemoryRequirements", "rrCmdBuildScene", "rrGetSceneBuildMemoryRequirements",
                 "rrCmdIntersect", "rrGetTraceMemoryRequirements", "rrAllocateCommandStream", "rrReleaseCommandStream",
                 "rrSumbitCommandStream", "rrReleaseEvent", "rrWaitEvent", "rrReleaseDevicePtr", "rrReleaseExternalCommandStream"):
        assert name in decl


def test_struct_layouts_match_the_reference_abi():
    assert C.sizeof(api.RRBuildOptions) == 16
    assert C.sizeof(api.RRTriangleMeshPrimitive) == 32
    assert C.sizeof(api.RRGeometryBuildInput) == 16
    assert C.sizeof(api.RRInstance) == 56
    assert C.sizeof(api.RRSceneBuildInput) == 16
    assert C.sizeof(api.RRMemoryRequirements) == 24
    from radeonrays_sdk_b200.workloads import HIT_DTYPE, NODE_DTYPE, RAY_DTYPE
    assert (RAY_DTYPE.itemsize, HIT_DTYPE.itemsize, NODE_DTYPE.itemsize) == (32, 16, 64)


def test_null_argument_checks_without_a_device():
    lib = api.load()
    ctx = C.c_void_p()
    assert lib.rrCreateContext(api.RR_API_VERSION, api.RR_API_CUDA, None) == api.RR_ERROR_INVALID_PARAMETER
    assert lib.rrCreateContext(api.RR_API_VERSION, api.RR_API_VK, C.byref(ctx)) == api.RR_ERROR_UNSUPPORTED_API
    assert lib.rrCreateContext(api.RR_API_VERSION, api.RR_API_DX, C.byref(ctx)) == api.RR_ERROR_UNSUPPORTED_API
    assert lib.rrDestroyContext(None) == api.RR_ERROR_INVALID_PARAMETER
    assert lib.rrCmdIntersect(None, None, 0, None, 1, None, 0, None, None, None) == api.RR_ERROR_INVALID_PARAMETER
    assert lib.rrCmdBuildGeometry(None, 1, None, None, None, None, None) == api.RR_ERROR_INVALID_PARAMETER
    assert lib.rrCmdBuildScene(None, None, None, None, None, None) == api.RR_ERROR_INVALID_PARAMETER
    assert lib.rrGetTraceMemoryRequirements(None, 1, None) == api.RR_ERROR_INVALID_PARAMETER
    assert lib.rrSumbitCommandStream(None, None, None, None) == api.RR_ERROR_INVALID_PARAMETER
    assert lib.rrWaitEvent(None, None) == api.RR_ERROR_INVALID_PARAMETER
    assert lib.rrReleaseDevicePtr(None, None) == api.RR_ERROR_INVALID_PARAMETER
    assert lib.rrSetLogLevel(0) == api.RR_ERROR_INVALID_PARAMETER
    assert lib.rrSetLogLevel(5) == api.RR_SUCCESS
    assert lib.rrSetLogLevel(3) == api.RR_SUCCESS


def test_no_device_means_error_not_fallback():
    """Without a GPU the product path must fail loudly (RR_ERROR_INTERNAL), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    lib = api.load()
    ctx = C.c_void_p()
    lib.rrSetLogLevel(5)
    try:
        assert lib.rrCreateContext(api.RR_API_VERSION, api.RR_API_CUDA, C.byref(ctx)) == api.RR_ERROR_INTERNAL
        assert not ctx.value
    finally:
        lib.rrSetLogLevel(3)


def test_product_does_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under radeonrays_sdk_b200/ may import, link or execute it."""
    pkg = os.path.join(ROOT, "radeonrays_sdk_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("#include \"rr_oracle", "librr_oracle", "from oracle", "import oracle", "oracle.binding"):
                    assert needle not in text, (f, needle)
