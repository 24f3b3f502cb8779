"""radeonrays_sdk_b200 -- host-side mirror of the rr* C API over libradeonrays_b200.so (sm_100a kernels).

The product is the C-ABI shared library built from csrc/ (include/radeonrays.h, radeonrays_cuda.h).  This
package only binds it with ctypes so that tests and bench.py can drive exactly the calls a C/C++ client of
RadeonRays makes.  There is no CPU fallback: importing `api` raises if the library has not been built.
"""
from . import workloads  # noqa: F401

__all__ = ["workloads", "api"]


def __getattr__(name):
    if name == "api":
        import importlib
        return importlib.import_module(".api", __name__)
    raise AttributeError(name)
