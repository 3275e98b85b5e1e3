"""rocquantum_b200 -- B200-native state-vector engine behind the rocQuantum hipStateVec C ABI.

The product is the shared library (rocquantum_b200/lib/libhipStateVec{,_f64}.so, built by
rocquantum_b200/build.py from csrc/).  This package is the thin Python host side:

  capi      ctypes prototypes of every entry point declared in include/hipStateVec.h
  backend   the free-function surface of the reference's `_rocq_hip_backend` pybind module
            (python/rocq/bindings.cpp:142-494), same names and argument order

There is no CPU fallback anywhere: loading fails loudly if the library is missing, and
rocsvCreate fails with HIP_ERROR if no CUDA device is usable.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
