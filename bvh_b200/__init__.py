"""bvh_b200 -- B200-native (sm_100a) build / flatten / batched-ray-traversal path of svenstaro/bvh.

The package is a thin host layer over libbvh_b200.so (include/bvh_b200.h).  Importing `bvh_b200.api`
loads the shared library and raises if it is missing: there is no CPU fallback.
"""
__all__ = ["api", "capi", "dtypes", "build"]
