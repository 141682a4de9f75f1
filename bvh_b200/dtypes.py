"""numpy views of the C-ABI PODs declared in include/bvh_b200.h."""
import numpy as np

U32_MAX = 0xFFFFFFFF


def _make(f):
    aabb = np.dtype([("min", f, (3,)), ("max", f, (3,))])
    ray = np.dtype([("origin", f, (3,)), ("direction", f, (3,)), ("inv_direction", f, (3,))])
    node = np.dtype([("parent", "<u4"), ("child_l", "<u4"), ("child_r", "<u4"), ("shape", "<u4"), ("l_aabb", aabb), ("r_aabb", aabb)])
    flat = np.dtype(
        {
            "names": ["aabb", "entry_index", "exit_index", "shape_index"],
            "formats": [aabb, "<u4", "<u4", "<u4"],
            "offsets": [0, aabb.itemsize, aabb.itemsize + 4, aabb.itemsize + 8],
            "itemsize": 36 if f == "<f4" else 64,
        }
    )
    return aabb, ray, node, flat


AABB3F, RAY3F, NODE3F, FLAT3F = _make("<f4")
AABB3D, RAY3D, NODE3D, FLAT3D = _make("<f8")

BY_PREC = {
    "f32": dict(scalar=np.float32, aabb=AABB3F, ray=RAY3F, node=NODE3F, flat=FLAT3F, suffix="f32x3"),
    "f64": dict(scalar=np.float64, aabb=AABB3D, ray=RAY3D, node=NODE3D, flat=FLAT3D, suffix="f64x3"),
}
assert AABB3F.itemsize == 24 and RAY3F.itemsize == 36 and NODE3F.itemsize == 64 and FLAT3F.itemsize == 36
assert AABB3D.itemsize == 48 and RAY3D.itemsize == 72 and NODE3D.itemsize == 112 and FLAT3D.itemsize == 64
