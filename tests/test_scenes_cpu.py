"""The bench's vectorised input generators (bvh_b200/scenes.py) == the scalar restatement of
src/testbase.rs in the oracle, bit for bit."""
import numpy as np
import pytest

from bvh_b200 import scenes
from oracle import oracle as O


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_cubes(prec):
    got = scenes.create_n_cubes_aabbs(300, prec)
    want = O.create_n_cubes(300, prec=prec)
    assert np.array_equal(got["min"], want["min"]) and np.array_equal(got["max"], want["max"])


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_rays(prec):
    o, d = scenes.ray_endpoints(500, 0, prec)
    rays, _ = O.create_rays(500, prec=prec)
    want = O.ray_new(o, d, prec)
    assert np.array_equal(rays["origin"], o)
    assert np.array_equal(rays["direction"], want["direction"]) and np.array_equal(rays["inv_direction"], want["inv_direction"])
    o2, d2 = scenes.ray_endpoints(100, 400, prec)          # a shard of the same chain
    assert np.array_equal(o2, o[400:]) and np.array_equal(d2, d[400:])
