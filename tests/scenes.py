"""Seeded input scenes shared by the parity tests (oracle side = test infrastructure)."""
import numpy as np

from oracle import oracle as O


def scene(name: str, prec: str = "f32") -> np.ndarray:
    rng = np.random.default_rng(abs(hash(name)) % (2**32) if False else sum(map(ord, name)))
    if name == "empty":
        return np.zeros(0, dtype=O._DT[prec]["aabb"])
    if name == "boxes21":
        return O.aligned_boxes(prec)
    if name.startswith("cubes"):
        return O.create_n_cubes(int(name[5:]), prec=prec)
    if name.startswith("random"):          # random boxes, varied sizes
        n = int(name[6:])
        mn = rng.uniform(-1000, 1000, (n, 3))
        return O.make_aabbs(mn, mn + rng.uniform(0, 50, (n, 3)) ** 2 / 50, prec)
    if name.startswith("points"):          # clusters of coincident degenerate boxes: exercises the halving branch
        n = int(name[6:])
        base = rng.integers(-5, 5, (max(n // 7, 1), 3)).astype(float)
        mn = base[rng.integers(0, len(base), n)]
        return O.make_aabbs(mn, mn, prec)
    if name.startswith("line"):            # all centroids on one axis, many ties
        n = int(name[4:])
        x = rng.integers(0, max(n // 3, 2), n).astype(float)
        mn = np.stack([x, np.zeros(n), np.zeros(n)], axis=1)
        return O.make_aabbs(mn - 0.25, mn + 0.25, prec)
    if name.startswith("huge"):            # surface areas overflow in f32: "no split wins" fallthrough
        n = int(name[4:])
        mn = rng.uniform(-1e30, 1e30, (n, 3))
        return O.make_aabbs(mn, mn + rng.uniform(0, 1e29, (n, 3)), prec)
    if name.startswith("skew"):            # one far outlier per level: very unbalanced tree, deep recursion
        n = int(name[4:])
        x = 2.0 ** np.arange(n) if n < 100 else np.cumsum(rng.uniform(0, 1, n) ** 8 * 1e4)
        mn = np.stack([x, rng.uniform(0, 1, n), rng.uniform(0, 1, n)], axis=1)
        return O.make_aabbs(mn, mn + 0.5, prec)
    raise KeyError(name)


def rays_for(shapes: np.ndarray, n: int, prec: str = "f32", seed: int = 0, axis_aligned: int = 0) -> np.ndarray:
    """Random rays through the scene's bounds; `axis_aligned` of them get exact zero direction components."""
    rng = np.random.default_rng(seed)
    if len(shapes):
        lo, hi = shapes["min"].min(axis=0).astype(float), shapes["max"].max(axis=0).astype(float)
    else:
        lo, hi = np.full(3, -1.0), np.full(3, 1.0)
    pad = (hi - lo) * 0.1 + 1e-3
    origins = rng.uniform(lo - pad, hi + pad, (n, 3))
    targets = rng.uniform(lo, hi, (n, 3))
    dirs = targets - origins
    for i in range(min(axis_aligned, n)):
        d = np.zeros(3)
        d[rng.integers(0, 3)] = rng.choice([-1.0, 1.0])
        dirs[i] = d
        if len(shapes) and i % 2 == 0:       # start exactly on a box face plane: NaN rule
            origins[i] = shapes["min"][rng.integers(0, len(shapes))].astype(float)
    return O.ray_new(origins, dirs, prec)
