"""Pins the CPU oracle (oracle/) against every golden / known-answer test the reference holds for
the build -> flatten -> traverse path, and against an independent numpy restatement
(tests/pyref.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import pyref

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
MODES = {"recursive": O.MODE_RECURSIVE, "flat": O.MODE_FLAT, "iterator": O.MODE_ITERATOR}


def _trav(mode, res, flat, shapes, rays, prec="f32"):
    tree = flat if mode == O.MODE_FLAT else res.nodes
    return O.traverse(tree, shapes, rays, mode, prec)


# ---- (1) 21-box scene hit sets: testbase.rs:174-225, iter.rs:269-308 ------------------------------
@pytest.mark.parametrize("threads", [1, 4])           # Bvh::build and Bvh::build_par
@pytest.mark.parametrize("mode", list(MODES))
def test_aligned_boxes_hit_sets(mode, threads):
    boxes = O.aligned_boxes()
    res = O.build(boxes, threads=threads)
    flat = O.flatten(res.nodes)
    for case in G["aligned_boxes_21"]["rays"]:
        ray = O.ray_new([case["origin"]], [case["direction"]])
        t = _trav(MODES[mode], res, flat, boxes, ray)
        ids = sorted(int(h) - 10 for h in t.hits)
        assert ids == sorted(case["hit_ids"]), (mode, case)
        assert len(t.hits) == len(case["hit_ids"])


# ---- (2) degenerate sizes: bvh_impl.rs:564-574, 665-690; flat_bvh.rs:620-625 ----------------------
def test_empty():
    e = np.zeros(0, dtype=O.AABB3F)
    res = O.build(e)
    assert len(res.nodes) == 0
    assert len(O.flatten(res.nodes)) == 0
    ray = O.ray_new([[0, 0, 0]], [[1, 0, 0]])
    for m in MODES.values():
        tree = O.flatten(res.nodes) if m == O.MODE_FLAT else res.nodes
        assert len(O.traverse(tree, e, ray, m).hits) == 0
    assert O.is_consistent(res.nodes, e) and O.is_tight(res.nodes)


@pytest.mark.parametrize("case", G["one_node_bvh"]["cases"])
def test_one_node(case):
    boxes = O.unit_boxes([case["box_center"]])
    res = O.build(boxes)
    assert len(res.nodes) == 1 and res.nodes[0]["child_l"] == O.U32_MAX and res.nodes[0]["shape"] == 0
    flat = O.flatten(res.nodes)
    assert len(flat) == 1 and flat[0]["entry_index"] == O.U32_MAX and flat[0]["exit_index"] == 1
    ray = O.ray_new([case["origin"]], [case["direction"]])
    for m in MODES.values():
        assert len(_trav(m, res, flat, boxes, ray).hits) == case["hits"]


# ---- (3)+(4) every shape in exactly one leaf; consistent + tight: bvh_impl.rs:588-614, optimization.rs:648-656
@pytest.mark.parametrize("scene", ["boxes21", "cubes100", "cubes1000"])
def test_structure_invariants(scene):
    shapes = O.aligned_boxes() if scene == "boxes21" else O.create_n_cubes(int(scene[5:]))
    for threads in (1, 4):
        res = O.build(shapes, threads=threads)
        n = len(shapes)
        assert len(res.nodes) == 2 * n - 1
        leaves = res.nodes[res.nodes["child_l"] == O.U32_MAX]
        assert sorted(leaves["shape"].tolist()) == list(range(n))
        assert np.array_equal(res.nodes["shape"][res.node_index], np.arange(n))
        assert O.is_consistent(res.nodes, shapes)
        assert O.is_tight(res.nodes)


def test_moved_shapes_make_tree_inconsistent():        # optimization.rs:657-663 (negative control for the checker)
    shapes = O.create_n_cubes(100)
    res = O.build(shapes)
    moved = shapes.copy()
    moved["min"][::3] += 1000.0
    moved["max"][::3] += 1000.0
    assert not O.is_consistent(res.nodes, moved)


# ---- (5) SAH pairs boxes -50/-40 against 50: optimization.rs:421-455 ------------------------------
def test_sah_pairing():
    k = G["sah_pairing"]
    boxes = O.unit_boxes(k["box_centers"])
    res = O.build(boxes)
    a, b = k["same_parent"]
    na, nb = res.nodes[res.node_index[a]], res.nodes[res.node_index[b]]
    assert na["child_l"] == O.U32_MAX and nb["child_l"] == O.U32_MAX
    assert na["parent"] == nb["parent"]


# ---- (6) fuzz.rs:299-329 grid-mode property: traverse == traverse_iterator == flatten().traverse ----
@pytest.mark.parametrize("seed", range(8))
def test_grid_mode_agreement(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2, 33))                       # fuzz.rs:427-438 caps trees at 32 shapes (iterator stack)
    mins = rng.integers(-20, 20, size=(n, 3)).astype(np.float32)
    size = rng.integers(0, 4, size=(n, 3)).astype(np.float32)
    shapes = O.make_aabbs(mins, mins + size)
    res = O.build(shapes)
    flat = O.flatten(res.nodes)
    assert O.is_consistent(res.nodes, shapes) and O.is_tight(res.nodes)
    origins = rng.integers(-25, 25, size=(64, 3)).astype(np.float32) + np.float32(1.0 / 3.0)
    dirs = np.zeros((64, 3), dtype=np.float32)
    dirs[np.arange(64), rng.integers(0, 3, 64)] = rng.choice([-1.0, 1.0], 64)
    rays = O.ray_new(origins, dirs)
    r0 = O.traverse(res.nodes, shapes, rays, O.MODE_RECURSIVE)
    r1 = O.traverse(flat, shapes, rays, O.MODE_FLAT)
    r2 = O.traverse(res.nodes, shapes, rays, O.MODE_ITERATOR)
    assert not r2.iter_overflow
    for a, b, c in zip(O.per_ray_lists(r0.offsets, r0.hits), O.per_ray_lists(r1.offsets, r1.hits), O.per_ray_lists(r2.offsets, r2.hits)):
        assert set(a.tolist()) == set(b.tolist()) == set(c.tolist())
    # brute force over shape AABBs must contain every reported hit (a BVH never invents hits)
    for r, lst in zip(rays, O.per_ray_lists(r0.offsets, r0.hits)):
        for h in lst:
            assert O.ray_intersects_aabb(r, shapes[h])


# ---- (7) ray / aabb KATs --------------------------------------------------------------------------
@pytest.mark.parametrize("case", G["ray_aabb"]["cases"], ids=lambda c: c["name"])
def test_ray_aabb_kats(case):
    ray = O.ray_new([case["origin"]], [case["direction"]])
    box = O.make_aabbs([case["min"]], [case["max"]])
    assert O.ray_intersects_aabb(ray, box) == case["hit"]


def test_ray_new_unit_direction():                      # ray_impl.rs:60-65
    ray = O.ray_new([[0, 0, 0]], [[1, 0, 0]])[0]
    assert ray["direction"].tolist() == [1.0, 0.0, 0.0]
    assert ray["inv_direction"][0] == 1.0 and np.isinf(ray["inv_direction"][1]) and np.isinf(ray["inv_direction"][2])


def test_ray_points_at_center_proptest():               # ray_impl.rs:304-330 restated with a seeded generator
    rng = np.random.default_rng(0)
    for _ in range(300):
        a, b, pos = (rng.uniform(-1e11, 1e11, 3).astype(np.float32) for _ in range(3))
        box = O.make_aabbs([np.minimum(a, b)], [np.maximum(a, b)])
        c, _, _ = O.aabb_ops(box)
        ray = O.ray_new([pos], [c - pos])
        assert O.ray_intersects_aabb(ray, box)
        back = ray.copy()
        back["direction"] = -back["direction"]
        back["inv_direction"] = -back["inv_direction"]
        inside = bool(np.all(pos >= box["min"][0]) and np.all(pos <= box["max"][0]))
        assert (not O.ray_intersects_aabb(back, box)) or inside


def test_aabb_kats():
    k = G["aabb"]
    _, _, axis = O.aabb_ops(O.make_aabbs([k["largest_axis"]["min"]], [k["largest_axis"]["max"]]))
    assert axis == k["largest_axis"]["axis"]
    p1, p2 = np.float32(k["center_overflow"]["p1"]), np.float32(k["center_overflow"]["p2"])
    box = O.make_aabbs([np.minimum(p1, p2)], [np.maximum(p1, p2)])
    with np.errstate(over="ignore"):
        assert np.isinf(box["max"][0][0] - box["min"][0][0])
    c, _, _ = O.aabb_ops(box)
    assert np.isfinite(c[0]) and box["min"][0][0] <= c[0] <= box["max"][0][0]
    rng = np.random.default_rng(1)
    for _ in range(200):                                 # aabb_impl.rs:889-899
        s = np.float32(10 ** rng.uniform(-6, 15))
        pos = rng.uniform(-1e3, 1e3, 3).astype(np.float32) * np.float32(0)   # exact corner so that size == s
        _, sa, _ = O.aabb_ops(O.make_aabbs([pos], [pos + s]))
        want = float(np.float32(6) * s * s)
        # 2*rn(3*rn(s*s)) vs rn(rn(6s)*s): any association of the 3-term dot gives the former, which can sit
        # 2 ulp from the latter for rare s (the reference's proptest samples 256 values); allow 2 eps here.
        assert abs(float(sa) - want) <= 2 * float(np.finfo(np.float32).eps) * max(abs(float(sa)), abs(want))


# ---- derived fingerprints (independent survey-session restatement; not reference-published) --------
def test_derived_fingerprints():
    d = G["derived"]
    res = O.build(O.aligned_boxes())
    assert res.node_index.tolist() == d["aligned_boxes_21_node_index"]
    for key, sc in d["scenes"].items():
        n = int(key)
        shapes = O.aligned_boxes() if n == 21 else O.create_n_cubes(n // 12)
        res = O.build(shapes)
        flat = O.flatten(res.nodes)
        assert len(res.nodes) == sc["bvh_nodes"] and len(flat) == sc["flat_nodes"]
        assert res.max_depth == sc["max_depth"] and res.degenerate_splits == sc["degenerate_splits"]
        # the survey counted bucketed nodes only; the oracle's P also counts degenerate nodes (2 prims each here)
        assert res.prim_visits - 2 * res.degenerate_splits == sc["bucketed_prim_visits"]
        if "rays" in sc:
            rays, _ = O.create_rays(sc["rays"])
            t = O.traverse(flat, shapes, rays, O.MODE_FLAT)
            assert len(t.hits) / sc["rays"] == sc["hits_per_ray"]


# ---- oracle == independent numpy restatement, bit for bit ------------------------------------------
def _cmp_nodes(py_nodes, nodes):
    assert len(py_nodes) == len(nodes)
    for i, p in enumerate(py_nodes):
        o = nodes[i]
        assert p[1] == o["parent"], i
        if p[0] == "leaf":
            assert o["child_l"] == O.U32_MAX and p[2] == o["shape"], i
        else:
            assert (p[2], p[3]) == (o["child_l"], o["child_r"]), i
            assert np.array_equal(np.array(p[4][0]), o["l_aabb"]["min"]) and np.array_equal(np.array(p[4][1]), o["l_aabb"]["max"]), i
            assert np.array_equal(np.array(p[5][0]), o["r_aabb"]["min"]) and np.array_equal(np.array(p[5][1]), o["r_aabb"]["max"]), i


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("scene", ["boxes21", "cubes40", "random300", "huge", "points"])
def test_oracle_equals_pyref(scene, prec):
    F = np.float32 if prec == "f32" else np.float64
    rng = np.random.default_rng(7)
    if scene == "boxes21":
        shapes = O.aligned_boxes(prec)
    elif scene == "cubes40":
        shapes = O.create_n_cubes(40, prec=prec)
    elif scene == "random300":
        mn = rng.uniform(-100, 100, (300, 3))
        shapes = O.make_aabbs(mn, mn + rng.uniform(0, 30, (300, 3)), prec)
    elif scene == "huge":                                # SA overflows to +inf in f32: "no split wins" fallthrough
        mn = rng.uniform(-1e30, 1e30, (64, 3))
        shapes = O.make_aabbs(mn, mn + rng.uniform(0, 1e29, (64, 3)), prec)
    else:                                                # many coincident centroids: degenerate halving
        mn = np.repeat(rng.integers(-3, 3, (20, 3)).astype(float), 5, axis=0)
        shapes = O.make_aabbs(mn, mn, prec)
    res = O.build(shapes, prec)
    py_nodes, py_idx = pyref.build(shapes, F)
    _cmp_nodes(py_nodes, res.nodes)
    assert py_idx == res.node_index.tolist()
    flat = O.flatten(res.nodes, prec)
    py_flat = pyref.flatten(py_nodes)
    assert len(py_flat) == len(flat)
    for pf, of in zip(py_flat, flat):
        assert (pf[1], pf[2], pf[3]) == (of["entry_index"], of["exit_index"], of["shape_index"])
        if pf[0] is not None:
            assert np.array_equal(np.array(pf[0][0]), of["aabb"]["min"]) and np.array_equal(np.array(pf[0][1]), of["aabb"]["max"])
    if scene == "huge" and prec == "f32":
        assert res.nosplit_fallthrough > 0
    if scene == "points":
        assert res.degenerate_splits > 0
    # traversal: oracle (three variants) == pyref recursive, as sequences
    b = shapes["min"].min(axis=0), shapes["max"].max(axis=0)
    origins = rng.uniform(b[0], b[1], (40, 3))
    targets = rng.uniform(b[0], b[1], (40, 3))
    rays = O.ray_new(origins, targets - origins, prec)
    r0 = O.traverse(res.nodes, shapes, rays, O.MODE_RECURSIVE, prec)
    r1 = O.traverse(flat, shapes, rays, O.MODE_FLAT, prec)
    if res.nosplit_fallthrough == 0:
        assert np.array_equal(r0.hits, r1.hits) and np.array_equal(r0.offsets, r1.offsets)
    else:
        # "no split wins" nodes store Aabb::empty() for their children (bvh_node.rs:225-230); an empty AABB
        # passes the slab test for every ray (tmin=-inf, tmax=+inf), so Bvh::traverse (no leaf re-test)
        # reports a superset of FlatBvh::traverse (re-tests the shape AABB, flat_bvh.rs:412-416).
        for a, b in zip(O.per_ray_lists(r0.offsets, r0.hits), O.per_ray_lists(r1.offsets, r1.hits)):
            assert set(b.tolist()) <= set(a.tolist())
    for ray, lst in zip(rays, O.per_ray_lists(r0.offsets, r0.hits)):
        pr = pyref.ray_new(F, ray["origin"], ray["direction"])
        pr = ([F(v) for v in ray["origin"]], None, [F(v) for v in ray["inv_direction"]])
        got = pyref.traverse_recursive(py_nodes, shapes, (pr[0], pr[2]), F)
        assert got == lst.tolist()


def test_ray_new_matches_pyref():
    rng = np.random.default_rng(3)
    o = rng.uniform(-1e5, 1e5, (50, 3)).astype(np.float32)
    d = rng.uniform(-1e5, 1e5, (50, 3)).astype(np.float32)
    rays = O.ray_new(o, d)
    for i in range(50):
        po, pd, pinv = pyref.ray_new(np.float32, o[i], d[i])
        assert np.array_equal(np.array(pd), rays[i]["direction"]) and np.array_equal(np.array(pinv), rays[i]["inv_direction"])


# ---- closed-form flatten == literal recursion ------------------------------------------
@pytest.mark.parametrize("n_cubes", [1, 5, 100])
def test_closed_form_flatten(n_cubes):
    shapes = O.create_n_cubes(n_cubes)
    res = O.build(shapes)
    flat = O.flatten(res.nodes)
    nodes = res.nodes
    n = len(shapes)
    is_leaf = nodes["child_l"] == O.U32_MAX
    leaves_before = np.concatenate([[0], np.cumsum(is_leaf)[:-1]])
    count = np.where(is_leaf, 1, nodes["shape"])
    assert len(flat) == 3 * n - 2
    for i in range(1, len(nodes)):
        nav = (i - 1) + leaves_before[i]
        par = nodes[nodes[i]["parent"]]
        aabb = par["l_aabb"] if par["child_l"] == i else par["r_aabb"]
        f = flat[nav]
        assert f["entry_index"] == nav + 1 and f["exit_index"] == nav + 3 * count[i] - 1 and f["shape_index"] == O.U32_MAX
        assert f["aabb"] == aabb
        if is_leaf[i]:
            g = flat[nav + 1]
            assert g["entry_index"] == O.U32_MAX and g["exit_index"] == nav + 2 and g["shape_index"] == nodes[i]["shape"]


# ---- the non-ray queries of traverse_some_built_bh (testbase.rs:227-266) ---------------------------------------------
@pytest.mark.parametrize("use_flat", [False, True])
def test_aligned_boxes_point_aabb_sphere_queries(use_flat):
    boxes = O.aligned_boxes()
    res = O.build(boxes)
    flat = O.flatten(res.nodes) if use_flat else None
    cases = [(O.QUERY_POINT, [[0, 0, 0]], {0}), (O.QUERY_POINT, [[0, 1000, 0]], set()),
             (O.QUERY_AABB, [[5.1, -1, -1, 9.9, 1, 1]], set(range(5, 11))), (O.QUERY_BALL, [[5, -1, -1, 1.4]], {4, 5, 6})]
    for kind, q, want in cases:
        _, hits = O.query(kind, q, res.nodes, boxes, flat)
        assert sorted(int(h) - 10 for h in hits) == sorted(want)


def test_ball_doctest():                                # src/ball.rs:74-84
    box = O.make_aabbs([[1.25, 1.25, 1.25]], [[3.0, 3.0, 3.0]])
    res = O.build(box)
    _, hits = O.query(O.QUERY_BALL, [[1.0, 1.0, 1.0, 1.0]], res.nodes, box)
    assert hits.tolist() == [0]


def test_ray_slice_kats():                              # ray_impl.rs:256-299
    box = O.make_aabbs([[-6.0, -8.0, -5.0]], [[-3.0, -4.0, 5.0]])
    ray = O.ray_new([[2.0, 2.0, 2.0]], [[-5.0, -8.66666, -3.666666]])
    tmin, tmax = O.ray_slice(ray, box)
    assert abs(tmin - 10.6562) < 0.01 and abs(tmax - 12.3034) < 0.01
    unit = O.unit_boxes([[-50.0, -50.0, -25.0]])
    assert O.ray_slice(O.ray_new([[-50.0, -50.0, -50.0]], [[1.0, 0.0, 0.0]]), unit) is None          # parallel ray
    assert O.ray_slice(O.ray_new([[0.0, 0.0, -0.5]], [[1.0, 0.0, 0.0]]), O.unit_boxes([[0.0, 0.0, 0.0]])) is None   # in-plane
    # fuzz.rs:387-407: intersects_aabb <=> slice.is_some()
    rng = np.random.default_rng(4)
    for _ in range(500):
        mn = rng.integers(-5, 5, 3).astype(np.float32); b = O.make_aabbs([mn], [mn + rng.integers(0, 4, 3)])
        r = O.ray_new([rng.integers(-6, 6, 3) + 1.0 / 3.0], [rng.choice([-1.0, 0.0, 1.0, 0.3], 3) + np.array([0.0, 0.0, 1e-3])])
        assert O.ray_intersects_aabb(r, b) == (O.ray_slice(r, b) is not None)


def test_bench_build_byte_model_constant():
    """bench.py's algorithmic-byte model of the build uses P = sum over internal nodes of their range size for the 120k scene."""
    import bench

    res = O.build(O.create_n_cubes(10_000))
    assert res.prim_visits == bench.BUILD_PRIM_VISITS


# ---------------------------------------------------------------------------------------------------------------------
# Bvh::update_shapes (src/bvh/optimization.rs): the reference's own tests, restated on the oracle.
# ---------------------------------------------------------------------------------------------------------------------
def _leaf_parent(nodes, node_index, shape):
    return int(nodes[int(node_index[shape])]["parent"])


def test_update_shapes_simple_update():                    # optimization.rs:420-487
    pos = np.array([[-50.0, 0, 0], [-40.0, 0, 0], [50.0, 0, 0]])
    shapes = O.unit_boxes(pos)
    b = O.build(shapes)
    assert _leaf_parent(b.nodes, b.node_index, 0) == _leaf_parent(b.nodes, b.node_index, 1)     # SAH joined #0 and #1
    pos[1] = [40.0, 0, 0]
    shapes = O.unit_boxes(pos)
    nodes, node_index = O.update_shapes(b.nodes, b.node_index, shapes, [1])
    assert O.is_consistent(nodes, shapes)
    assert _leaf_parent(nodes, node_index, 1) == _leaf_parent(nodes, node_index, 2)             # now #1 and #2


def test_consistent_after_update_shapes():                 # optimization.rs:405-417
    shapes = O.aligned_boxes()
    b = O.build(shapes)
    moved = O.unit_boxes([[10.0, 1.0, 2.0], [-10.0, -10.0, 10.0], [-10.0, 10.0, 10.0], [-10.0, 10.0, -10.0], [11.0, 1.0, 2.0], [11.0, 2.0, 2.0]])
    shapes = shapes.copy()
    shapes[:6] = moved
    nodes, node_index = O.update_shapes(b.nodes, b.node_index, shapes, np.arange(6))
    assert len(nodes) == 41
    assert O.is_consistent(nodes, shapes)
    for s in range(21):                                       # every shape is still in exactly one leaf, and knows which
        nd = nodes[int(node_index[s])]
        assert nd["child_l"] == 0xFFFFFFFF and nd["shape"] == s


def _predictable_bvh():                                     # optimization.rs:489-541
    shapes = O.unit_boxes([[0.0, 0, 0], [2.0, 0, 0], [4.0, 0, 0], [6.0, 0, 0]])
    nodes = np.zeros(7, dtype=O.NODE3F)
    INV = 0xFFFFFFFF
    def join(a, b):
        out = np.zeros((), dtype=O.AABB3F)
        out["min"] = np.minimum(a["min"], b["min"]); out["max"] = np.maximum(a["max"], b["max"])
        return out
    empty = np.zeros((), dtype=O.AABB3F); empty["min"] = np.inf; empty["max"] = -np.inf
    def node(i, parent, l, r, la, ra, cnt):
        nodes[i]["parent"], nodes[i]["child_l"], nodes[i]["child_r"], nodes[i]["shape"] = parent, l, r, cnt
        nodes[i]["l_aabb"], nodes[i]["r_aabb"] = la, ra
    node(0, 0, 1, 2, join(shapes[0], shapes[1]), join(shapes[2], shapes[3]), 4)
    node(1, 0, 3, 4, shapes[0], shapes[1], 2)
    node(2, 0, 5, 6, shapes[2], shapes[3], 2)
    for i, (p, s) in enumerate([(1, 0), (1, 1), (2, 2), (2, 3)]):
        node(3 + i, p, INV, INV, empty, empty, s)
    return shapes, nodes


def _same(a, b):
    return np.array_equal(a["min"], b["min"]) and np.array_equal(a["max"], b["max"])


def test_connect_grandchildren():                          # optimization.rs:543-589
    shapes, nodes = _predictable_bvh()
    nodes = O.connect_nodes(nodes, shapes, 3, 2, True)
    nodes = O.connect_nodes(nodes, shapes, 5, 1, True)
    assert [int(nodes[i]["parent"]) for i in range(7)] == [0, 0, 0, 2, 1, 1, 2]
    assert (int(nodes[0]["child_l"]), int(nodes[0]["child_r"])) == (1, 2)
    assert (int(nodes[1]["child_l"]), int(nodes[1]["child_r"])) == (5, 4)
    assert (int(nodes[2]["child_l"]), int(nodes[2]["child_r"])) == (3, 6)
    assert _same(nodes[1]["l_aabb"], shapes[2]) and _same(nodes[1]["r_aabb"], shapes[1])
    assert _same(nodes[2]["l_aabb"], shapes[0]) and _same(nodes[2]["r_aabb"], shapes[3])


def test_connect_child_grandchild():                       # optimization.rs:591-637
    shapes, nodes = _predictable_bvh()
    nodes = O.connect_nodes(nodes, shapes, 1, 2, True)
    nodes = O.connect_nodes(nodes, shapes, 5, 0, True)
    assert [int(nodes[i]["parent"]) for i in range(7)] == [0, 2, 0, 1, 1, 0, 2]
    assert (int(nodes[0]["child_l"]), int(nodes[0]["child_r"])) == (5, 2)
    assert (int(nodes[1]["child_l"]), int(nodes[1]["child_r"])) == (3, 4)
    assert (int(nodes[2]["child_l"]), int(nodes[2]["child_r"])) == (1, 6)
    assert _same(nodes[0]["l_aabb"], shapes[2]) and _same(nodes[2]["r_aabb"], shapes[3])
    assert _same(nodes[1]["l_aabb"], shapes[0]) and _same(nodes[1]["r_aabb"], shapes[1])


def _move_shapes(shapes, amount, rng, max_offset=None):
    """randomly_transform_scene analogue (testbase.rs:640-681): translate `amount` distinct shapes by a random offset that
    keeps them inside the default bounds (the reference shuffles with StdRng, which cannot be restated: seeded numpy here)."""
    bounds = O.default_bounds()
    idx = rng.permutation(len(shapes))[:amount]
    out = shapes.copy()
    lo = bounds["min"][0] - shapes["min"][idx]
    hi = bounds["max"][0] - shapes["max"][idx]
    off = rng.uniform(lo, hi).astype(np.float32)
    if max_offset is not None:
        off = np.clip(off, -max_offset, max_offset)
    out["min"][idx] = (shapes["min"][idx] + off).astype(np.float32)
    out["max"][idx] = (shapes["max"][idx] + off).astype(np.float32)
    return out, idx.astype(np.uint32)


def _brute_force(shapes, ray):
    """Every shape whose AABB the ray hits (intersect_default.rs:16-37, vectorised in f32: sub, mul, min/max -- no FMA in numpy)."""
    with np.errstate(all="ignore"):
        l = (shapes["min"] - ray["origin"]) * ray["inv_direction"]
        r = (shapes["max"] - ray["origin"]) * ray["inv_direction"]
    nan = np.isnan(l).any(axis=1) | np.isnan(r).any(axis=1)
    tmin, tmax = np.minimum(l, r).max(axis=1), np.maximum(l, r).min(axis=1)
    return np.nonzero(~nan & (tmax >= np.where(tmin > 0, tmin, np.float32(0))))[0].tolist()


def test_update_shapes_bvh_12k_75p():                      # optimization.rs:639-662
    shapes = O.create_n_cubes(1000)
    b = O.build(shapes)
    assert O.is_consistent(b.nodes, shapes) and O.is_tight(b.nodes)
    moved, idx = _move_shapes(shapes, 9000, np.random.default_rng(0))
    assert not O.is_consistent(b.nodes, moved)
    nodes, node_index = O.update_shapes(b.nodes, b.node_index, moved, idx)
    assert len(nodes) == 2 * len(shapes) - 1
    assert O.is_consistent(nodes, moved) and O.is_tight(nodes)
    # and the updated (non-preorder) tree still answers ray queries like a brute-force scan of the moved shapes
    rays, _ = O.create_rays(64)
    r = O.traverse(nodes, moved, rays, O.MODE_RECURSIVE)
    for k, got in enumerate(O.per_ray_lists(r.offsets, r.hits)):
        assert sorted(got.tolist()) == _brute_force(moved, rays[k])


# ---------------------------------------------------------------------------------------------------------------------
# nearest_to (src/bvh/bvh_impl.rs:221-238, src/flat_bvh.rs:513-562)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("flat", [False, True], ids=["bvh", "flat"])
def test_nearest_to_doc_example(flat):                     # bvh_impl.rs / flat_bvh.rs doc tests: 1000 unit boxes on the diagonal
    pos = np.repeat(np.arange(1000, dtype=np.float32)[:, None], 3, axis=1)
    shapes = O.unit_boxes(pos)
    b = O.build(shapes)
    tree = O.flatten(b.nodes) if flat else b.nodes
    s, d = O.nearest_to(tree, shapes, [[5.0, 5.7, 5.3]], flat=flat)
    assert int(s[0]) == 5
    q = abs(np.float32(5.7) - np.float32(5.0)) - np.float32(0.5)                 # outside only along y
    assert d[0] == np.sqrt(q * q)


def test_aabb_min_distance_squared_doc():                  # aabb_impl.rs:598-614: the doc test's 10.0
    box = O.make_aabbs([[-1.0, -1.0, -1.0]], [[1.0, 1.0, 1.0]])
    assert np.sqrt(O.shape_distances_squared(box, [11.0, 0.0, 0.0])[0]) == 10.0
    assert O.shape_distances_squared(box, [0.5, -0.5, 0.0])[0] == 0.0            # inside


@pytest.mark.parametrize("flat", [False, True], ids=["bvh", "flat"])
def test_nearest_to_some_bh(flat):                         # testbase.rs:270-312 on 12 000 triangles
    shapes, tris = O.create_n_cubes(1000, want_tris=True)
    b = O.build(shapes)
    tree = O.flatten(b.nodes) if flat else b.nodes
    bounds = O.make_aabbs([[-1000.0] * 3], [[1000.0] * 3])
    ref_point, _ = O.next_points(1, bounds=bounds, seed=0)                       # the reference re-seeds with 0 for every query
    rng = np.random.default_rng(3)
    pts = np.concatenate([ref_point.reshape(1, 3), rng.uniform(-100000, 100000, (60, 3)).astype(np.float32),
                          tris.reshape(-1, 3)[rng.integers(0, len(tris) * 3, 20)]])          # some points ON the geometry
    s, d = O.nearest_to(tree, shapes, pts, flat=flat, kind=O.DIST_TRIANGLE, tris=tris)
    for k, p in enumerate(pts):
        d2 = O.shape_distances_squared(shapes, p, kind=O.DIST_TRIANGLE, tris=tris)
        assert d[k] == np.sqrt(d2.min()), k                                       # same arithmetic on both sides: exact
        assert d2[int(s[k])] == d2.min()


def test_nearest_to_empty_and_single():
    s, d = O.nearest_to(np.zeros(0, dtype=O.NODE3F), np.zeros(0, dtype=O.AABB3F), [[0.0, 0.0, 0.0]])
    assert int(s[0]) == 0xFFFFFFFF
    one = O.unit_boxes([[3.0, 0.0, 0.0]])
    b = O.build(one)
    s, d = O.nearest_to(b.nodes, one, [[0.0, 0.0, 0.0]])
    assert int(s[0]) == 0 and d[0] == 2.5


# ---------------------------------------------------------------------------------------------------------------------
# Property tests (the reference uses proptest for the same purpose, src/bvh/optimization.rs / src/testbase.rs fuzz helpers)
# ---------------------------------------------------------------------------------------------------------------------
from hypothesis import given, settings, strategies as st
from tests.scenes import rays_for


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 400), st.integers(0, 2**31 - 1), st.floats(0.0, 1.0), st.sampled_from([0.5, 50.0, 5000.0]))
def test_update_shapes_property(n, seed, frac, reach):
    """Any scene, any subset moved anywhere: update_shapes leaves a consistent, tight tree over all shapes with 2n-1 nodes, whose
    traversal agrees with a brute-force scan."""
    rng = np.random.default_rng(seed)
    mn = rng.uniform(-100, 100, (n, 3))
    shapes = O.make_aabbs(mn, mn + rng.uniform(0, 4, (n, 3)))
    b = O.build(shapes)
    m = int(round(n * frac))
    idx = rng.permutation(n)[:m].astype(np.uint32)
    moved = shapes.copy()
    d = rng.uniform(-reach, reach, (m, 3)).astype(np.float32)
    moved["min"][idx] += d
    moved["max"][idx] += d
    nodes, node_index = O.update_shapes(b.nodes, b.node_index, moved, idx)
    assert len(nodes) == 2 * n - 1
    assert O.is_consistent(nodes, moved) and O.is_tight(nodes)
    assert np.array_equal(nodes["shape"][node_index], np.arange(n)) and np.all(nodes["child_l"][node_index] == 0xFFFFFFFF)
    rays = rays_for(moved, 8, seed=seed % 1000)
    r = O.traverse(nodes, moved, rays, O.MODE_RECURSIVE)
    for k, got in enumerate(O.per_ray_lists(r.offsets, r.hits)):
        assert sorted(got.tolist()) == _brute_force(moved, rays[k])


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 600), st.integers(0, 2**31 - 1), st.booleans())
def test_nearest_to_property(n, seed, flat):
    """nearest_to returns a shape at the minimum AABB distance (brute force), for the Bvh and the FlatBvh walk."""
    rng = np.random.default_rng(seed)
    mn = rng.uniform(-50, 50, (n, 3))
    shapes = O.make_aabbs(mn, mn + rng.uniform(0, 6, (n, 3)))
    b = O.build(shapes)
    tree = O.flatten(b.nodes) if flat else b.nodes
    pts = rng.uniform(-80, 80, (12, 3)).astype(np.float32)
    s, d = O.nearest_to(tree, shapes, pts, flat=flat)
    for k in range(len(pts)):
        d2 = O.shape_distances_squared(shapes, pts[k])
        assert d2[int(s[k])] == d2.min() and d[k] == np.sqrt(d2.min())


# ---- Ray::intersects_triangle (src/ray/ray_impl.rs:154-213) -----------------------------------------------------------------
def test_oracle_ray_triangle_property_from_the_reference():
    """The reference's own property test (ray_impl.rs:361-420, test_ray_hits_triangle), restated: a ray aimed at a point u*AB + v*AC of a
    triangle hits it (distance < inf, u+v in [0,1]) unless it looks at the back face, in which case the distance is +inf."""
    rng = np.random.default_rng(1234)
    checked = 0
    for _ in range(4000):
        a, b, c, org = (rng.integers(-1000, 1000, 3).astype(np.float32) for _ in range(4))
        u = int(rng.integers(0, 101)); v = min(100 - u, int(rng.integers(0, 101)))
        uf, vf = np.float32(u / 100.0), np.float32(v / 100.0)
        uvec, vvec = b - a, c - a
        normal = np.cross(uvec, vvec)
        p = a + uf * uvec + vf * vvec
        if np.allclose(p, org):
            continue
        ray = O.ray_new((org,), (p - org,))
        on_back = float(np.dot(normal.astype(np.float64), (org - a).astype(np.float64))) <= 0.0
        dist, iu, iv = O.ray_triangle(ray, np.concatenate([a, b, c]))
        if on_back:
            if abs(float(np.dot(normal.astype(np.float64), (org - a).astype(np.float64)))) > 1e-3 * np.linalg.norm(normal):   # clearly behind
                assert np.isinf(dist)
        else:
            inside = 0.0 <= float(iu) + float(iv) <= 1.0 and np.isfinite(dist)
            eps = np.finfo(np.float32).eps
            border = abs(uf) < eps or abs(uf - 1) < eps or abs(vf) < eps or abs(vf - 1) < eps or abs(uf + vf - 1) < eps
            degenerate = np.linalg.norm(normal) < 1e-3
            assert inside or border or degenerate, (a, b, c, org, uf, vf, dist, iu, iv)
        checked += 1
    assert checked > 3000


def test_oracle_ray_triangle_kats():
    ray = O.ray_new(((0.0, 0.0, -1.0),), ((0.0, 0.0, 1.0),))
    d, u, v = O.ray_triangle(ray, [-1, -1, 0, 0, 1, 0, 1, -1, 0])        # front face: det > 0
    assert d == np.float32(1.0) and u == np.float32(0.5) and v == np.float32(0.25)
    d, u, v = O.ray_triangle(ray, [-1, -1, 0, 1, -1, 0, 0, 1, 0])        # same triangle, other winding: culled
    assert np.isinf(d)
    ray2 = O.ray_new(((5.0, 5.0, -1.0),), ((0.0, 0.0, 1.0),))            # misses
    assert np.isinf(O.ray_triangle(ray2, [-1, -1, 0, 0, 1, 0, 1, -1, 0])[0])


def test_oracle_closest_hit_equals_brute_force():
    shapes, tris = O.create_n_cubes(60, want_tris=True)
    res = O.build(shapes)
    rng = np.random.default_rng(3)
    centres = (shapes["min"][::12] + shapes["max"][::12]) * 0.5
    tgt = centres[rng.integers(0, len(centres), 400)].astype(np.float64) + rng.uniform(-0.4, 0.4, (400, 3))
    org = tgt + rng.normal(0, 1, (400, 3)) * 500
    rays = O.ray_new(org, tgt - org)
    s, d, _ = O.closest_hit(res.nodes, shapes, rays, tris)
    assert (s != O.U32_MAX).sum() > 300
    for i in range(0, 400, 7):                                             # brute force over ALL triangles
        best, bd = O.U32_MAX, np.inf
        for t in range(len(tris.reshape(-1, 9))):
            dist = O.ray_triangle(rays[i:i + 1], tris.reshape(-1, 9)[t])[0]
            if dist < bd:
                best, bd = t, dist
        assert s[i] == best and (np.isinf(bd) or d[i] == np.float32(bd))
