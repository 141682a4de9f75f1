"""GPU parity tests: libbvh_b200.so (through the C ABI) against the CPU oracle, bit-exact.
Run on the B200 box:  python -m pytest tests -m gpu"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.scenes import rays_for, scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from bvh_b200 import api as A

    return A


def _nodes_equal(a, b):
    """Bitwise on integers, float equality on AABBs (only the sign of a zero may differ)."""
    for f in ("parent", "child_l", "child_r", "shape"):
        if not np.array_equal(a[f], b[f]):
            return False
    for f in ("l_aabb", "r_aabb"):
        for g in ("min", "max"):
            if not np.array_equal(a[f][g], b[f][g]):
                return False
    return True


def _flat_equal(a, b):
    if len(a) != len(b):
        return False
    for f in ("entry_index", "exit_index", "shape_index"):
        if not np.array_equal(a[f], b[f]):
            return False
    return np.array_equal(a["aabb"]["min"], b["aabb"]["min"]) and np.array_equal(a["aabb"]["max"], b["aabb"]["max"])


SCENES_F32 = ["empty", "cubes1", "boxes21", "random2", "random3", "random31", "random32", "random33", "random255", "random256",
              "random257", "random513", "random1000", "random5000", "cubes100", "cubes1000", "points200", "points3000",
              "line700", "huge64", "huge2000", "skew40", "skew3000"]


@pytest.mark.parametrize("name", SCENES_F32)
def test_build_flatten_parity_f32(api, name):
    shapes = scene(name)
    want = O.build(shapes)
    bvh = api.Bvh.build(shapes)
    assert bvh.num_shapes == len(shapes)
    assert _nodes_equal(bvh.nodes, want.nodes), name
    assert np.array_equal(bvh.node_index, want.node_index)
    flat = bvh.flatten()
    assert _flat_equal(flat.nodes, O.flatten(want.nodes)), name
    if name.startswith("huge"):
        assert want.nosplit_fallthrough > 0
    if name.startswith("points"):
        assert want.degenerate_splits > 0
    bvh.free()


@pytest.mark.parametrize("name", ["empty", "cubes1", "boxes21", "random33", "random257", "random3000", "cubes200", "points500", "huge300", "skew500"])
def test_build_flatten_parity_f64(api, name):
    shapes = scene(name, "f64")
    want = O.build(shapes, "f64")
    bvh = api.Bvh.build(shapes, prec="f64")
    assert _nodes_equal(bvh.nodes, want.nodes), name
    assert np.array_equal(bvh.node_index, want.node_index)
    assert _flat_equal(bvh.flatten().nodes, O.flatten(want.nodes, "f64")), name
    bvh.free()


def _check_traverse(api, bvh, want_nodes, shapes, rays, prec="f32"):
    from bvh_b200 import capi

    flat = O.flatten(want_nodes, prec)
    r_rec = O.traverse(want_nodes, shapes, rays, O.MODE_RECURSIVE, prec)
    r_flat = O.traverse(flat, shapes, rays, O.MODE_FLAT, prec)
    for slots, pers in ((4, 2), (0, 1), (1, 0), (-1, 1)):   # single-pass / two-pass / forced overflow re-walk x pass-1 kernel choice
        bvh.ctx.set_option("traverse_slots", slots)
        bvh.ctx.set_option("traverse_persistent", pers)
        off, hits = bvh.traverse_batch(rays, mode=capi.TRAVERSE_BVH)
        assert np.array_equal(off.astype(np.uint64), r_rec.offsets), slots
        assert np.array_equal(hits, r_rec.hits), slots
        off, hits = bvh.traverse_batch(rays, mode=capi.TRAVERSE_FLAT)
        assert np.array_equal(off.astype(np.uint64), r_flat.offsets), slots
        assert np.array_equal(hits, r_flat.hits), slots
    bvh.ctx.set_option("traverse_slots", -1)
    bvh.ctx.set_option("traverse_persistent", 2)
    visits, total = bvh.traverse_stats()
    assert total == len(r_flat.hits)


@pytest.mark.parametrize("name", ["cubes1", "boxes21", "random2", "random257", "random5000", "cubes1000", "points3000", "huge2000", "skew3000"])
def test_traverse_parity_f32(api, name):
    shapes = scene(name)
    want = O.build(shapes)
    bvh = api.Bvh.build(shapes)
    rays = rays_for(shapes, 3000, seed=1, axis_aligned=300)
    _check_traverse(api, bvh, want.nodes, shapes, rays)
    bvh.free()


@pytest.mark.parametrize("name", ["boxes21", "random3000", "cubes200", "huge300"])
def test_traverse_parity_f64(api, name):
    shapes = scene(name, "f64")
    want = O.build(shapes, "f64")
    bvh = api.Bvh.build(shapes, prec="f64")
    rays = rays_for(shapes, 2000, "f64", seed=2, axis_aligned=200)
    _check_traverse(api, bvh, want.nodes, shapes, rays, "f64")
    bvh.free()


def test_empty_tree_and_empty_batch(api):
    e = scene("empty")
    bvh = api.Bvh.build(e)
    assert len(bvh.nodes) == 0 and len(bvh.flatten()) == 0
    rays = rays_for(e, 10)
    off, hits = bvh.traverse_batch(rays)
    assert off.tolist() == [0] * 11 and len(hits) == 0
    shapes = scene("boxes21")
    b2 = api.Bvh.build(shapes)
    off, hits = b2.traverse_batch(rays[:0])
    assert off.tolist() == [0] and len(hits) == 0


def test_reference_kats_through_the_api(api):
    """The reference's own fixed-scene tests (testbase.rs:174-225, bvh_impl.rs:665-690) driven through the product API."""
    import json, os

    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))

    class UnitBox:                                  # testbase.rs:66-103
        def __init__(self, id, pos):
            self.id, self.pos, self.node_index = id, np.array(pos, dtype=np.float32), 0

        def aabb(self):
            return self.pos + np.float32(-0.5), self.pos + np.float32(0.5)

        def set_bh_node_index(self, i):
            self.node_index = i

    boxes = [UnitBox(x, (x, 0.0, 0.0)) for x in range(-10, 11)]
    for builder in (api.Bvh.build, api.Bvh.build_par):
        bvh = builder(boxes)
        assert [b.node_index for b in boxes] == G["derived"]["aligned_boxes_21_node_index"]
        flat = bvh.flatten()
        for case in G["aligned_boxes_21"]["rays"]:
            ray = api.Ray.new([case["origin"]], [case["direction"]])
            for got in (bvh.traverse(ray, boxes), list(bvh.traverse_iterator(ray, boxes)), flat.traverse(ray, boxes)):
                assert sorted(b.id for b in got) == sorted(case["hit_ids"])
    for case in G["one_node_bvh"]["cases"]:
        one = [UnitBox(0, case["box_center"])]
        bvh = api.Bvh.build(one)
        ray = api.Ray.new([case["origin"]], [case["direction"]])
        assert len(bvh.traverse(ray, one)) == case["hits"]
        assert len(list(bvh.traverse_iterator(ray, one))) == case["hits"]
        assert len(bvh.flatten().traverse(ray, one)) == case["hits"]
    k = G["sah_pairing"]
    three = [UnitBox(i, c) for i, c in enumerate(k["box_centers"])]
    bvh = api.Bvh.build(three)
    a, b = k["same_parent"]
    assert bvh.nodes[three[a].node_index]["parent"] == bvh.nodes[three[b].node_index]["parent"]


def test_ray_new_parity(api):
    rng = np.random.default_rng(5)
    o = rng.uniform(-1e5, 1e5, (5000, 3))
    d = rng.uniform(-1e5, 1e5, (5000, 3))
    d[:50, 0] = 0.0
    d[50:60, 1:] = 0.0
    for prec in ("f32", "f64"):
        got = api.Ray.new(o, d, prec)
        want = O.ray_new(o, d, prec)
        for f in ("origin", "direction", "inv_direction"):
            assert np.array_equal(got[f], want[f], equal_nan=True), (prec, f)


def test_tree_from_nodes_roundtrip(api):
    shapes = scene("cubes100")
    want = O.build(shapes)
    bvh = api.Bvh.from_nodes(want.nodes, shapes)
    assert _nodes_equal(bvh.nodes, want.nodes)
    assert np.array_equal(bvh.node_index, want.node_index)
    assert _flat_equal(bvh.flatten().nodes, O.flatten(want.nodes))
    rays = rays_for(shapes, 2000, seed=3)
    _check_traverse(api, bvh, want.nodes, shapes, rays)
    # a node array that is not in preorder layout is refused, not mis-traversed
    from bvh_b200 import capi
    bad = want.nodes.copy()
    bad[0]["child_l"], bad[0]["child_r"] = bad[0]["child_r"], bad[0]["child_l"]
    with pytest.raises(capi.BvhGpuError):
        api.Bvh.from_nodes(bad, shapes)


def test_nan_input_is_an_error(api):
    from bvh_b200 import capi

    shapes = scene("random1000").copy()
    shapes["min"][123][1] = np.nan
    with pytest.raises(capi.BvhGpuError) as e:
        api.Bvh.build(shapes)
    assert e.value.status == capi.ERR_NAN


def test_sah_cost_matches_oracle(api):
    shapes = scene("cubes1000")
    want = O.build(shapes)
    bvh = api.Bvh.build(shapes)
    got, ref = bvh.sah_cost(), O.sah_cost(want.nodes)
    assert got[0] == pytest.approx(ref[0], rel=1e-12) and got[1] == pytest.approx(ref[1], rel=1e-12)


def test_refit_keeps_the_tree_valid(api):
    """bvhgpu_refit = the data-parallel part of update_shapes (optimization.rs:304-351): after moving
    shapes the tree must be consistent and tight again (optimization.rs:648-669 asserts exactly that)."""
    shapes = scene("cubes1000").copy()
    bvh = api.Bvh.build(shapes)
    assert O.is_consistent(bvh.nodes, shapes) and O.is_tight(bvh.nodes)
    rng = np.random.default_rng(9)
    moved = rng.choice(len(shapes), 9000, replace=False)
    delta = rng.uniform(-5000, 5000, (9000, 3)).astype(np.float32)
    shapes["min"][moved] += delta
    shapes["max"][moved] += delta
    assert not O.is_consistent(bvh.nodes, shapes)
    topo = bvh.nodes[["parent", "child_l", "child_r", "shape"]].copy()
    bvh.refit(shapes)
    assert O.is_consistent(bvh.nodes, shapes) and O.is_tight(bvh.nodes)
    assert np.array_equal(bvh.nodes[["parent", "child_l", "child_r", "shape"]], topo)
    # traversal of the refitted tree == oracle traversal of the same node array
    rays = rays_for(shapes, 2000, seed=4)
    _check_traverse(api, bvh, bvh.nodes, shapes, rays)


def _preorder_layout_ok(nodes):
    """Bvh::build's layout rule (bvh_node.rs:138-142): child_l = i + 1, child_r = i + 2 * (shapes under the left child)."""
    inner = np.nonzero(nodes["child_l"] != 0xFFFFFFFF)[0]
    cnt = np.where(nodes["child_l"] == 0xFFFFFFFF, 1, nodes["shape"]).astype(np.int64)
    l, r = nodes["child_l"][inner].astype(np.int64), nodes["child_r"][inner].astype(np.int64)
    return bool(np.all(l == inner + 1) and np.all(r == inner + 2 * cnt[l]) and np.all(cnt[inner] == cnt[l] + cnt[r])
                and np.all(nodes["parent"][l] == inner) and np.all(nodes["parent"][r] == inner))


@pytest.mark.parametrize("name,prec,frac,offset", [("cubes1000", "f32", 0.01, 10.0), ("cubes1000", "f32", 0.10, 10.0), ("cubes1000", "f32", 0.75, None),
                                                   ("random5000", "f32", 0.05, 300.0), ("points3000", "f32", 0.2, 3.0),
                                                   ("random3000", "f64", 0.1, 200.0), ("cubes200", "f64", 0.5, None)])
def test_optimize_after_shapes_moved(api, name, prec, frac, offset):
    """bvhgpu_optimize = the counterpart of Bvh::update_shapes (optimization.rs:290-302): refit + in-place exact rebuild of the
    degraded subtrees.  Not the reference's tree (that is a sequential re-insertion), so parity is on the reference's own
    acceptance criteria (optimization.rs:639-662: consistent and tight), on the layout rule, on hit sets, and on SAH cost
    against the oracle's update_shapes run on the same motion."""
    shapes = scene(name, prec).copy()
    F = shapes["min"].dtype
    bvh = api.Bvh.build(shapes, prec=prec)
    built = O.build(shapes, prec)
    rng = np.random.default_rng(5)
    m = max(1, int(len(shapes) * frac))
    moved = rng.choice(len(shapes), m, replace=False)
    ext = float(shapes["max"].max() - shapes["min"].min())
    delta = rng.uniform(-(offset or ext / 2), (offset or ext / 2), (m, 3)).astype(F)
    shapes["min"][moved] += delta
    shapes["max"][moved] += delta
    # reference path: sequential re-insertion on the oracle
    ref_nodes, _ = O.update_shapes(built.nodes, built.node_index, shapes, moved, prec)
    assert O.is_consistent(ref_nodes, shapes, prec)
    c_ref = O.sah_cost(ref_nodes, prec)[0]
    # pure refit (on a second tree: optimize measures growth against the tree's state before the call), for comparison
    other = api.Bvh.build(scene(name, prec), prec=prec)
    other.refit(shapes)
    c_refit = other.sah_cost()[0]
    other.free()
    rebuilt = bvh.optimize(shapes, 1.5)
    nodes, node_index = bvh.nodes, bvh.node_index
    assert 0 < rebuilt <= len(shapes)
    assert _preorder_layout_ok(nodes)
    assert O.is_consistent(nodes, shapes, prec) and O.is_tight(nodes, prec)
    assert np.array_equal(nodes["shape"][node_index], np.arange(len(shapes)))           # set_bh_node_index targets
    assert np.all(nodes["child_l"][node_index] == 0xFFFFFFFF)
    c_opt = bvh.sah_cost()[0]
    c_fresh = O.sah_cost(O.build(shapes, prec).nodes, prec)[0]
    assert c_opt <= c_refit * (1 + 1e-9), (c_opt, c_refit)
    assert c_opt <= 1.10 * c_ref, (c_opt, c_ref, c_fresh)                               # at least as good as the reference's update
    rays = rays_for(shapes, 2000, seed=6, prec=prec)
    _check_traverse(api, bvh, nodes, shapes, rays, prec=prec)
    assert _flat_equal(bvh.flatten().nodes, O.flatten(nodes, prec))
    bvh.free()


@pytest.mark.parametrize("pct", [1, 10, 50])
def test_optimize_vs_reference_update_120k(api, pct):
    """The reference's update benchmark shape (optimization.rs:693-725: 120 k triangles, p % moved by at most 10.0): the oracle's
    update_shapes and bvhgpu_optimize on the same motion; SAH cost of the result and wall time of both are recorded in
    gpurun_out/optimize_vs_reference.json (the timing is informative, the assertions are on validity and cost)."""
    import json, os, time
    from bvh_b200 import scenes as S
    a = S.create_n_cubes_aabbs(10000)
    ob = O.build(a)
    rng = np.random.default_rng(pct)
    mv = rng.choice(len(a), len(a) * pct // 100, replace=False)
    am = a.copy()
    dl = rng.uniform(-10.0, 10.0, (len(mv), 3)).astype(np.float32)
    am["min"][mv] += dl
    am["max"][mv] += dl
    t0 = time.perf_counter()
    rn, _ = O.update_shapes(ob.nodes, ob.node_index, am, mv)
    t_ref = (time.perf_counter() - t0) * 1e3
    assert O.is_consistent(rn, am) and O.is_tight(rn)
    g = api.Bvh.build(a)
    g.ctx.synchronize()
    t0 = time.perf_counter()
    rebuilt = g.optimize(am, 1.5)
    t_gpu = (time.perf_counter() - t0) * 1e3
    nodes = g.nodes
    assert O.is_consistent(nodes, am) and O.is_tight(nodes) and _preorder_layout_ok(nodes)
    c_ref, c_opt = O.sah_cost(rn)[0], g.sah_cost()[0]
    c_fresh = O.sah_cost(O.build(am, threads=O.hardware_threads()).nodes)[0]
    assert c_opt <= 1.10 * c_ref, (c_opt, c_ref)
    g.free()
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "optimize_vs_reference.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    rec = json.load(open(path)) if os.path.exists(path) else {}
    rec[f"120k_f32_{pct}pct"] = {"moved": int(len(mv)), "oracle_update_shapes_ms_1thread": t_ref, "gpu_optimize_host_call_ms": t_gpu,
                                 "rebuilt_shapes": int(rebuilt), "sah_cost_oracle_update_shapes": c_ref, "sah_cost_gpu_optimize": c_opt,
                                 "sah_cost_fresh_build": c_fresh}
    json.dump(rec, open(path, "w"), indent=1)


# ---- nearest_to ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,prec", [("cubes1000", "f32"), ("random5000", "f32"), ("points3000", "f32"), ("skew3000", "f32"), ("boxes21", "f32"),
                                       ("cubes1", "f32"), ("random3000", "f64"), ("huge300", "f64")])
def test_nearest_to_matches_the_reference_walk(api, name, prec):
    """bvhgpu_nearest_* replays Bvh::nearest_to / FlatBvh::nearest_to for AABB-distance shapes: same shape (ties included: the walk
    order is the reference's) and bit-identical distance as the oracle, in both visiting orders."""
    from bvh_b200 import capi
    shapes = scene(name, prec)
    bvh = api.Bvh.build(shapes, prec=prec)
    want = O.build(shapes, prec)
    rng = np.random.default_rng(12)
    lo, hi = shapes["min"].min(axis=0).astype(float), shapes["max"].max(axis=0).astype(float)
    pts = rng.uniform(lo - (hi - lo) * 0.2 - 1, hi + (hi - lo) * 0.2 + 1, (3000, 3))
    pts[:200] = (shapes["min"][rng.integers(0, len(shapes), 200)] + shapes["max"][rng.integers(0, len(shapes), 200)]) * 0.5   # inside / on boxes: ties at 0
    for flat, mode in ((False, capi.TRAVERSE_BVH), (True, capi.TRAVERSE_FLAT)):
        tree = O.flatten(want.nodes, prec) if flat else want.nodes
        ws, wd = O.nearest_to(tree, shapes, pts, prec, flat=flat)
        gs, gd = bvh.nearest_to_batch(pts, mode=mode)
        assert np.array_equal(gs, ws), (name, flat, int(np.sum(gs != ws)))
        assert np.array_equal(gd, wd), (name, flat)
    bvh.free()


def test_nearest_to_empty_tree(api):
    bvh = api.Bvh.build(scene("empty"))
    s, d = bvh.nearest_to_batch([[0.0, 0.0, 0.0], [1.0, 2.0, 3.0]])
    assert np.all(s == 0xFFFFFFFF) and np.all(d == 0)
    off, cand = bvh.nearest_candidates([[0.0, 0.0, 0.0]])
    assert off.tolist() == [0, 0] and len(cand) == 0


def test_nearest_to_doc_example_and_some_bh(api):
    """The reference's doc example (1000 unit boxes on the diagonal, query (5, 5.7, 5.3) -> shape 5) and nearest_to_some_bh
    (testbase.rs:270-312: 12 000 triangles with their own PointDistance, brute force as the judge): the candidate lists from
    the device plus the shape's distance function on the host give the brute-force answer."""
    pos = np.repeat(np.arange(1000, dtype=np.float32)[:, None], 3, axis=1)
    boxes = O.unit_boxes(pos)
    b = api.Bvh.build(boxes)
    s, d = b.nearest_to_batch([[5.0, 5.7, 5.3]])
    assert int(s[0]) == 5
    b.free()
    shapes, tris = O.create_n_cubes(1000, want_tris=True)
    bvh = api.Bvh.build(shapes)
    bounds = O.make_aabbs([[-1000.0] * 3], [[1000.0] * 3])
    ref_point, _ = O.next_points(1, bounds=bounds, seed=0)
    rng = np.random.default_rng(3)
    pts = np.concatenate([ref_point.reshape(1, 3), rng.uniform(-100000, 100000, (300, 3)).astype(np.float32),
                          tris.reshape(-1, 3)[rng.integers(0, len(tris) * 3, 50)]])
    off, cand = bvh.nearest_candidates(pts)
    assert np.all(np.diff(off.astype(np.int64)) >= 1)
    sizes = np.diff(off.astype(np.int64))
    assert sizes.mean() < 200, sizes.mean()                                       # short lists: this is a pruned search, not a scan
    for k, p in enumerate(pts):
        d2 = O.shape_distances_squared(shapes, p, kind=O.DIST_TRIANGLE, tris=tris)          # PointDistance of every triangle
        mine = cand[off[k]:off[k + 1]]
        assert d2[mine].min() == d2.min(), k
    bvh.free()


def test_optimize_large_scene_without_gangs(api):
    """1.2 M shapes: above the gang threshold, so rebuild roots larger than 512 shapes run as queue-mode BIN / SCATTER tile tasks."""
    from bvh_b200 import scenes as S
    shapes = S.create_n_cubes_aabbs(100000).copy()
    bvh = api.Bvh.build(shapes)
    rng = np.random.default_rng(21)
    mv = rng.choice(len(shapes), len(shapes) // 20, replace=False)
    dl = rng.uniform(-3000.0, 3000.0, (len(mv), 3)).astype(np.float32)
    shapes["min"][mv] += dl
    shapes["max"][mv] += dl
    rebuilt = bvh.optimize(shapes, 1.5)
    nodes = bvh.nodes
    assert 512 < rebuilt <= len(shapes)
    assert _preorder_layout_ok(nodes)
    assert O.is_consistent(nodes, shapes) and O.is_tight(nodes)
    assert np.array_equal(nodes["shape"][bvh.node_index], np.arange(len(shapes)))
    rays = rays_for(shapes, 20000, seed=2)
    r = O.traverse(nodes, shapes, rays, O.MODE_RECURSIVE, threads=O.hardware_threads())
    off, hits = bvh.traverse_batch(rays)
    assert np.array_equal(off.astype(np.uint64), r.offsets) and np.array_equal(hits, r.hits)
    bvh.free()


def test_optimize_without_motion_is_a_no_op(api):
    shapes = scene("random5000")
    bvh = api.Bvh.build(shapes)
    before, idx = bvh.nodes.copy(), bvh.node_index.copy()
    assert bvh.optimize(shapes, 1.5) == 0
    assert _nodes_equal(bvh.nodes, before) and np.array_equal(bvh.node_index, idx)
    bvh.free()


def test_optimize_global_motion_is_a_full_rebuild(api):
    """Everything moved far: the tree root itself is degraded, the rebuild root is node 0, and the result is Bvh::build of the
    moved shapes (general-position scene: the only order dependence of the builder, the degenerate halving, does not occur)."""
    shapes = scene("random5000").copy()
    bvh = api.Bvh.build(shapes)
    rng = np.random.default_rng(8)
    delta = rng.uniform(-20000, 20000, (len(shapes), 3)).astype(np.float32)
    shapes["min"] += delta
    shapes["max"] += delta
    assert bvh.optimize(shapes, 1.5) == len(shapes)
    want = O.build(shapes)
    assert _nodes_equal(bvh.nodes, want.nodes) and np.array_equal(bvh.node_index, want.node_index)
    bvh.free()


def test_config2_full_size(api):
    """BASELINE.json configs[1]: 120k-triangle scene, 1M create_ray rays: bit-exact build + flatten, identical
    hit lists for all 1M rays (the oracle needs a few seconds for this)."""
    shapes = O.create_n_cubes(10_000)
    want = O.build(shapes)
    bvh = api.Bvh.build(shapes)
    assert _nodes_equal(bvh.nodes, want.nodes)
    assert np.array_equal(bvh.node_index, want.node_index)
    flat = O.flatten(want.nodes)
    assert _flat_equal(bvh.flatten().nodes, flat)
    rays, _ = O.create_rays(1_000_000)
    r = O.traverse(flat, shapes, rays, O.MODE_FLAT, threads=O.hardware_threads())
    off, hits = bvh.traverse_batch(rays)
    assert np.array_equal(off.astype(np.uint64), r.offsets) and np.array_equal(hits, r.hits)
    # create_ray shares seed 0 with create_n_cubes, so ray k starts at the centre of cube 2k while 2k < 10 000 and
    # hits exactly the two triangles of one face; later rays hit nothing in this sparse scene (derived).
    assert len(hits) == 10_000 and np.all((off[1:] - off[:-1])[:5000] == 2)


def test_large_build_properties(api):
    """1.2M shapes (beyond what the reference benches; this size also runs the thread-per-range kernel for the bottom of
    the tree): GPU == oracle, plus the size-independent invariants."""
    shapes = O.create_n_cubes(100_000)
    bvh = api.Bvh.build(shapes)
    nodes, idx = bvh.nodes, bvh.node_index
    n = len(shapes)
    leaves = nodes["child_l"] == O.U32_MAX
    assert leaves.sum() == n and np.array_equal(np.sort(nodes["shape"][leaves]), np.arange(n))
    assert np.array_equal(nodes["shape"][idx], np.arange(n))
    assert O.is_consistent(nodes, shapes) and O.is_tight(nodes)
    want = O.build(shapes, threads=O.hardware_threads())
    assert _nodes_equal(nodes, want.nodes)


def test_device_pointer_entry_points(api):
    import torch

    from bvh_b200 import capi
    from bvh_b200.dtypes import AABB3F, RAY3F

    shapes = scene("cubes1000")
    want = O.build(shapes)
    rays = rays_for(shapes, 5000, seed=6)
    dev = torch.device("cuda", 0)
    ctx = api.Context.default()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        d_aabb = torch.from_numpy(shapes.view(np.uint8).reshape(-1)).to(dev)
        d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(dev)
        d_off = torch.empty(len(rays) + 1, dtype=torch.int32, device=dev)
        d_hits = torch.empty(16 * len(rays), dtype=torch.int32, device=dev)
        bvh = api.Bvh.build_dev(d_aabb.data_ptr(), len(shapes))
        total = bvh.traverse_dev(d_rays.data_ptr(), len(rays), d_off.data_ptr(), d_hits.data_ptr(), d_hits.numel(), want_total=True)
        r = O.traverse(want.nodes, shapes, rays, O.MODE_RECURSIVE)
        assert total == len(r.hits)
        assert np.array_equal(d_off.cpu().numpy().view(np.uint32).astype(np.uint64), r.offsets)
        assert np.array_equal(d_hits[:total].cpu().numpy().view(np.uint32), r.hits)
        assert _nodes_equal(bvh.nodes, want.nodes)
    finally:
        ctx.set_stream(None)


# ---- the other BASELINE.json configurations as parity cases -------------------------------------------------------
def test_sponza_build_and_coherent_rays(api):
    """configs[2]: Sponza (66 450 triangles through the reference loader), 2048 x 2048 coherent pinhole rays."""
    from bvh_b200 import scenes as S
    from tests.scenes import sponza

    shapes = sponza()
    assert len(shapes) == 66_450
    want = O.build(shapes, threads=O.hardware_threads())
    bvh = api.Bvh.build(shapes)
    assert _nodes_equal(bvh.nodes, want.nodes) and np.array_equal(bvh.node_index, want.node_index)
    assert _flat_equal(bvh.flatten().nodes, O.flatten(want.nodes))
    o, d = S.pinhole_rays(2048, 2048)
    rays = api.Ray.new(o, d)
    assert np.array_equal(rays["inv_direction"], O.ray_new(o, d)["inv_direction"], equal_nan=True)
    r = O.traverse(want.nodes, shapes, rays, O.MODE_RECURSIVE, threads=O.hardware_threads())
    off, hits = bvh.traverse_batch(rays)
    assert np.array_equal(off.astype(np.uint64), r.offsets) and np.array_equal(hits, r.hits)
    assert len(hits) > 4 * len(rays)                 # dense scene: many hits per ray, exercises the > K-slots re-walk


def test_sponza_incoherent_shard(api):
    """configs[3]: Sponza, create_ray rays inside the scene bounds (testbase.rs:628-631, 687-691): one 2 M-ray shard."""
    from bvh_b200 import scenes as S
    from tests.scenes import sponza

    shapes = sponza()
    bmin, bmax = shapes["min"].min(axis=0), shapes["max"].max(axis=0)
    want = O.build(shapes, threads=O.hardware_threads())
    bvh = api.Bvh.build(shapes)
    first = 6_000_000                                  # the shard rank 3 of 8 would own
    o, d = S.ray_endpoints(2_000_000, first_ray=first, bounds=(bmin, bmax))
    rays = api.Ray.new(o, d)
    r = O.traverse(want.nodes, shapes, rays, O.MODE_RECURSIVE, threads=O.hardware_threads())
    off, hits = bvh.traverse_batch(rays)
    assert np.array_equal(off.astype(np.uint64), r.offsets) and np.array_equal(hits, r.hits)
    # the vectorised chain == the scalar chain for a shard that does not start at ray 0
    b = np.zeros(1, dtype=O.AABB3F); b["min"], b["max"] = bmin, bmax
    seed_rays, _ = O.create_rays(1000, bounds=b)
    o0, d0 = S.ray_endpoints(1000, 0, bounds=(bmin, bmax))
    assert np.array_equal(seed_rays["origin"], o0)


def test_config5_ten_million_f64(api):
    """configs[4]: 10 M triangles, f64: full exact-SAH build bit-identical to the oracle, SAH cost equal,
    then move 1 % of the shapes, refit, and check the reference's consistency + tightness invariants."""
    from bvh_b200 import scenes as S

    shapes = S.create_n_cubes_aabbs(833_334, "f64")[:10_000_000]
    want = O.build(shapes, "f64", threads=O.hardware_threads())
    bvh = api.Bvh.build(shapes, prec="f64")
    nodes = bvh.nodes
    assert _nodes_equal(nodes, want.nodes)
    assert np.array_equal(bvh.node_index, want.node_index)
    got, ref = bvh.sah_cost(), O.sah_cost(want.nodes, "f64")
    assert got[0] == pytest.approx(ref[0], rel=1e-9) and got[1] == pytest.approx(ref[1], rel=1e-9)
    rng = np.random.default_rng(11)
    moved = rng.choice(len(shapes), len(shapes) // 100, replace=False)
    delta = rng.uniform(-10.0, 10.0, (len(moved), 3))          # max offset 10.0 as in optimization.rs:702
    shapes = shapes.copy()
    shapes["min"][moved] += delta
    shapes["max"][moved] += delta
    bvh.refit(shapes)
    nodes = bvh.nodes
    assert O.is_consistent(nodes, shapes, "f64") and O.is_tight(nodes, "f64")
    # SAH cost after refit vs a fresh oracle rebuild of the moved scene: stated tolerance 10 %
    rebuilt = O.build(shapes, "f64", threads=O.hardware_threads())
    c_refit, c_rebuild = bvh.sah_cost()[0], O.sah_cost(rebuilt.nodes, "f64")[0]
    assert c_refit <= 1.10 * c_rebuild, (c_refit, c_rebuild)
    # a second motion, this time fixed by bvhgpu_optimize (refit + in-place rebuild; queue-mode tile tasks and the thread-per-range
    # kernel at this size): same acceptance criteria, cost no worse than the refit-only tree
    moved2 = rng.choice(len(shapes), len(shapes) // 100, replace=False)
    shapes["min"][moved2] += 25.0
    shapes["max"][moved2] += 25.0
    n_rebuilt = bvh.optimize(shapes, 1.5)
    nodes = bvh.nodes
    assert 0 < n_rebuilt <= len(shapes)
    assert _preorder_layout_ok(nodes)
    assert O.is_consistent(nodes, shapes, "f64") and O.is_tight(nodes, "f64")
    assert np.array_equal(nodes["shape"][bvh.node_index], np.arange(len(shapes)))


# ---- BVHGPU_BUILD_LBVH: different topology, same layout rule, same hit sets ------------------------------------------
@pytest.mark.parametrize("mode", [1, 2], ids=["lbvh", "lbvh_treelet"])
@pytest.mark.parametrize("name", ["cubes1", "random2", "random3", "random33", "random257", "random1000", "cubes1000", "points3000", "line700", "skew3000", "huge300"])
def test_lbvh_mode_is_a_valid_reference_layout_bvh(api, name, mode):
    from bvh_b200 import capi

    shapes = scene(name)
    bvh = api.Bvh.build(shapes, mode=mode)
    if mode == capi.BUILD_LBVH_TREELET and len(shapes) <= 512 and name.startswith("random"):
        # a scene that fits one treelet is rebuilt entirely by the SAH kernel: identical to the exact builder (for shapes in
        # general position; the degenerate "halve by position" branch sees the Morton order instead of the input order)
        assert _nodes_equal(bvh.nodes, O.build(shapes).nodes)
    nodes, idx = bvh.nodes, bvh.node_index
    n = len(shapes)
    assert len(nodes) == 2 * n - 1
    leaves = nodes["child_l"] == O.U32_MAX
    assert leaves.sum() == n and np.array_equal(np.sort(nodes["shape"][leaves]), np.arange(n))
    assert np.array_equal(nodes["shape"][idx], np.arange(n))
    if not (name.startswith("huge") and mode == capi.BUILD_LBVH_TREELET):   # ("no split wins" SAH nodes store empty child AABBs, as the reference does)
        assert O.is_consistent(nodes, shapes) and O.is_tight(nodes)      # the reference's own acceptance checks
    # preorder rule child_l = i+1, child_r = i + 2*n_l: the host-side validator of tree_from_nodes accepts it
    again = api.Bvh.from_nodes(nodes, shapes)
    assert np.array_equal(again.node_index, idx)
    # flatten() of that tree == the literal reference recursion applied to the same node array
    assert _flat_equal(bvh.flatten().nodes, O.flatten(nodes))
    # traversal of the LBVH tree == oracle traversal of the same node array (sequences), and == the exact-SAH tree (sets)
    rays = rays_for(shapes, 2000, seed=8)
    r = O.traverse(nodes, shapes, rays, O.MODE_RECURSIVE)
    off, hits = bvh.traverse_batch(rays)
    assert np.array_equal(off.astype(np.uint64), r.offsets) and np.array_equal(hits, r.hits)
    if not name.startswith("huge"):
        ref = O.build(shapes)
        rr = O.traverse(ref.nodes, shapes, rays, O.MODE_RECURSIVE)
        for a, b in zip(O.per_ray_lists(r.offsets, r.hits), O.per_ray_lists(rr.offsets, rr.hits)):
            assert sorted(a.tolist()) == sorted(b.tolist())


def test_lbvh_sah_cost_ratio_config2(api):
    """every lbvh result reports C_gpu / C_oracle (stated tolerance: informational in round 1)."""
    from bvh_b200 import capi

    shapes = O.create_n_cubes(10_000)
    exact = api.Bvh.build(shapes)
    lbvh = api.Bvh.build(shapes, mode=capi.BUILD_LBVH)
    tre = api.Bvh.build(shapes, mode=capi.BUILD_LBVH_TREELET)
    ce, cl, ct = exact.sah_cost(), lbvh.sah_cost(), tre.sah_cost()
    print(f"SAH cost (pseudo-area) exact {ce[0]:.4f} lbvh {cl[0]:.4f} ratio {cl[0] / ce[0]:.3f}; geometric ratio {cl[1] / ce[1]:.3f}; "
          f"lbvh+treelet ratio {ct[0] / ce[0]:.3f} (geometric {ct[1] / ce[1]:.3f})")
    assert cl[0] / ce[0] < 3.0
    assert ct[0] / ce[0] <= 1.10                           # stated tolerance
    rays, _ = O.create_rays(200_000)
    a = exact.traverse_batch(rays)
    b = lbvh.traverse_batch(rays)
    assert np.array_equal(a[0], b[0])                    # same hit counts per ray
    for x, y in zip(O.per_ray_lists(a[0], a[1])[:5000], O.per_ray_lists(b[0], b[1])[:5000]):
        assert sorted(x.tolist()) == sorted(y.tolist())


# ---- Aabb / Point / Ball queries (the other IntersectsAabb implementors) ------------------------------
def test_query_kats_from_the_reference(api):
    """testbase.rs:227-266: point (0,0,0) -> {0}; point (0,1000,0) -> {}; aabb [5.1,-1,-1]..[9.9,1,1] -> {5..10};
    sphere c=(5,-1,-1) r=1.4 -> {4,5,6}; through Bvh and FlatBvh semantics."""
    from bvh_b200 import capi

    boxes = O.aligned_boxes()
    bvh = api.Bvh.build(boxes)
    cases = [(capi.QUERY_POINT, [[0, 0, 0]], {0}), (capi.QUERY_POINT, [[0, 1000, 0]], set()),
             (capi.QUERY_AABB, [[5.1, -1, -1, 9.9, 1, 1]], set(range(5, 11))), (capi.QUERY_BALL, [[5, -1, -1, 1.4]], {4, 5, 6})]
    for kind, q, want in cases:
        for mode in (capi.TRAVERSE_BVH, capi.TRAVERSE_FLAT):
            _, hits = bvh.query_batch(kind, q, mode)
            assert sorted(int(h) - 10 for h in hits) == sorted(want)


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("name", ["boxes21", "cubes500", "random3000", "points500", "huge300"])
def test_query_parity(api, name, prec):
    from bvh_b200 import capi

    shapes = scene(name, prec)
    want = O.build(shapes, prec)
    flat = O.flatten(want.nodes, prec)
    bvh = api.Bvh.build(shapes, prec=prec)
    rng = np.random.default_rng(12)
    lo, hi = shapes["min"].min(axis=0).astype(float), shapes["max"].max(axis=0).astype(float)
    ext = hi - lo + 1e-3
    n = 3000
    pts = rng.uniform(lo - 0.05 * ext, hi + 0.05 * ext, (n, 3))
    pts[:200] = shapes["min"][rng.integers(0, len(shapes), 200)]          # exactly on box corners: >= / <= edge cases
    amin = rng.uniform(lo, hi, (n, 3))
    aab = np.concatenate([amin, amin + rng.uniform(0, 0.2, (n, 3)) * ext], axis=1)
    balls = np.concatenate([rng.uniform(lo, hi, (n, 3)), (rng.uniform(0, 0.15, (n, 1)) * ext.max())], axis=1)
    for kind, q in ((capi.QUERY_POINT, pts), (capi.QUERY_AABB, aab), (capi.QUERY_BALL, balls)):
        for mode, fl in ((capi.TRAVERSE_BVH, None), (capi.TRAVERSE_FLAT, flat)):
            off, hits = bvh.query_batch(kind, q, mode)
            woff, whits = O.query(kind, q, want.nodes, shapes, fl, prec)
            assert np.array_equal(off.astype(np.uint64), woff) and np.array_equal(hits, whits), (kind, mode)


# ---- randomized sweep (in the spirit of the reference's fuzz target, fuzz/fuzz_targets/fuzz.rs) ----------------------
@pytest.mark.parametrize("seed", range(24))
def test_random_scene_sweep(api, seed):
    """Random size (1..6000), random distribution family, random precision: build + flatten + traverse + point query parity."""
    from bvh_b200 import capi

    rng = np.random.default_rng(1000 + seed)
    prec = "f32" if seed % 3 else "f64"
    n = int(rng.choice([1, 2, 3, 5, 31, 32, 33, 64, 255, 256, 257, 511, 513, int(rng.integers(600, 6000))]))
    fam = seed % 4
    if fam == 0:        # uniform boxes
        mn = rng.uniform(-100, 100, (n, 3)); size = rng.uniform(0, 5, (n, 3))
    elif fam == 1:      # integer grid with many exact ties (Grid mode of the fuzzer: coordinates in thirds)
        mn = rng.integers(-12, 12, (n, 3)).astype(float) / 3.0; size = rng.integers(0, 3, (n, 3)).astype(float) / 3.0
    elif fam == 2:      # clustered, wildly different scales
        c = rng.uniform(-1e4, 1e4, (max(n // 50, 1), 3)); mn = c[rng.integers(0, len(c), n)] + rng.normal(0, 1, (n, 3)) * 10.0 ** rng.uniform(-3, 2, (n, 1)); size = 10.0 ** rng.uniform(-4, 1, (n, 3))
    else:               # flat sheet: one axis degenerate
        mn = rng.uniform(-50, 50, (n, 3)); mn[:, int(rng.integers(0, 3))] = 7.0; size = rng.uniform(0, 2, (n, 3)); size[:, 1] = 0.0
    shapes = O.make_aabbs(mn, mn + size, prec)
    want = O.build(shapes, prec)
    bvh = api.Bvh.build(shapes, prec=prec)
    assert _nodes_equal(bvh.nodes, want.nodes), (seed, n, fam, prec)
    assert np.array_equal(bvh.node_index, want.node_index)
    assert _flat_equal(bvh.flatten().nodes, O.flatten(want.nodes, prec))
    rays = rays_for(shapes, 1500, prec, seed=seed, axis_aligned=150)
    r = O.traverse(want.nodes, shapes, rays, O.MODE_RECURSIVE, prec)
    off, hits = bvh.traverse_batch(rays)
    assert np.array_equal(off.astype(np.uint64), r.offsets) and np.array_equal(hits, r.hits), (seed, n, fam, prec)
    pts = rng.uniform(mn.min(axis=0), (mn + size).max(axis=0), (500, 3))
    o2, h2 = bvh.query_batch(capi.QUERY_POINT, pts)
    wo, wh = O.query(O.QUERY_POINT, pts, want.nodes, shapes, None, prec)
    assert np.array_equal(o2.astype(np.uint64), wo) and np.array_equal(h2, wh)
    bvh.free()


@pytest.mark.parametrize("name,prec", [("cubes1000", "f32"), ("random5000", "f32"), ("points3000", "f32"), ("huge2000", "f32"), ("skew3000", "f32"),
                                       ("random33", "f32"), ("cubes200", "f64"), ("random3000", "f64"), ("huge300", "f64")])
def test_thread_per_range_kernel_forced(api, name, prec):
    """The bottom-of-tree kernel (one thread per range of <= 16 shapes) is normally used from 400 k shapes up; force it on
    small scenes, including the degenerate and the "no split wins" branches."""
    shapes = scene(name, prec)
    want = O.build(shapes, prec)
    ctx = api.Context.default()
    ctx.set_option("build_small", 1)
    try:
        bvh = api.Bvh.build(shapes, prec=prec)
        assert _nodes_equal(bvh.nodes, want.nodes), name
        assert np.array_equal(bvh.node_index, want.node_index)
        assert _flat_equal(bvh.flatten().nodes, O.flatten(want.nodes, prec))
    finally:
        ctx.set_option("build_small", -1)


@pytest.mark.parametrize("small,subtree,gang", [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 1, 1), (1, 1, 1), (1, 0, 1)])
@pytest.mark.parametrize("name,prec", [("cubes1000", "f32"), ("random5000", "f32"), ("points3000", "f32"), ("huge2000", "f32"), ("skew3000", "f32"),
                                       ("random33", "f32"), ("random700", "f32"), ("line2000", "f32"), ("cubes200", "f64"), ("random3000", "f64"), ("huge300", "f64")])
def test_builder_strategies_are_bit_identical(api, name, prec, small, subtree, gang):
    """The exact builder picks its strategies by size and type (warp gangs for the top levels, in-register subtrees or the
    thread-per-range kernel for the bottom); every combination must produce the reference's bits, on the degenerate
    ("halve by position") and "no split wins" scenes too."""
    shapes = scene(name, prec)
    want = O.build(shapes, prec)
    ctx = api.Context.default()
    ctx.set_option("build_small", small); ctx.set_option("build_subtree", subtree); ctx.set_option("build_gang", gang)
    try:
        bvh = api.Bvh.build(shapes, prec=prec)
        assert _nodes_equal(bvh.nodes, want.nodes), name
        assert np.array_equal(bvh.node_index, want.node_index)
        assert _flat_equal(bvh.flatten().nodes, O.flatten(want.nodes, prec))
        bvh.free()
    finally:
        ctx.set_option("build_small", -1); ctx.set_option("build_subtree", -1); ctx.set_option("build_gang", -1)


def test_forced_gangs_on_a_large_scene(api):
    """Gangs are normally off above ~150 k shapes; forced on, they coexist with queue-mode tile tasks (300 k shapes)."""
    from bvh_b200 import scenes as S
    shapes = S.create_n_cubes_aabbs(25000)
    want = O.build(shapes, "f32")
    ctx = api.Context.default()
    ctx.set_option("build_gang", 1)
    try:
        bvh = api.Bvh.build(shapes)
        assert _nodes_equal(bvh.nodes, want.nodes)
        assert np.array_equal(bvh.node_index, want.node_index)
        bvh.free()
    finally:
        ctx.set_option("build_gang", -1)


def test_concurrent_builds_on_two_contexts(api):
    """Two contexts building at the same time: the cooperative launch serialises the gang kernels instead of letting them
    starve each other of SMs."""
    import threading
    from bvh_b200 import scenes as S
    shapes = S.create_n_cubes_aabbs(8000)
    want = O.build(shapes, "f32")
    ctxs = [api.Context(0), api.Context(0)]
    res = [None, None]
    def work(i):
        ok = True
        for _ in range(10):
            b = api.Bvh.build(shapes, ctx=ctxs[i])
            ok = ok and np.array_equal(b.node_index, want.node_index)
            b.free()
        res[i] = ok
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th: t.start()
    for t in th: t.join(120)
    assert res == [True, True]


def test_capacity_error_and_fetch(api):
    """Caller buffer too small: BVHGPU_ERR_CAPACITY with the needed size, and bvhgpu_traverse_fetch_* returns the retained result."""
    import ctypes as C

    from bvh_b200 import capi
    from tests.scenes import sponza

    shapes = sponza()
    bvh = api.Bvh.build(shapes)
    o, d = np.tile([[-15.0, 2.0, 0.0]], (64, 1)), np.tile([[1.0, 0.05, 0.02]], (64, 1))
    rays = api.Ray.new(o, d)
    off = np.zeros(65, dtype=np.uint32); small = np.zeros(8, dtype=np.uint32); tot = C.c_size_t(0)
    L = capi.lib()
    st = L.bvhgpu_traverse_f32x3(bvh._h, 0, rays.ctypes.data_as(C.c_void_p), 64, off.ctypes.data_as(C.c_void_p), small.ctypes.data_as(C.c_void_p), 8, C.byref(tot))
    assert st == capi.ERR_CAPACITY and tot.value > 8 and off[-1] == tot.value
    full = np.zeros(tot.value, dtype=np.uint32)
    capi.check(L.bvhgpu_traverse_fetch_f32x3(bvh._h, full.ctypes.data_as(C.c_void_p), tot.value))
    off2, hits2 = bvh.traverse_batch(rays)
    assert np.array_equal(off2, off) and np.array_equal(hits2, full)


# ---- distance-ordered traversal ---------------------------------------------------------------------
@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("name", ["boxes21", "cubes300", "random3000"])
def test_ordered_traversal(api, name, prec):
    """Same hit set as Bvh::traverse, sorted by AABB entry distance (ascending) / exit distance (descending); the distances
    are Ray::intersection_slice_for_aabb of the leaf's AABB, bit for bit; ties keep the reference's DFS order."""
    shapes = scene(name, prec)
    want = O.build(shapes, prec)
    bvh = api.Bvh.build(shapes, prec=prec)
    rays = rays_for(shapes, 800, prec, seed=13, axis_aligned=100)
    ref = O.traverse(want.nodes, shapes, rays, O.MODE_RECURSIVE, prec)
    lists = O.per_ray_lists(ref.offsets, ref.hits)
    for ascending in (True, False):
        off, hits, dists = bvh.traverse_ordered(rays, ascending)
        assert np.array_equal(off.astype(np.uint64), ref.offsets)
        for i, (ray, lst) in enumerate(zip(rays, lists)):
            got_h, got_d = hits[off[i]:off[i + 1]], dists[off[i]:off[i + 1]]
            sl = [O.ray_slice(ray, shapes[h], prec) for h in lst]
            key = [s[0] if ascending else -s[1] for s in sl]
            order = sorted(range(len(lst)), key=lambda j: key[j])          # stable: ties keep DFS order
            assert got_h.tolist() == [int(lst[j]) for j in order], (name, i)
            assert got_d.tolist() == [sl[j][0] if ascending else sl[j][1] for j in order]
    # the reference's own monotonicity property (distance_traverse.rs:186-266)
    off, hits, dists = bvh.traverse_ordered(rays, True)
    for i in range(len(rays)):
        d = dists[off[i]:off[i + 1]]
        assert np.all(d[1:] >= d[:-1])
