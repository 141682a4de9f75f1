// tests/cpp/test_api.cpp -- the reference's fixed-scene tests (src/testbase.rs:66-267, src/bvh/bvh_impl.rs:557-690,
// src/flat_bvh.rs:602-625, src/bvh/iter.rs:256-308, src/bvh/optimization.rs:405-487) written against the C++ host
// mirror include/bvh_b200.hpp, i.e. through the C ABI on the GPU.  Exit code 0 = all passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

#include "bvh_b200.hpp"

using TAabb3 = bvh::Aabb<float>;
using TRay3 = bvh::Ray<float>;
using TBvh3 = bvh::Bvh<float>;

#define REQUIRE(cond)                                                              \
    do {                                                                           \
        if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); std::exit(1); } \
    } while (0)

// testbase.rs:66-103
struct UnitBox {
    int id;
    float pos[3];
    size_t node_index = 0;
    UnitBox(int i, float x, float y, float z) : id(i), pos{x, y, z} {}
    TAabb3 aabb() const {
        TAabb3 a;
        for (int k = 0; k < 3; ++k) { a.min[k] = pos[k] + -0.5f; a.max[k] = pos[k] + 0.5f; }
        return a;
    }
    void set_bh_node_index(size_t i) { node_index = i; }
    size_t bh_node_index() const { return node_index; }
    // testbase.rs:101-105: PointDistance = Aabb::min_distance_squared (aabb_impl.rs:618-629)
    float distance_squared(const float (&p)[3]) const {
        float d2 = 0.0f;
        for (int k = 0; k < 3; ++k) {
            const float hs = 0.5f, c = (pos[k] + -0.5f) + hs;
            float q = std::fabs(p[k] - c) - hs;
            q = q > 0.0f ? q : 0.0f;
            d2 += q * q;
        }
        return d2;
    }
};

// testbase.rs:109-116
static std::vector<UnitBox> generate_aligned_boxes() {
    std::vector<UnitBox> shapes;
    for (int x = -10; x < 11; ++x) shapes.emplace_back(x, (float)x, 0.0f, 0.0f);
    return shapes;
}

// testbase.rs:142-157
template <class BH>
static void traverse_and_verify(const TRay3& ray, const std::vector<UnitBox>& all, const BH& bh, const std::set<int>& expected) {
    auto hit = bh.traverse(ray, all);
    REQUIRE(hit.size() == expected.size());
    for (const UnitBox* s : hit) REQUIRE(expected.count(s->id) == 1);
}

// testbase.rs:174-225 (the ray cases)
template <class BH> static void traverse_some_built_bh(const std::vector<UnitBox>& all, const BH& bh) {
    {
        std::set<int> e;
        for (int id = -10; id < 11; ++id) e.insert(id);
        traverse_and_verify(TRay3({-1000.0f, 0.0f, 0.0f}, {1.0f, 0.0f, 0.0f}), all, bh, e);
    }
    traverse_and_verify(TRay3({0.0f, -1000.0f, 0.0f}, {0.0f, 1.0f, 0.0f}), all, bh, {0});
    traverse_and_verify(TRay3({6.0f, 0.5f, 0.0f}, {-2.0f, -1.0f, 0.0f}), all, bh, {4, 5, 6});
}

int main() {
    // test_build_bvh / test_traverse_bvh / test_traverse_flat_bvh / build_par twins
    for (int par = 0; par < 2; ++par) {
        auto boxes = generate_aligned_boxes();
        TBvh3 bvh = par ? TBvh3::build_par(boxes) : TBvh3::build(boxes);
        traverse_some_built_bh(boxes, bvh);
        auto flat = bvh.flatten();
        REQUIRE(flat.size() == 3 * boxes.size() - 2);
        {   // flatten_custom (flat_bvh.rs:214-251): a caller-defined node type built by a caller-defined constructor
            struct CustomStruct { TAabb3 aabb; uint32_t entry_index, exit_index, shape_index; int tag; };
            auto custom = bvh.flatten_custom([](const TAabb3& aabb, uint32_t entry, uint32_t exit, uint32_t shape) { return CustomStruct{aabb, entry, exit, shape, 42}; });
            REQUIRE(custom.size() == flat.size());
            for (size_t i = 0; i < custom.size(); ++i) {
                REQUIRE(custom[i].tag == 42 && custom[i].entry_index == flat.nodes[i].entry_index && custom[i].exit_index == flat.nodes[i].exit_index);
                REQUIRE(custom[i].shape_index == flat.nodes[i].shape_index && custom[i].aabb == flat.nodes[i].aabb);
            }
        }
        traverse_some_built_bh(boxes, flat);
        // iterator twin (iter.rs:269-308)
        auto it = bvh.traverse_iterator(TRay3({6.0f, 0.5f, 0.0f}, {-2.0f, -1.0f, 0.0f}), boxes);
        REQUIRE(it.size() == 3);
        // every shape in exactly one leaf, and the leaf knows it (bvh_impl.rs:588-614)
        auto nodes = bvh.nodes();
        REQUIRE(nodes.size() == 2 * boxes.size() - 1);
        std::set<size_t> seen;
        for (const auto& n : nodes) if (n.leaf) REQUIRE(seen.insert(n.shape_index).second);
        REQUIRE(seen.size() == boxes.size());
        for (size_t i = 0; i < boxes.size(); ++i) {
            REQUIRE(nodes[boxes[i].bh_node_index()].leaf);
            REQUIRE(nodes[boxes[i].bh_node_index()].shape_index == i);
        }
        // assert_tight (bvh_impl.rs:448-485): parent AABB == join of the children, exactly
        for (size_t i = 0; i < nodes.size(); ++i) {
            if (nodes[i].leaf) continue;
            for (size_t c : {nodes[i].child_l_index, nodes[i].child_r_index}) {
                REQUIRE(nodes[c].parent_index == i);
                if (!nodes[c].leaf) {
                    const TAabb3 mine = c == nodes[i].child_l_index ? nodes[i].child_l_aabb : nodes[i].child_r_aabb;
                    REQUIRE(nodes[c].child_l_aabb.join(nodes[c].child_r_aabb) == mine);
                }
            }
        }
    }
    // test_build_empty_bvh / test_flatten_empty_bvh (bvh_impl.rs:564-574, flat_bvh.rs:620-625)
    {
        std::vector<UnitBox> none;
        TBvh3 bvh = TBvh3::build(none);
        REQUIRE(bvh.nodes().empty());
        REQUIRE(bvh.flatten().empty());
        REQUIRE(bvh.traverse(TRay3({0.0f, 0.0f, 0.0f}, {1.0f, 0.0f, 0.0f}), none).empty());
    }
    // test_traverse_one_node_bvh_{no_,}intersection (bvh_impl.rs:665-690)
    {
        std::vector<UnitBox> miss{UnitBox(0, 0.0f, 1.0f, 2.0f)}, hit{UnitBox(0, 10.0f, 0.0f, 0.0f)};
        const TRay3 ray({0.0f, 0.0f, 0.0f}, {1.0f, 0.0f, 0.0f});
        TBvh3 a = TBvh3::build(miss), b = TBvh3::build(hit);
        REQUIRE(a.traverse(ray, miss).empty() && a.traverse_iterator(ray, miss).empty() && a.flatten().traverse(ray, miss).empty());
        REQUIRE(b.traverse(ray, hit).size() == 1 && b.traverse_iterator(ray, hit).size() == 1 && b.flatten().traverse(ray, hit).size() == 1);
    }
    // test_update_shapes_simple_update, the build half (optimization.rs:421-455): SAH pairs #0 and #1
    {
        std::vector<UnitBox> s{UnitBox(0, -50.0f, 0.0f, 0.0f), UnitBox(1, -40.0f, 0.0f, 0.0f), UnitBox(2, 50.0f, 0.0f, 0.0f)};
        TBvh3 bvh = TBvh3::build(s);
        auto nodes = bvh.nodes();
        REQUIRE(nodes[s[0].bh_node_index()].leaf && nodes[s[1].bh_node_index()].leaf);
        REQUIRE(nodes[s[0].bh_node_index()].parent_index == nodes[s[1].bh_node_index()].parent_index);
        // move #1 next to #2 and refit (the AABB half of update_shapes): tree stays tight
        s[1].pos[0] = 40.0f;
        bvh.refit(s);
        auto after = bvh.nodes();
        const auto& root = after[0];
        REQUIRE(root.child_l_aabb.join(root.child_r_aabb).max[0] == 50.5f);
    }
    // test_update_shapes_simple_update, the update half (optimization.rs:456-487): after moving #1 to x = 40, update_shapes
    // must make #1 and #2 siblings
    {
        std::vector<UnitBox> s{UnitBox(0, -50.0f, 0.0f, 0.0f), UnitBox(1, -40.0f, 0.0f, 0.0f), UnitBox(2, 50.0f, 0.0f, 0.0f)};
        TBvh3 bvh = TBvh3::build(s);
        s[1].pos[0] = 40.0f;
        REQUIRE(bvh.update_shapes(s) == 3);                       // the whole (tiny) tree is the degraded subtree
        auto nodes = bvh.nodes();
        for (size_t i = 0; i < s.size(); ++i) REQUIRE(nodes[s[i].bh_node_index()].leaf && nodes[s[i].bh_node_index()].shape_index == i);
        REQUIRE(nodes[s[1].bh_node_index()].parent_index == nodes[s[2].bh_node_index()].parent_index);
    }
    // test_consistent_after_update_shapes (optimization.rs:405-417): 21 boxes, six moved; the tree must contain every shape
    // inside every ancestor's stored AABB afterwards (assert_consistent)
    {
        auto s = generate_aligned_boxes();
        TBvh3 bvh = TBvh3::build(s);
        const float to[6][3] = {{10, 1, 2}, {-10, -10, 10}, {-10, 10, 10}, {-10, 10, -10}, {11, 1, 2}, {11, 2, 2}};
        for (int i = 0; i < 6; ++i) for (int k = 0; k < 3; ++k) s[i].pos[k] = to[i][k];
        bvh.update_shapes(s);
        auto nodes = bvh.nodes();
        for (size_t i = 0; i < s.size(); ++i) {
            size_t v = s[i].bh_node_index();
            REQUIRE(nodes[v].leaf && nodes[v].shape_index == i);
            const TAabb3 box = s[i].aabb();
            while (v != 0) {
                const size_t p = nodes[v].parent_index;
                const TAabb3 stored = nodes[p].child_l_index == v ? nodes[p].child_l_aabb : nodes[p].child_r_aabb;
                for (int k = 0; k < 3; ++k) REQUIRE(stored.min[k] <= box.min[k] && stored.max[k] >= box.max[k]);
                v = p;
            }
        }
        traverse_and_verify(TRay3({10.0f, -1000.0f, 2.0f}, {0.0f, 1.0f, 0.0f}), s, bvh, {-10});     // the box moved to (10, 1, 2)
    }
    // nearest_to doc example (bvh_impl.rs / flat_bvh.rs:493-507): 1000 unit boxes on the diagonal, query (5, 5.7, 5.3) -> id 5
    {
        std::vector<UnitBox> s;
        for (int i = 0; i < 1000; ++i) s.emplace_back(i, (float)i, (float)i, (float)i);
        TBvh3 bvh = TBvh3::build(s);
        const float q[3] = {5.0f, 5.7f, 5.3f};
        auto near = bvh.nearest_to(q, s);
        REQUIRE(near.first != nullptr && near.first->id == 5);
        REQUIRE(std::fabs(near.second - 0.2f) < 1e-6f);
        std::vector<UnitBox> none;
        TBvh3 e = TBvh3::build(none);
        REQUIRE(e.nearest_to(q, none).first == nullptr);
    }
    // ray / aabb known answers (ray_impl.rs:244-299)
    {
        TAabb3 zero_depth{{-1.0f, -1.0f, 1.0f}, {1.0f, 1.0f, 1.0f}};
        REQUIRE(TRay3({0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 1.0f}).intersects_aabb(zero_depth));
        TAabb3 unit{{-0.5f, -0.5f, -0.5f}, {0.5f, 0.5f, 0.5f}};
        REQUIRE(!TRay3({0.0f, 0.0f, -0.5f}, {1.0f, 0.0f, 0.0f}).intersects_aabb(unit));     // in-plane ray: NaN rule
        REQUIRE(!TRay3({0.0f, 0.5f, 0.0f}, {0.0f, 0.0f, 1.0f}).intersects_aabb(unit));
    }
    // a GPU error surfaces as an exception (the reference panics): NaN shape
    {
        std::vector<UnitBox> bad{UnitBox(0, 0.0f, 0.0f, 0.0f), UnitBox(1, NAN, 0.0f, 0.0f), UnitBox(2, 3.0f, 0.0f, 0.0f)};
        bool threw = false;
        try { TBvh3::build(bad); } catch (const bvh::Error& e) { threw = e.status == BVHGPU_ERR_NAN; }
        REQUIRE(threw);
    }
    std::printf("test_api: all reference fixed-scene tests passed\n");
    return 0;
}
