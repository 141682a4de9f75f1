// tools/sim_coherence.cpp -- DEV TOOL (not product, not shipped): CPU model of the traversal kernel's L1 tag-stage load.
//   g++ -O3 -march=x86-64-v3 -ffp-contract=off -std=c++17 -pthread -I oracle tools/sim_coherence.cpp -o /tmp/sim && /tmp/sim
// Model: a warp = 32 consecutive rays of an ordering, walking the preorder skip-link records in lock step; per warp step the
// tag stage pays one lookup per DISTINCT record among the active lanes.  Reports, per ray ordering, lane visits, lookups and
// warp steps, and the share of visits that fall into the top of the tree (nodes with >= C shapes below them).
#include "bvh_oracle.hpp"
#include <algorithm>
#include <cstdio>
#include <numeric>
using namespace orc;

struct Rec { float mn[3]; uint32_t skip; float mx[3]; uint32_t shape; uint32_t count; };

static uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) {
    auto spread = [](uint64_t v) { v &= 0x1FFFFF; v = (v | v << 32) & 0x1F00000000FFFFull; v = (v | v << 16) & 0x1F0000FF0000FFull;
                                   v = (v | v << 8) & 0x100F00F00F00F00Full; v = (v | v << 4) & 0x10C30C30C30C30C3ull; v = (v | v << 2) & 0x1249249249249249ull; return v; };
    return spread(x) | (spread(y) << 1) | (spread(z) << 2);
}

int main(int argc, char** argv) {
    const uint32_t n_cubes = argc > 1 ? atoi(argv[1]) : 10000;
    const uint32_t R = argc > 2 ? atoi(argv[2]) : 200000;
    std::vector<float> tris;
    const Aabb3<float> bounds = default_bounds<float>();
    create_n_cubes(n_cubes, bounds, tris);
    const uint32_t n = n_cubes * 12;
    std::vector<Aabb3<float>> shapes(n);
    for (uint32_t i = 0; i < n; ++i) shapes[i] = triangle_aabb(&tris[9 * i], &tris[9 * i + 3], &tris[9 * i + 6]);
    std::vector<Node<float>> nodes(2 * n - 1);
    std::vector<uint32_t> nidx(n);
    build(shapes.data(), n, nodes.data(), nidx.data());
    // records: record r = node r+1
    std::vector<uint32_t> cnt(2 * n - 1);
    for (uint32_t i = 2 * n - 1; i-- > 0;) cnt[i] = nodes[i].is_leaf() ? 1 : cnt[nodes[i].child_l] + cnt[nodes[i].child_r];
    std::vector<Rec> rec(2 * n - 2);
    for (uint32_t i = 1; i < 2 * n - 1; ++i) {
        const Node<float>& p = nodes[nodes[i].parent];
        const Aabb3<float>& b = p.child_l == i ? p.l_aabb : p.r_aabb;
        Rec r;
        for (int k = 0; k < 3; ++k) { r.mn[k] = b.min[k]; r.mx[k] = b.max[k]; }
        r.skip = (i - 1) + (2 * cnt[i] - 1);
        r.shape = nodes[i].is_leaf() ? nodes[i].shape : U32_MAX;
        r.count = cnt[i];
        rec[i - 1] = r;
    }
    std::vector<Ray3<float>> rays(R);
    uint64_t seed = 0;
    for (uint32_t i = 0; i < R; ++i) rays[i] = create_ray(seed, bounds);

    auto hit = [&](const Ray3<float>& ray, const Rec& r) { Aabb3<float> b; for (int k = 0; k < 3; ++k) { b.min[k] = r.mn[k]; b.max[k] = r.mx[k]; } return ray_intersects_aabb(ray, b); };

    // top-of-tree share
    {
        const uint32_t Cs[] = {1, 16, 32, 64, 128, 256, 512, 1024, 4096};
        uint64_t tot = 0, top[9] = {};
        uint32_t ntop[9] = {};
        for (auto& r : rec) for (int c = 0; c < 9; ++c) if (r.count >= Cs[c]) ntop[c]++;
        for (uint32_t q = 0; q < std::min<uint32_t>(R, 50000); ++q) {
            uint32_t i = 0;
            while (i < rec.size()) { ++tot; for (int c = 0; c < 9; ++c) if (rec[i].count >= Cs[c]) top[c]++; i = hit(rays[q], rec[i]) ? i + 1 : rec[i].skip; }
        }
        printf("visits/ray %.1f\n", (double)tot / std::min<uint32_t>(R, 50000));
        for (int c = 0; c < 9; ++c) printf("  count >= %5u : %7u records (+fringe), %5.1f %% of visits\n", Cs[c], ntop[c], 100.0 * top[c] / tot);
    }

    auto simulate = [&](const char* name, const std::vector<uint32_t>& order) {
        uint64_t lane_visits = 0, lookups = 0, steps = 0;
        for (uint32_t w = 0; w + 32 <= R; w += 32) {
            uint32_t pos[32];
            for (int l = 0; l < 32; ++l) pos[l] = 0;
            for (;;) {
                uint32_t act[32]; int na = 0;
                for (int l = 0; l < 32; ++l) if (pos[l] < rec.size()) act[na++] = pos[l];
                if (!na) break;
                ++steps; lane_visits += na;
                std::sort(act, act + na);
                lookups += std::unique(act, act + na) - act;
                for (int l = 0; l < 32; ++l) if (pos[l] < rec.size()) { const Rec& r = rec[pos[l]]; pos[l] = hit(rays[order[w + l]], r) ? pos[l] + 1 : r.skip; }
            }
        }
        printf("%-28s lane-visits/ray %6.1f  lookups/ray %6.1f  warp-steps/ray %5.2f  lanes/step %5.1f  distinct/step %5.1f\n", name,
               (double)lane_visits / R, (double)lookups / R, (double)steps / R, (double)lane_visits / steps, (double)lookups / steps);
    };
    // the persistent refill kernel: a pool of W warps pulls rays from a ticket counter; a warp refills when >= 8 lanes are idle
    auto simulate_refill = [&](const char* name, const std::vector<uint32_t>& order) {
        const int W = 2048;                                   // concurrently resident warps (model: round-robin stepping)
        struct Warp { uint32_t pos[32], ray[32]; };
        std::vector<Warp> warps(W);
        for (auto& w : warps) for (int l = 0; l < 32; ++l) w.ray[l] = U32_MAX;
        uint32_t ticket = 0;
        uint64_t lane_visits = 0, lookups = 0, steps = 0;
        bool any = true;
        while (any) {
            any = false;
            for (auto& w : warps) {
                int idle = 0;
                for (int l = 0; l < 32; ++l) idle += w.ray[l] == U32_MAX;
                if ((idle >= 8 || idle == 32) && ticket < R) for (int l = 0; l < 32 && ticket < R; ++l) if (w.ray[l] == U32_MAX) { w.ray[l] = order[ticket++]; w.pos[l] = 0; }
                uint32_t act[32]; int na = 0;
                for (int l = 0; l < 32; ++l) if (w.ray[l] != U32_MAX) act[na++] = w.pos[l];
                if (!na) continue;
                any = true;
                ++steps; lane_visits += na;
                std::sort(act, act + na);
                lookups += std::unique(act, act + na) - act;
                for (int l = 0; l < 32; ++l) if (w.ray[l] != U32_MAX) {
                    const Rec& r = rec[w.pos[l]];
                    w.pos[l] = hit(rays[w.ray[l]], r) ? w.pos[l] + 1 : r.skip;
                    if (w.pos[l] >= rec.size()) w.ray[l] = U32_MAX;
                }
            }
        }
        printf("refill %-21s lane-visits/ray %6.1f  lookups/ray %6.1f  warp-steps/ray %5.2f  lanes/step %5.1f  distinct/step %5.1f\n", name,
               (double)lane_visits / R, (double)lookups / R, (double)steps / R, (double)lane_visits / steps, (double)lookups / steps);
    };
    std::vector<uint32_t> id(R);
    std::iota(id.begin(), id.end(), 0);
    simulate("original order", id);
    simulate_refill("original order", id);
    auto key_origin = [&](uint32_t i, int bits) {
        uint32_t q[3];
        for (int k = 0; k < 3; ++k) { double f = (rays[i].origin[k] + 100000.0) / 200000.0; f = std::min(std::max(f, 0.0), 0.999999); q[k] = (uint32_t)(f * (1u << bits)); }
        return morton3(q[0], q[1], q[2]);
    };
    for (int bits : {4, 7, 10}) {
        std::vector<uint32_t> o = id;
        std::vector<uint64_t> key(R);
        for (uint32_t i = 0; i < R; ++i) key[i] = key_origin(i, bits);
        std::stable_sort(o.begin(), o.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
        char nm[64]; snprintf(nm, sizeof nm, "morton(origin) %d bits/axis", bits);
        simulate(nm, o);
    }
    {   // direction octant major, then origin
        std::vector<uint32_t> o = id;
        std::vector<uint64_t> key(R);
        for (uint32_t i = 0; i < R; ++i) { uint64_t oct = (rays[i].direction[0] < 0) | ((rays[i].direction[1] < 0) << 1) | ((rays[i].direction[2] < 0) << 2); key[i] = (oct << 60) | key_origin(i, 10); }
        std::stable_sort(o.begin(), o.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
        simulate("octant, morton(origin)", o);
        simulate_refill("octant, morton(origin)", o);
    }
    {   // origin coarse (4 bits/axis), then direction quantised (morton of dir 5 bits)
        std::vector<uint32_t> o = id;
        std::vector<uint64_t> key(R);
        for (uint32_t i = 0; i < R; ++i) {
            uint32_t q[3];
            for (int k = 0; k < 3; ++k) q[k] = (uint32_t)(std::min(std::max((rays[i].direction[k] + 1.0) * 0.5, 0.0), 0.999999) * 32);
            key[i] = (key_origin(i, 3) << 15) | morton3(q[0], q[1], q[2]);
        }
        std::stable_sort(o.begin(), o.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
        simulate("origin 3b, morton(dir) 5b", o);
        simulate_refill("origin 3b, morton(dir) 5b", o);
    }
    return 0;
}
