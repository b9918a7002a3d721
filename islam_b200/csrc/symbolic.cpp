// Symbolic analysis, part 1: unique node pairs of the Hessian pattern and their CSR maps (see symbolic.h).
#include "symbolic.h"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <unordered_map>

namespace islam {

int build_plan(int N, int E, const int64_t* links, const SymbolicOpts& opts, Plan& p) {
    if (N < 2 || E < 0) return -1;
    p = Plan();
    p.N = N; p.E = E; p.M = N - 1;

    // ---- unique pairs: VO / loop-closure edges + IMU chain pairs -------------------------------------
    std::map<std::pair<int, int>, int> pair_id;
    auto get_pair = [&](int a, int b) {
        if (a > b) std::swap(a, b);
        auto key = std::make_pair(a, b);
        auto it = pair_id.find(key);
        if (it != pair_id.end()) return it->second;
        int id = (int)pair_id.size();
        pair_id.emplace(key, id);
        return id;
    };
    for (int i = 0; i + 1 < N; ++i) get_pair(i, i + 1);          // pair id == i for the IMU chain
    p.edge_pair.resize(E);
    int band = 1;
    for (int e = 0; e < E; ++e) {
        int64_t a = links[2 * e], b = links[2 * e + 1];
        if (a < 0 || b < 0 || a >= N || b >= N || a == b) return -2;
        p.edge_pair[e] = get_pair((int)a, (int)b);
        int span = (int)std::llabs(a - b);
        if (span <= opts.band_max) band = std::max(band, span);
    }
    p.band = band;
    p.P = (int)pair_id.size();
    p.pair_lo.resize(p.P); p.pair_hi.resize(p.P); p.pair_adj.resize(p.P);
    for (auto& kv : pair_id) {
        p.pair_lo[kv.second] = kv.first.first;
        p.pair_hi[kv.second] = kv.first.second;
        p.pair_adj[kv.second] = (kv.first.second == kv.first.first + 1);
    }
    // CSR pair -> edges, node -> edges (edge order preserved => deterministic summation order)
    p.pair_eoff.assign(p.P + 1, 0);
    p.node_eoff.assign(N + 1, 0);
    for (int e = 0; e < E; ++e) {
        p.pair_eoff[p.edge_pair[e] + 1]++;
        p.node_eoff[links[2 * e] + 1]++;
        p.node_eoff[links[2 * e + 1] + 1]++;
    }
    for (int i = 0; i < p.P; ++i) p.pair_eoff[i + 1] += p.pair_eoff[i];
    for (int i = 0; i < N; ++i) p.node_eoff[i + 1] += p.node_eoff[i];
    p.pair_edges.resize(E);
    p.node_edges.resize(2 * (size_t)E);
    {
        std::vector<int> pc(p.pair_eoff.begin(), p.pair_eoff.end() - 1), nc(p.node_eoff.begin(), p.node_eoff.end() - 1);
        for (int e = 0; e < E; ++e) {
            p.pair_edges[pc[p.edge_pair[e]]++] = e;
            p.node_edges[nc[links[2 * e]]++] = e;
            p.node_edges[nc[links[2 * e + 1]]++] = e;
        }
    }
    return 0;
}

}  // namespace islam
