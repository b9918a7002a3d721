// Symbolic analysis: ordering, fronts, elimination tree and gather maps (see symbolic.h).
#include "symbolic.h"

#include <algorithm>
#include <cstdlib>
#include <map>
#include <unordered_map>

namespace islam {

namespace {

struct Builder {
    const SymbolicOpts& o;
    int N, band;
    const std::vector<char>& is_root;
    std::vector<std::vector<int>> fronts;   // pivot lists in elimination order
    std::vector<int> parts;

    void emit(std::vector<int>&& piv, int part) {
        // split over-wide fronts into chained chunks of <= pivot_max poses
        for (size_t s = 0; s < piv.size(); s += o.pivot_max) {
            size_t e = std::min(piv.size(), s + (size_t)o.pivot_max);
            fronts.emplace_back(piv.begin() + s, piv.begin() + e);
            parts.push_back(part);
        }
    }
    std::vector<int> collect(int lo, int hi) const {
        std::vector<int> v;
        for (int n = lo; n < hi; ++n)
            if (!is_root[n]) v.push_back(n);
        return v;
    }
    // nested dissection of the index interval [lo, hi); `nparts` windows starting at `part0` live inside it
    void recurse(int lo, int hi, int part0, int nparts) {
        if (hi <= lo) return;
        std::vector<int> all = collect(lo, hi);
        if (all.empty()) return;
        bool can_split = (hi - lo) >= band + 2;
        if (nparts <= 1 && ((int)all.size() <= o.leaf_max || !can_split)) {
            emit(std::move(all), part0);
            return;
        }
        if (!can_split) {   // more windows requested than the interval can host: shared front
            emit(std::move(all), -1);
            return;
        }
        int m = lo + (hi - lo - band) / 2;
        if (m <= lo) m = lo + 1;
        if (m + band >= hi) m = hi - band - 1;
        int lparts = nparts > 1 ? nparts / 2 : 1;
        recurse(lo, m, part0, nparts > 1 ? lparts : 1);
        recurse(m + band, hi, nparts > 1 ? part0 + lparts : part0, nparts > 1 ? nparts - lparts : 1);
        std::vector<int> sep = collect(m, m + band);
        if (!sep.empty()) emit(std::move(sep), nparts > 1 ? -1 : part0);
    }
};

}  // namespace

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
    std::vector<char> is_root(N, 0);
    for (int e = 0; e < E; ++e) {
        int64_t a = links[2 * e], b = links[2 * e + 1];
        if (a < 0 || b < 0 || a >= N || b >= N || a == b) return -2;
        p.edge_pair[e] = get_pair((int)a, (int)b);
        int span = (int)std::llabs(a - b);
        if (span <= opts.band_max) band = std::max(band, span);
        else { is_root[a] = 1; is_root[b] = 1; }
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
    // adjacency
    std::vector<std::vector<int>> adj(N);
    for (int i = 0; i < p.P; ++i) {
        adj[p.pair_lo[i]].push_back(p.pair_hi[i]);
        adj[p.pair_hi[i]].push_back(p.pair_lo[i]);
    }

    // ---- ordering -----------------------------------------------------------------------------------
    Builder bld{opts, N, band, is_root, {}, {}};
    bld.recurse(0, N, 0, std::max(1, opts.n_parts));
    {
        std::vector<int> root;
        for (int n = 0; n < N; ++n)
            if (is_root[n]) root.push_back(n);
        p.root_pivots = (int)root.size();
        if ((int)root.size() >= opts.dense_root_min) {          // one dense front, factored by the tiled dense path
            p.dense_root = (int)bld.fronts.size();
            bld.fronts.emplace_back(root);
            bld.parts.push_back(-1);
        } else if (!root.empty()) bld.emit(std::move(root), -1);
    }
    p.F = (int)bld.fronts.size();
    p.f_part = bld.parts;
    p.node_front.assign(N, -1); p.node_slot.assign(N, -1); p.node_pos.assign(N, -1);
    {
        int pos = 0;
        for (int f = 0; f < p.F; ++f)
            for (size_t s = 0; s < bld.fronts[f].size(); ++s) {
                int n = bld.fronts[f][s];
                p.node_front[n] = f; p.node_slot[n] = (int)s; p.node_pos[n] = pos++;
            }
        if (pos != N) return -3;
    }

    // ---- symbolic factorisation: boundaries, parents, children --------------------------------------
    std::vector<std::vector<int>> boundary(p.F), children(p.F);
    p.f_parent.assign(p.F, -1);
    std::vector<int> stamp(N, -1);
    for (int f = 0; f < p.F; ++f) {
        std::vector<int>& B = boundary[f];
        for (int n : bld.fronts[f])
            for (int q : adj[n])
                if (p.node_front[q] > f && stamp[q] != f) { stamp[q] = f; B.push_back(q); }
        for (int c : children[f])
            for (int q : boundary[c])
                if (p.node_front[q] != f && stamp[q] != f) { stamp[q] = f; B.push_back(q); }
        std::sort(B.begin(), B.end(), [&](int a, int b) { return p.node_pos[a] < p.node_pos[b]; });
        if (!B.empty()) {
            int par = p.node_front[B[0]];
            p.f_parent[f] = par;
            children[par].push_back(f);
        }
    }
    // a shared (multi-GPU) front's ancestors must be shared too
    for (int f = 0; f < p.F; ++f)
        if (p.f_part[f] < 0)
            for (int a = p.f_parent[f]; a >= 0 && p.f_part[a] >= 0; a = p.f_parent[a]) p.f_part[a] = -1;
    for (int f = 0; f < p.F; ++f)
        if (p.f_parent[f] >= 0 && p.f_part[p.f_parent[f]] >= 0 && p.f_part[f] != p.f_part[p.f_parent[f]])
            for (int a = p.f_parent[f]; a >= 0; a = p.f_parent[a]) p.f_part[a] = -1;

    // ---- flatten -----------------------------------------------------------------------------------
    p.f_np.resize(p.F); p.f_nb.resize(p.F); p.f_nodes_off.assign(p.F + 1, 0);
    p.f_Loff.resize(p.F); p.f_Uoff.resize(p.F); p.f_level.assign(p.F, 0);
    p.f_child_off.assign(p.F + 1, 0); p.f_hmap_off.assign(p.F + 1, 0);
    for (int f = 0; f < p.F; ++f) {
        int np = (int)bld.fronts[f].size(), nb = (int)boundary[f].size();
        p.f_np[f] = np; p.f_nb[f] = nb;
        p.f_nodes_off[f + 1] = p.f_nodes_off[f] + np + nb;
        p.f_nodes.insert(p.f_nodes.end(), bld.fronts[f].begin(), bld.fronts[f].end());
        p.f_nodes.insert(p.f_nodes.end(), boundary[f].begin(), boundary[f].end());
        long long rows = 9LL * (np + nb) + 1, cols = 9LL * np, ub = 9LL * nb + 1;
        p.f_Loff[f] = p.L_doubles; p.L_doubles += rows * cols;
        p.f_Uoff[f] = p.U_doubles; p.U_doubles += ub * ub;
        p.max_rows = std::max<int>(p.max_rows, (int)rows);
        p.max_cols = std::max<int>(p.max_cols, (int)cols);
        p.factor_flops += 0.5 * (double)rows * cols * cols + 0.5 * (double)ub * ub * cols;
        for (int c : children[f]) p.f_level[f] = std::max(p.f_level[f], p.f_level[c] + 1);
        p.n_levels = std::max(p.n_levels, p.f_level[f] + 1);
        p.f_child_off[f + 1] = p.f_child_off[f] + (int)children[f].size();
        p.f_children.insert(p.f_children.end(), children[f].begin(), children[f].end());
        p.f_hmap_off[f + 1] = p.f_hmap_off[f] + (f == p.dense_root ? 0 : (np + nb) * np);
    }
    // inverse child maps and H gather maps
    p.c_inv_off.assign(p.f_children.size() + 1, 0);
    p.hmap.assign(p.f_hmap_off[p.F], -1);
    std::vector<int> slot_in(N, -1);
    for (int f = 0; f < p.F; ++f) {
        int np = p.f_np[f], nb = p.f_nb[f], ns = np + nb;
        const int* nodes = &p.f_nodes[p.f_nodes_off[f]];
        for (int s = 0; s < ns; ++s) slot_in[nodes[s]] = s;
        for (int k = p.f_child_off[f]; k < p.f_child_off[f + 1]; ++k) {
            int c = p.f_children[k];
            if (f == p.dense_root) { p.c_inv_off[k + 1] = p.c_inv_off[k]; continue; }   // children are pushed, not pulled
            p.c_inv_off[k + 1] = p.c_inv_off[k] + ns;
            size_t base = p.c_inv.size();
            p.c_inv.resize(base + ns, -1);
            const int* cb = &p.f_nodes[p.f_nodes_off[c] + p.f_np[c]];
            for (int b = 0; b < p.f_nb[c]; ++b) {
                int s = slot_in[cb[b]];
                if (s < 0) return -4;      // multifrontal containment violated
                p.c_inv[base + s] = b;
            }
        }
        int* hm = p.hmap.data() + p.f_hmap_off[f];
        for (int cs = 0; cs < np && f != p.dense_root; ++cs) {
            int nc = nodes[cs];
            for (int q : adj[nc]) {
                int rs = slot_in[q];
                if (rs < 0 || rs <= cs) continue;     // earlier-eliminated neighbour or upper triangle
                if (p.node_front[q] < f) continue;
                auto it = pair_id.find(std::make_pair(std::min(nc, q), std::max(nc, q)));
                int pid = it->second;
                int transpose = (q > nc) ? 1 : 0;     // row node == hi  => element (a,b) is Ho[pid][b][a]
                hm[rs * np + cs] = (pid << 1) | transpose;
            }
        }
        for (int s = 0; s < ns; ++s) slot_in[nodes[s]] = -1;
    }
    p.root_slot.assign(N, -1);
    if (p.dense_root >= 0)
        for (int k = 0; k < p.f_np[p.dense_root]; ++k) p.root_slot[p.f_nodes[p.f_nodes_off[p.dense_root] + k]] = k;
    // level schedule
    p.level_off.assign(p.n_levels + 1, 0);
    for (int f = 0; f < p.F; ++f) p.level_off[p.f_level[f] + 1]++;
    for (int l = 0; l < p.n_levels; ++l) p.level_off[l + 1] += p.level_off[l];
    p.level_fronts.resize(p.F);
    {
        std::vector<int> cur(p.level_off.begin(), p.level_off.end() - 1);
        for (int f = 0; f < p.F; ++f) p.level_fronts[cur[p.f_level[f]]++] = f;
    }
    return 0;
}

}  // namespace islam
