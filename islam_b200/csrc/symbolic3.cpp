// Symbolic analysis over 3-dof variables: ordering with trimmed separators, fronts, elimination tree, push maps.
#include "symbolic3.h"

#include <algorithm>
#include <cstdlib>
#include <unordered_map>

namespace islam {

namespace {

enum : char { SIDE_L = 1, SIDE_S = 2, SIDE_R = 3 };

struct Builder3 {
    const SymbolicOpts& o;
    int band, leaf_vars, pivot_vars;
    const std::vector<std::vector<int>>& adj;
    std::vector<std::vector<int>> fronts;   // pivot lists in elimination order
    std::vector<int> parts;
    std::vector<int> stamp;                 // call id that last classified a variable
    std::vector<char> side;
    int calls = 0;

    void emit(std::vector<int>&& piv, int part) {
        for (size_t s = 0; s < piv.size(); s += pivot_vars) {      // over-wide pivot sets become chained fronts
            size_t e = std::min(piv.size(), s + (size_t)pivot_vars);
            fronts.emplace_back(piv.begin() + s, piv.begin() + e);
            parts.push_back(part);
        }
    }

    bool touches(int v, int id, char sd) const {
        for (int q : adj[v])
            if (stamp[q] == id && side[q] == sd) return true;
        return false;
    }

    // nested dissection of a sorted variable list; `nparts` pose windows (multi-GPU) starting at `part0` live inside it
    void recurse(std::vector<int>&& vars, int part0, int nparts) {
        if (vars.empty()) return;
        // disconnected pieces (the chain is cut wherever a whole pose was promoted to the root) are independent subtrees:
        // no separator between them
        if (nparts <= 1 && (int)vars.size() > leaf_vars) {
            const int id = ++calls;
            for (int v : vars) { stamp[v] = id; side[v] = 0; }
            std::vector<std::vector<int>> comps;
            std::vector<int> stack;
            for (int v0 : vars) {
                if (side[v0]) continue;
                comps.emplace_back();
                side[v0] = 1; stack.push_back(v0);
                while (!stack.empty()) {
                    int v = stack.back(); stack.pop_back();
                    comps.back().push_back(v);
                    for (int q : adj[v])
                        if (stamp[q] == id && !side[q]) { side[q] = 1; stack.push_back(q); }
                }
            }
            if (comps.size() > 1) {
                std::vector<int>().swap(vars);
                for (auto& c : comps) { std::sort(c.begin(), c.end()); recurse(std::move(c), part0, 1); }
                return;
            }
        }
        // the cut is placed among the poses that still have a tau / phi variable here: velocity-only tails (left over
        // from an earlier trim) hang off the chain and must not skew the balance of the tree
        int pmin = -1, pmax = -1;
        for (int v : vars)
            if (v % 3 != 2) { if (pmin < 0) pmin = v / 3; pmax = v / 3; }
        const int span = pmin < 0 ? 0 : pmax - pmin + 1;
        const bool can_split = span >= band + 2;
        if (nparts <= 1 && ((int)vars.size() <= leaf_vars || !can_split)) { emit(std::move(vars), part0); return; }
        if (!can_split) { emit(std::move(vars), -1); return; }      // more windows than the interval can host: shared
        int m = pmin + (span - band) / 2;
        if (m <= pmin) m = pmin + 1;
        if (m + band > pmax) m = pmax - band;
        const int id = ++calls;
        for (int v : vars) {
            int p = v / 3;
            stamp[v] = id;
            side[v] = p < m ? SIDE_L : (p < m + band ? SIDE_S : SIDE_R);
        }
        // trim the window: a variable with no neighbour right of the cut belongs to the left part, and vice versa
        for (int v : vars)
            if (side[v] == SIDE_S && !touches(v, id, SIDE_R)) side[v] = SIDE_L;
        for (auto it = vars.rbegin(); it != vars.rend(); ++it)
            if (side[*it] == SIDE_S && !touches(*it, id, SIDE_L)) side[*it] = SIDE_R;
        std::vector<int> L, S, R;
        for (int v : vars) (side[v] == SIDE_L ? L : side[v] == SIDE_S ? S : R).push_back(v);
        if (L.empty() || R.empty()) { emit(std::move(vars), nparts > 1 ? -1 : part0); return; }
        std::vector<int>().swap(vars);
        const int lparts = nparts > 1 ? nparts / 2 : 1;
        recurse(std::move(L), part0, nparts > 1 ? lparts : 1);
        recurse(std::move(R), nparts > 1 ? part0 + lparts : part0, nparts > 1 ? nparts - lparts : 1);
        if (!S.empty()) emit(std::move(S), nparts > 1 ? -1 : part0);
    }
};

}  // namespace

int build_plan3(const Plan& base, const int64_t* links, const SymbolicOpts& opts, Plan3& p) {
    const int N = base.N, V = 3 * N;
    if (N < 2) return -1;
    if (81LL * std::max(N, base.P) >= (1LL << 29)) return -8;      // orig_src packs the offset into 29 bits
    p = Plan3();
    p.N = N; p.V = V;

    // ---- variable graph ---------------------------------------------------------------------------------------
    std::vector<std::vector<int>> adj(V);
    auto link = [&](int a, int b) { adj[a].push_back(b); adj[b].push_back(a); };
    for (int i = 0; i < N; ++i) { link(3 * i, 3 * i + 1); link(3 * i, 3 * i + 2); }       // tau-phi (VO), tau-v (pvgo.py:51)
    std::unordered_map<long long, int> pair_id;
    pair_id.reserve(base.P * 2);
    for (int k = 0; k < base.P; ++k) {
        const int lo = base.pair_lo[k], hi = base.pair_hi[k];
        pair_id.emplace((long long)lo * N + hi, k);
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) link(3 * lo + a, 3 * hi + b);
        if (base.pair_adj[k]) { link(3 * lo + 2, 3 * hi + 2); link(3 * lo + 2, 3 * hi); }  // v-v (pvgo.py:42), v_i-tau_{i+1} (:51)
    }
    for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    std::vector<char> is_root(V, 0);
    for (int e = 0; e < base.E; ++e) {
        long long a = links[2 * e], b = links[2 * e + 1];
        if (std::llabs(a - b) > opts.band_max)
            for (int c = 0; c < 2; ++c) { is_root[3 * a + c] = 1; is_root[3 * b + c] = 1; }
    }

    // With a dense root (many closures) the chain below it must stay shallow and narrow: every `cut_every`-th closure
    // pose also gives its velocity to the root, which CUTS the chain there (nothing couples across a fully promoted
    // pose).  The pieces in between become independent subtrees whose fronts only see their own few closure poses,
    // instead of a 13-level tree whose top separators see thousands of root variables.
    {
        int n_root_poses = 0;
        for (int i = 0; i < N; ++i) n_root_poses += is_root[3 * i];
        if (n_root_poses >= opts.dense_root_min) {
            const int cut_every = 8;
            int k = 0;
            for (int i = 0; i < N; ++i)
                if (is_root[3 * i] && (k++ % cut_every) == 0) is_root[3 * i + 2] = 1;
        }
    }

    // ---- ordering ---------------------------------------------------------------------------------------------
    Builder3 bld{opts, base.band, 3 * opts.leaf_max, 3 * opts.pivot_max, adj, {}, {}, std::vector<int>(V, 0),
                 std::vector<char>(V, 0)};
    {
        std::vector<int> all;
        all.reserve(V);
        for (int v = 0; v < V; ++v)
            if (!is_root[v]) all.push_back(v);
        bld.recurse(std::move(all), 0, std::max(1, opts.n_parts));
        std::vector<int> root;
        for (int v = 0; v < V; ++v)
            if (is_root[v]) root.push_back(v);
        p.root_pivots = (int)root.size();
        if ((int)root.size() >= 2 * opts.dense_root_min) {          // one dense front, factored by the tiled dense path
            p.dense_root = (int)bld.fronts.size();
            bld.fronts.emplace_back(root);
            bld.parts.push_back(-1);
        } else if (!root.empty()) bld.emit(std::move(root), -1);
    }
    p.F = (int)bld.fronts.size();
    p.f_part = bld.parts;
    p.var_front.assign(V, -1); p.var_slot.assign(V, -1); p.var_pos.assign(V, -1);
    {
        int pos = 0;
        for (int f = 0; f < p.F; ++f)
            for (size_t s = 0; s < bld.fronts[f].size(); ++s) {
                int v = bld.fronts[f][s];
                if (p.var_front[v] >= 0) return -3;
                p.var_front[v] = f; p.var_slot[v] = (int)s; p.var_pos[v] = pos++;
            }
        if (pos != V) return -3;
    }

    // ---- symbolic factorisation: boundaries, parents, children ------------------------------------------------
    std::vector<std::vector<int>> boundary(p.F), children(p.F);
    p.f_parent.assign(p.F, -1);
    std::vector<int> mark(V, -1);
    for (int f = 0; f < p.F; ++f) {
        std::vector<int>& B = boundary[f];
        for (int v : bld.fronts[f])
            for (int q : adj[v])
                if (p.var_front[q] > f && mark[q] != f) { mark[q] = f; B.push_back(q); }
        for (int c : children[f])
            for (int q : boundary[c])
                if (p.var_front[q] != f && mark[q] != f) { mark[q] = f; B.push_back(q); }
        std::sort(B.begin(), B.end(), [&](int a, int b) { return p.var_pos[a] < p.var_pos[b]; });
        if (!B.empty()) {
            int par = p.var_front[B[0]];
            p.f_parent[f] = par;
            children[par].push_back(f);
        }
    }
    // a shared (multi-GPU) front's ancestors must be shared too; so must any front whose children belong to two windows
    for (int f = 0; f < p.F; ++f)
        if (p.f_part[f] < 0)
            for (int a = p.f_parent[f]; a >= 0 && p.f_part[a] >= 0; a = p.f_parent[a]) p.f_part[a] = -1;
    for (int f = 0; f < p.F; ++f)
        if (p.f_parent[f] >= 0 && p.f_part[p.f_parent[f]] >= 0 && p.f_part[f] != p.f_part[p.f_parent[f]])
            for (int a = p.f_parent[f]; a >= 0; a = p.f_parent[a]) p.f_part[a] = -1;

    // ---- flatten ----------------------------------------------------------------------------------------------
    p.f_np.resize(p.F); p.f_npad.resize(p.F); p.f_nb.resize(p.F); p.f_vars_off.assign(p.F + 1, 0);
    p.f_Loff.resize(p.F); p.f_Uoff.resize(p.F); p.f_Ioff.resize(p.F); p.f_level.assign(p.F, 0);
    p.f_child_off.assign(p.F + 1, 0);
    for (int f = 0; f < p.F; ++f) {
        const int np = (int)bld.fronts[f].size(), nb = (int)boundary[f].size();
        const int npad = (f == p.dense_root) ? np : 3 * ((np + 2) / 3);
        p.f_np[f] = np; p.f_npad[f] = npad; p.f_nb[f] = nb;
        p.f_vars_off[f + 1] = p.f_vars_off[f] + npad + nb;
        p.f_vars.insert(p.f_vars.end(), bld.fronts[f].begin(), bld.fronts[f].end());
        p.f_vars.insert(p.f_vars.end(), npad - np, -1);
        p.f_vars.insert(p.f_vars.end(), boundary[f].begin(), boundary[f].end());
        const long long cols = 3LL * npad, ub = 3LL * nb + 1, rows = cols + ub;
        // even: every panel starts 16-byte aligned (TMA bulk copies); the dense root's leading dimension is even too
        // (dense_root.cuh: its operand panels are fetched by bulk copies of whole column segments)
        p.f_Loff[f] = p.L_doubles; p.L_doubles += (((f == p.dense_root) ? ((rows + 1) & ~1LL) : rows) * cols + 1) & ~1LL;
        const long long ube = (ub + 1) & ~1LL;                     // paired-column layout of solver3.cuh (f3_ulen)
        p.f_Uoff[f] = p.U_doubles; p.U_doubles += ube * (ube / 2 + 1);
        p.f_Ioff[f] = p.I_doubles; p.I_doubles += (f == p.dense_root) ? 0 : 81LL * (npad / 3);
        p.max_rows = std::max<int>(p.max_rows, (int)rows);
        p.max_cols = std::max<int>(p.max_cols, (int)cols);
        p.max_ub = std::max<int>(p.max_ub, (int)ub);
        p.factor_flops += 0.5 * (double)rows * cols * cols + 0.5 * (double)ub * ub * cols;
        for (int c : children[f]) p.f_level[f] = std::max(p.f_level[f], p.f_level[c] + 1);
        p.n_levels = std::max(p.n_levels, p.f_level[f] + 1);
        p.f_child_off[f + 1] = p.f_child_off[f] + (int)children[f].size();
        p.f_children.insert(p.f_children.end(), children[f].begin(), children[f].end());
    }
    // push maps (child boundary -> parent slot) and original-entry lists
    p.c_map_off.assign(p.f_children.size() + 1, 0);
    p.f_orig_off.assign(p.F + 1, 0);
    std::vector<int> slot_in(V, -1);
    for (int f = 0; f < p.F; ++f) {
        const int npad = p.f_npad[f], nb = p.f_nb[f], ns = npad + nb;
        const int* vars = &p.f_vars[p.f_vars_off[f]];
        for (int s = 0; s < ns; ++s)
            if (vars[s] >= 0) slot_in[vars[s]] = s;
        for (int k = p.f_child_off[f]; k < p.f_child_off[f + 1]; ++k) {
            const int c = p.f_children[k];
            const int* cb = &p.f_vars[p.f_vars_off[c] + p.f_npad[c]];
            p.c_map_off[k + 1] = p.c_map_off[k] + p.f_nb[c];
            int prev = -1;
            for (int b = 0; b < p.f_nb[c]; ++b) {
                int s = slot_in[cb[b]];
                if (s < 0 || s <= prev) return -4;      // multifrontal containment / monotonicity violated
                prev = s;
                p.c_map.push_back(s);
            }
        }
        for (int cs = 0; cs < p.f_np[f]; ++cs) {
            const int vc = vars[cs], pc = vc / 3, cc = vc % 3;
            p.orig_rs.push_back(cs); p.orig_cs.push_back(cs);
            p.orig_src.push_back((81 * pc + 27 * cc + 3 * cc) << 2);
            for (int q : adj[vc]) {
                const int rs = slot_in[q];
                if (rs < 0 || rs <= cs || p.var_front[q] < f) continue;     // earlier-eliminated or upper triangle
                const int pr = q / 3, cr = q % 3;
                int src;
                if (pr == pc) src = (81 * pc + 27 * cr + 3 * cc) << 2;                       // Hd[pose][3cr+r][3cc+c]
                else {
                    auto it = pair_id.find((long long)std::min(pr, pc) * N + std::max(pr, pc));
                    if (it == pair_id.end()) return -4;
                    const int pid = it->second;
                    if (pr < pc) src = ((81 * pid + 27 * cr + 3 * cc) << 2) | 2;             // row is lo: Ho[p][3cr+r][3cc+c]
                    else src = ((81 * pid + 27 * cc + 3 * cr) << 2) | 2 | 1;                  // row is hi: Ho[p][3cc+c][3cr+r]
                }
                p.orig_rs.push_back(rs); p.orig_cs.push_back(cs); p.orig_src.push_back(src);
            }
        }
        p.f_orig_off[f + 1] = (int)p.orig_rs.size();
        for (int s = 0; s < ns; ++s)
            if (vars[s] >= 0) slot_in[vars[s]] = -1;
    }
    p.root_slot.assign(V, -1);
    if (p.dense_root >= 0)
        for (int k = 0; k < p.f_np[p.dense_root]; ++k) p.root_slot[p.f_vars[p.f_vars_off[p.dense_root] + k]] = k;
    // level schedule
    p.level_off.assign(p.n_levels + 1, 0);
    for (int f = 0; f < p.F; ++f) p.level_off[p.f_level[f] + 1]++;
    for (int l = 0; l < p.n_levels; ++l) p.level_off[l + 1] += p.level_off[l];
    p.level_fronts.resize(p.F);
    {
        std::vector<int> cur(p.level_off.begin(), p.level_off.end() - 1);
        for (int f = 0; f < p.F; ++f) p.level_fronts[cur[p.f_level[f]]++] = f;
    }
    p.scalars = {p.F, p.n_levels, p.dense_root, p.max_rows, p.max_cols, p.max_ub, p.root_pivots};
    return 0;
}

}  // namespace islam
