// Host-side symbolic analysis for the block-sparse multifrontal Cholesky used by the PVGO LM step.
//
// The reference solves the damped normal equations with a DENSE Cholesky inside PyPose
// (/root/reference/pvgo.py:169-178 -> pp.optim.LM + solver.Cholesky).  The graph structure
// (pvgo.py:36-51: VO / loop-closure edges + consecutive IMU pairs) is fixed over the LM iterations, so the
// elimination order, the fronts and every gather map are computed once here and uploaded.
//
// Ordering: 1-D nested dissection over the pose index.  Edges with |i-j| <= band_max are "short"; the
// endpoints of longer edges (loop closures) are promoted to the root.  An index interval is split by
// `band` consecutive indices (no short edge can jump over them) until it holds <= leaf_max poses.
// Every unknown block is 9x9: [tau(3), phi(3), v(3)] of one pose-velocity node.
#pragma once
#include <cstdint>
#include <vector>

namespace islam {

struct SymbolicOpts {
    int band_max = 16;   // edges with span above this are treated as loop closures
    int leaf_max = 8;    // poses per leaf front
    int pivot_max = 8;   // poses eliminated per front (9*pivot_max columns)
    int n_parts = 1;     // multi-GPU: number of contiguous pose windows (power of two)
    int dense_root_min = 33;   // loop-closure roots with at least this many poses become ONE dense front (dense_root.cuh)
};

struct Plan {
    int N = 0, E = 0, M = 0;
    int band = 1;
    // unique off-diagonal node pairs (lo < hi) of the Hessian pattern
    int P = 0;
    std::vector<int> pair_lo, pair_hi, pair_adj;      // pair_adj: 1 if hi == lo + 1 (IMU pair)
    std::vector<int> pair_eoff, pair_edges;            // CSR pair -> VO edges
    std::vector<int> node_eoff, node_edges;            // CSR node -> incident VO edges
    std::vector<int> edge_pair;                        // edge -> pair id
    // fronts in elimination order
    int F = 0;
    std::vector<int> f_np, f_nb, f_nodes_off, f_nodes; // node list: pivots then boundary (elimination order)
    std::vector<long long> f_Loff, f_Uoff;             // offsets (doubles) into the L / U arenas
    std::vector<int> f_parent, f_level, f_part;        // f_part: owning window (multi-GPU) or -1 = shared top
    std::vector<int> f_child_off, f_children;          // CSR front -> children
    std::vector<int> c_inv_off, c_inv;                 // per child entry: parent slot -> child boundary idx | -1
    std::vector<int> f_hmap_off, hmap;                 // per front (np+nb) x np: (pair<<1 | transpose) | -1
    std::vector<int> level_off, level_fronts;          // fronts grouped by level (ascending)
    std::vector<int> node_front, node_slot, node_pos;  // owner front / slot / elimination position
    long long L_doubles = 0, U_doubles = 0;
    int n_levels = 0, max_rows = 0, max_cols = 0, root_pivots = 0;
    int dense_root = -1;                                // front id of the dense root, or -1
    std::vector<int> root_slot;                         // [N] slot of a pose inside the dense root, or -1
    double factor_flops = 0;                            // multiply-adds of one numeric factorisation
};

// links: E x 2 (int64, as sample['link'] at /root/reference/train.py:254).  Returns 0 on success.
int build_plan(int N, int E, const int64_t* links, const SymbolicOpts& opts, Plan& plan);

}  // namespace islam
