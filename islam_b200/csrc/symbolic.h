// Host-side symbolic analysis, part 1: the block pattern of J^T W J that the assembly kernels fill.
//
// The reference solves the damped normal equations with a DENSE Cholesky inside PyPose
// (/root/reference/pvgo.py:169-178 -> pp.optim.LM + solver.Cholesky).  The graph structure
// (pvgo.py:36-51: VO / loop-closure edges + consecutive IMU pairs) is fixed over the LM iterations, so the unique
// node pairs, their CSR maps (here) and the elimination order / fronts (symbolic3.h) are computed once and uploaded.
#pragma once
#include <cstdint>
#include <vector>

namespace islam {

struct SymbolicOpts {
    int band_max = 16;   // edges with span above this are treated as loop closures
    int leaf_max = 8;    // leaf fronts hold up to 3*leaf_max variables
    int pivot_max = 8;   // a front eliminates up to 3*pivot_max variables (9*pivot_max columns)
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
};

// links: E x 2 (int64, as sample['link'] at /root/reference/train.py:254).  Returns 0 on success.
int build_plan(int N, int E, const int64_t* links, const SymbolicOpts& opts, Plan& plan);   // pairs + CSR only

}  // namespace islam
