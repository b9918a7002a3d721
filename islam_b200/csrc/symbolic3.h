// Symbolic analysis, generation 2: the unknowns are 3-dof VARIABLES (tau_i, phi_i, v_i of every pose-velocity node)
// instead of 9-dof pose slots.
//
// Why: in the Hessian of /root/reference/pvgo.py:26-64 the velocity v_i only couples to v_{i-1}, v_{i+1}, tau_i and
// tau_{i+1} (pvgo.py:42 and :51), while the VO band (pvgo.py:36-39) couples the 6-dof poses up to `band` nodes apart.
// A vertex separator of the band-b chain therefore needs the tau/phi variables of b consecutive poses but only ONE
// velocity: 6b+3 unknowns instead of 9b (51 instead of 72 at b = 8).  Front cost is quadratic-to-cubic in that
// width, so every level of the elimination tree gets ~2.5x cheaper.  The trimming is generic (a window variable with
// no neighbour on one side of the cut moves to the other side), not specific to this pattern.
//
// Same linear system as PyPose's dense Cholesky; only the elimination order differs.
#pragma once
#include <cstdint>
#include <vector>

#include "symbolic.h"

namespace islam {

struct Plan3 {
    int N = 0, V = 0;                                  // poses, variables (3N: var u = 3*pose + {0 tau, 1 phi, 2 v})
    int F = 0;
    // fronts in elimination order.  Slots: [0, npad) pivots (np real ones, then dummies = -1 up to a multiple of 3),
    // [npad, npad + nb) boundary variables, both in elimination order.
    std::vector<int> f_np, f_npad, f_nb, f_vars_off, f_vars;
    std::vector<long long> f_Loff, f_Uoff, f_Ioff;     // L panel (Rf x Cf), update matrix (ub x ub), inverse 9x9 diagonal blocks
    std::vector<int> f_parent, f_level, f_part;
    std::vector<int> f_child_off, f_children;
    std::vector<int> c_map_off, c_map;                 // per child entry: child boundary index -> parent slot
    // original entries of J^T W J that are first touched by this front: 3x3 block at (row slot, col slot) read from
    // (src >> 2) doubles into Hd (src & 2 == 0) or Ho (src & 2), transposed if src & 1; element (r, c) of the block
    // is at offset 9 r + c (9 c + r when transposed)
    std::vector<int> f_orig_off, orig_rs, orig_cs, orig_src;
    std::vector<int> level_off, level_fronts;
    std::vector<int> var_front, var_slot, var_pos;
    long long L_doubles = 0, U_doubles = 0, I_doubles = 0;
    int n_levels = 0, max_rows = 0, max_cols = 0, max_ub = 0, root_pivots = 0;
    int dense_root = -1;                               // front id of the dense root, or -1
    std::vector<int> root_slot;                        // [V] slot of a variable inside the dense root, or -1
    double factor_flops = 0;
    std::vector<int> scalars;                          // {F, n_levels, dense_root, max_rows, max_cols, max_ub, root_pivots} (introspection)
};

// `base` must have been filled by build_plan (pairs, band).  Returns 0 on success.
int build_plan3(const Plan& base, const int64_t* links, const SymbolicOpts& opts, Plan3& plan);

}  // namespace islam
