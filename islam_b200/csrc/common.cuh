// Shared device-side definitions for the PVGO kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/islam_pvgo.h"

namespace islam {

typedef islam_lm_state LMState;

// Device views of the symbolic plan (symbolic3.h), uploaded once per graph.
struct Front3Meta {
    const int* np;          // [F] real pivot variables
    const int* npad;        // [F] pivot slots (np rounded up to a multiple of 3; the extra ones are dummies)
    const int* nb;          // [F] boundary variables
    const int* vars_off;    // [F+1]
    const int* vars;        // pivot slots (dummy = -1) then boundary, elimination order; variable u = 3 pose + {tau, phi, v}
    const long long* Loff;  // [F] offset of the (3(npad+nb)+1) x 3npad column-major panel
    const long long* Uoff;  // [F] offset of the packed lower triangle of the (3nb+1)^2 update matrix
    const long long* Ioff;  // [F] offset of the npad/3 inverse 9x9 diagonal blocks
    const int* parent;      // [F] parent front or -1
    const int* child_off;   // [F+1]
    const int* children;    // child front ids
    const int* cmap_off;    // [nchildren_total+1]
    const int* cmap;        // child boundary index -> parent slot
    const int* orig_off;    // [F+1] original 3x3 blocks first touched by the front
    const int* orig_rs;     // row slot
    const int* orig_cs;     // column slot (a pivot)
    const int* orig_src;    // (offset << 2) | (2: Ho, 0: Hd) | (1: transposed)
    const unsigned short* dmap;   // per element of every child's packed U: destination in the parent's shared memory | nullptr
    const int* part;        // [F] owning window or -1 (shared)
    const long long* shared_off;  // [F] offset into the shared all-reduce buffer or -1
    int mypart;                   // this rank's window (multi-GPU)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// deterministic block reduction (fixed tree); result valid in thread 0
template <int THREADS> __device__ __forceinline__ double block_sum(double v, double* sh /* >= THREADS/32 */) {
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0;
    if (w == 0) {
        r = (lane < THREADS / 32) ? sh[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

}  // namespace islam
