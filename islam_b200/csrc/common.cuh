// Shared device-side definitions for the PVGO kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/islam_pvgo.h"

namespace islam {

typedef islam_lm_state LMState;

// Device views of the symbolic plan (symbolic.h), uploaded once per graph.
struct FrontMeta {
    const int* np;          // [F] pivot poses
    const int* nb;          // [F] boundary poses
    const int* nodes_off;   // [F+1]
    const int* nodes;       // pivots then boundary (elimination order)
    const long long* Loff;  // [F] offset of the (9(np+nb)+1) x 9np column-major panel
    const long long* Uoff;  // [F] offset of the (9nb+1)^2 column-major update matrix
    const int* child_off;   // [F+1]
    const int* children;    // child front ids
    const int* cinv_off;    // [nchildren_total+1]
    const int* cinv;        // parent slot -> child boundary index | -1
    const int* hmap_off;    // [F+1]
    const int* hmap;        // (np+nb) x np : (pair<<1|transpose) | -1
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
