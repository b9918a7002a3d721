// Device-resident Levenberg-Marquardt controller: trust-region damping, accept / roll back, plateau scheduler.
// Restates pp.optim.LM.step + strategy.TrustRegion.update + scheduler.StopOnPlateau.step as configured at
// /root/reference/pvgo.py:169-180 with the control flow kept on the GPU: every kernel of a
// "try" is predicated on LMState flags, so the host never has to wait for a verdict between tries.
#pragma once
#include "common.cuh"
#include "lie.cuh"

namespace islam {

__global__ void k_begin_try(LMState* st) {
    cudaGridDependencySynchronize();           // PDL: only the launch latency overlaps the previous kernel
    cudaTriggerProgrammaticLaunchCompletion();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (!st->continual) { st->active = 0; return; }
    st->active = 1;
    st->do_lin = st->need_linearize;
    st->chol_fail = 0;
    st->tries_total += 1;
}

// deterministic sum of per-block partials: out[0] = sum part[2k] (sum r^2), out[1] = sum part[2k+1] (quality term)
__global__ void __launch_bounds__(256) k_reduce2(const LMState* __restrict__ st, const double* __restrict__ part, int nparts,
                                                 double* __restrict__ out, int lin_only) {
    if (!st->active || (lin_only && !st->do_lin)) return;
    __shared__ double sh[8], sh2[8];
    double s = 0.0, q = 0.0;
    for (int k = threadIdx.x; k < nparts; k += 256) { s += part[2 * k]; q += part[2 * k + 1]; }
    s = block_sum<256>(s, sh);
    q = block_sum<256>(q, sh2);
    // multi-GPU trial sums (lin_only == 0 is only used there): a Cholesky that failed on THIS rank (a private front, or the
    // block of the distributed dense root this rank owns) must fail the try on EVERY rank, or the ranks take different
    // decisions and the next collective never completes.  The flag travels as a NaN in the sum (lm_control_shared).
    if (threadIdx.x == 0) { out[0] = (!lin_only && st->chol_fail) ? __longlong_as_double(0x7ff8000000000000LL) : s; out[1] = q; }
}

// after the (possibly all-reduced) linearisation loss is known
__device__ __forceinline__ void lm_begin_step_b(LMState* st, const double* __restrict__ lin_sum) {
    if (!st->do_lin) return;
    st->lin_loss = lin_sum[0];
    if (!st->loss_valid) { st->loss = lin_sum[0]; st->loss_valid = 1; }      // first call: loss = model.loss()
    st->last = st->loss;
}
__global__ void k_begin_step_b(LMState* st, const double* __restrict__ lin_sum) {
    if (threadIdx.x != 0 || !st->active) return;
    lm_begin_step_b(st, lin_sum);
}

// opens the step / the try in ONE launch: deterministic sum of the linearisation partials, cumulative damping of the
// (clamped) diagonal (A.4) and, on a single GPU (with_b), the bookkeeping that needs the linearisation loss
__global__ void __launch_bounds__(256) k_begin_step(LMState* st, const double* __restrict__ part, int nparts,
                                                    double* __restrict__ lin_sum, int with_b) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (!st->active) return;
    __shared__ double sh[8], sh2[8];
    if (st->do_lin) {
        double s = 0.0, q = 0.0;
        for (int k = threadIdx.x; k < nparts; k += 256) { s += part[2 * k]; q += part[2 * k + 1]; }
        s = block_sum<256>(s, sh);
        q = block_sum<256>(q, sh2);
        if (threadIdx.x == 0) { lin_sum[0] = s; lin_sum[1] = q; }
    }
    if (threadIdx.x != 0) return;
    if (st->do_lin) { st->reject_count = 0; st->diag_scale = 1.0; }
    st->diag_scale *= (1.0 + st->damping);                                   // A.diag += A.diag * damping
    if (with_b) lm_begin_step_b(st, lin_sum);
}

// nodes <- Exp(d[:6]) nodes ; vels <- vels + d[6:9]   (LieTensor.add_ / update_parameter, A.1/A.4)
__global__ void __launch_bounds__(128)
k_retract(const LMState* __restrict__ st, float* nodes0, float* nodes1, float* vels0, float* vels1,
          const double* __restrict__ D, int N) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (!st->active) return;
    int cur = st->cur;
    const float* ns = cur ? nodes1 : nodes0;
    float* nd = cur ? nodes0 : nodes1;
    const float* vs = cur ? vels1 : vels0;
    float* vd = cur ? vels0 : vels1;
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double xi[6], X[7], E[7], O[7];
#pragma unroll
    for (int k = 0; k < 6; ++k) xi[k] = D[9 * (size_t)n + k];
#pragma unroll
    for (int k = 0; k < 7; ++k) X[k] = (double)ns[7 * (size_t)n + k];
    se3_exp(xi, E);
    se3_mul(E, X, O);
#pragma unroll
    for (int k = 0; k < 7; ++k) nd[7 * (size_t)n + k] = (float)O[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) vd[3 * (size_t)n + k] = (float)((double)vs[3 * (size_t)n + k] + D[9 * (size_t)n + 6 + k]);
}

__device__ __forceinline__ void lm_end_step(LMState* st, const islam_lm_params& p) {
    st->need_linearize = 1;
    st->steps_done += 1;
    if (st->steps_done >= p.max_steps) st->continual = 0;
    if (p.use_scheduler) {                                   // StopOnPlateau.step (A.4)
        if ((st->last - st->loss) < p.decreasing) st->patience_count += 1;
        else st->patience_count = 0;
        if (st->patience_count >= p.patience) st->continual = 0;
        if (st->reject_count >= p.reject) st->continual = 0;
    }
}

// trial loss + trust-region update + accept / roll back
__device__ __forceinline__ void lm_control(LMState* st, const islam_lm_params* __restrict__ pp, double s, double q) {
    const islam_lm_params p = *pp;
    st->loss_trial = s;
    if (st->chol_fail) {            // "Linear solver failed. Breaking optimization step..." : params untouched
        st->info = 1;
        st->loss = st->last;
        st->accepted_last = 0;
        lm_end_step(st, p);
        return;
    }
    const double last = st->last, loss = s;
    const double denom = -q;                                  // -((J D)^T (2 R + J D)), unweighted
    const double quality = (last - loss) / denom;
    st->denom = denom;
    st->quality = quality;
    double radius = 1.0 / st->damping, down = st->down;
    if (quality > p.high) { radius *= p.up; down = p.down; }
    else if (quality > p.low) { down = p.down; }
    else { radius *= down; down *= p.factor; }
    down = fmax(p.tr_min, fmin(down, p.tr_max));
    radius = fmax(p.tr_min, fmin(radius, p.tr_max));
    st->down = down;
    st->radius = radius;
    st->damping = 1.0 / radius;
    if (last < loss && st->reject_count < p.reject) {        // reject: roll back (the trial buffer is dropped)
        st->loss = last;
        st->reject_count += 1;
        st->need_linearize = 0;
        st->accepted_last = 0;
    } else {                                                  // accept: the trial buffer becomes the state
        st->cur ^= 1;
        st->loss = loss;
        st->accepted_last = 1;
        lm_end_step(st, p);
    }
}
// multi-GPU: the summed trial loss is NaN when any rank's factorisation failed (k_reduce2 / k_end_try_p2p)
__device__ __forceinline__ void lm_control_shared(LMState* st, const islam_lm_params* __restrict__ pp, double s, double q) {
    if (s != s) st->chol_fail = 1;
    lm_control(st, pp, s, q);
}
__global__ void k_lm_control(LMState* st, const islam_lm_params* __restrict__ pp, const double* __restrict__ sums) {
    if (threadIdx.x != 0 || !st->active) return;
    lm_control_shared(st, pp, sums[0], sums[1]);
}

// single GPU: closes the try in ONE launch — deterministic sum of the trial partials, then the controller
__global__ void __launch_bounds__(256) k_end_try(LMState* st, const islam_lm_params* __restrict__ pp,
                                                 const double* __restrict__ part, int nparts, double* __restrict__ sums) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (!st->active) return;
    __shared__ double sh[8], sh2[8];
    double s = 0.0, q = 0.0;
    for (int k = threadIdx.x; k < nparts; k += 256) { s += part[2 * k]; q += part[2 * k + 1]; }
    s = block_sum<256>(s, sh);
    q = block_sum<256>(q, sh2);
    if (threadIdx.x != 0) return;
    sums[0] = s; sums[1] = q;
    lm_control(st, pp, s, q);
}

// multi-GPU: closes the try in ONE launch and WITHOUT a second collective.  Every rank sums its trial partials, writes the
// two doubles (trial sum r^2, quality term) straight into every peer's mailbox over NVLink (CUDA IPC peer memory), waits
// for the G messages of this try, adds them in rank order (identical, deterministic result everywhere) and runs the
// controller.  The mailbox is double-buffered on the try sequence number: a rank can run at most one exchange ahead of a
// peer, because it cannot finish exchange k+1 before that peer has sent its message k+1, i.e. after it consumed message k.
// mailbox layout (per rank, in its own device memory): [2 slots][G senders] x {s, q, seq, pad} (4 x 8 bytes)
__global__ void __launch_bounds__(256) k_end_try_p2p(LMState* st, const islam_lm_params* __restrict__ pp,
                                                     const double* __restrict__ part, int nparts, double* __restrict__ sums,
                                                     unsigned long long* const* __restrict__ peers,
                                                     unsigned long long* __restrict__ seq_ctr, int rank, int G) {
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (!st->active) return;
    __shared__ double sh[8], sh2[8], in_s[64], in_q[64];
    __shared__ unsigned long long seq_s;
    __shared__ double my[2];
    __shared__ int failed;
    const int tid = threadIdx.x;
    double s = 0.0, q = 0.0;
    for (int k = tid; k < nparts; k += 256) { s += part[2 * k]; q += part[2 * k + 1]; }
    s = block_sum<256>(s, sh);
    q = block_sum<256>(q, sh2);
    if (tid == 0) { my[0] = st->chol_fail ? __longlong_as_double(0x7ff8000000000000LL) : s; my[1] = q; seq_s = ++(*seq_ctr); failed = 0; }
    __syncthreads();
    const unsigned long long seq = seq_s;
    const int slot = (int)(seq & 1);
    if (tid < G) {                                             // one thread per peer: payload, fence, then the flag
        volatile unsigned long long* m = peers[tid] + (size_t)(slot * G + rank) * 4;
        m[0] = (unsigned long long)__double_as_longlong(my[0]);
        m[1] = (unsigned long long)__double_as_longlong(my[1]);
        __threadfence_system();
        m[2] = seq;
    }
    if (tid < G) {                                             // one thread per sender
        volatile unsigned long long* m = peers[rank] + (size_t)(slot * G + tid) * 4;
        const long long t0 = clock64();
        while (m[2] != seq) {
            if (clock64() - t0 > (4LL << 30)) { failed = 1; break; }        // ~2 s: a peer is gone
        }
        __threadfence_system();
        in_s[tid] = __longlong_as_double((long long)m[0]);
        in_q[tid] = __longlong_as_double((long long)m[1]);
    }
    __syncthreads();
    if (tid != 0) return;
    if (failed) { st->info = 2; st->continual = 0; return; }
    double S = 0.0, Q = 0.0;
    for (int r = 0; r < G; ++r) { S += in_s[r]; Q += in_q[r]; }
    sums[0] = S; sums[1] = Q;
    lm_control_shared(st, pp, S, Q);
}

// lm_reset without a host round trip: the parameters travel as a kernel argument into device memory, the state is
// re-initialised in place (the current/trial buffer index survives)
__global__ void k_lm_reset(LMState* st, islam_lm_params* dst, islam_lm_params p) {
    if (threadIdx.x != 0) return;
    *dst = p;
    const int cur = st->cur;
    LMState z;
    memset(&z, 0, sizeof(z));
    z.cur = cur;
    z.damping = 1.0 / p.radius; z.radius = p.radius; z.down = p.down; z.diag_scale = 1.0;
    z.need_linearize = 1; z.continual = 1;
    *st = z;
}

// copy of the CURRENT state buffer (selected on the device, so no host synchronisation is needed)
__global__ void __launch_bounds__(128)
k_copy_state(const LMState* __restrict__ st, const float* __restrict__ n0, const float* __restrict__ n1,
             const float* __restrict__ v0, const float* __restrict__ v1, int N, float* __restrict__ nout, float* __restrict__ vout) {
    const float* ns = st->cur ? n1 : n0;
    const float* vs = st->cur ? v1 : v0;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (nout && i < 7 * N) nout[i] = ns[i];
    if (vout && i < 3 * N) vout[i] = vs[i];
}

}  // namespace islam
