// scale_from_disp_flow (/root/reference/dense_ba.py:88-176; call site TartanVO.py:159-171): the
// masked one-unknown least squares  s = sum(M w) / sum(M M)  that restores the metric scale of TartanVO's translation
// from stereo disparity (or depth) and optical flow, right before the PVGO back-end.
//
// The reference runs it per sample in a Python loop: ~40 elementwise torch kernels over the H x W grid, two boolean
// gathers and a tiny matmul.  Here the whole batch is ONE fused pass: every pixel is read once (disp | depth, 2 x flow,
// optional edge mask), the masks, the depth map and the four coefficients are formed in registers, and the two sums are
// reduced deterministically (fixed tree per CTA, fixed order over CTAs).  Bound: HBM — 12 (+1) bytes read and 6 bytes
// written per pixel.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace islam {

constexpr int SC_THREADS = 256;
constexpr int SC_PIX = 4;            // pixels per thread

__device__ __forceinline__ void quat_rot(const float* q, const float* p, float* o) {     // q = (x, y, z, w)
    const float tx = 2.f * (q[1] * p[2] - q[2] * p[1]), ty = 2.f * (q[2] * p[0] - q[0] * p[2]), tz = 2.f * (q[0] * p[1] - q[1] * p[0]);
    o[0] = p[0] + q[3] * tx + (q[1] * tz - q[2] * ty);
    o[1] = p[1] + q[3] * ty + (q[2] * tx - q[0] * tz);
    o[2] = p[2] + q[3] * tz + (q[0] * ty - q[1] * tx);
}

// GRAD: additionally reduce the nine sums the backward pass of s = num / den needs (d num / d a, d den / d a with a = K t_norm,
// and d num / d(left tangent of R = T.Inv().rotation())), so that autograd into `motion` (TartanVO.py:181 multiplies the
// normalised translation by the scale while the graph of the pose head is live) costs no second pass over the pixels.
constexpr int SC_NSUM = 12;
template <bool GRAD>
__global__ void __launch_bounds__(SC_THREADS)
k_scale_pixels(const float* __restrict__ disp, const float* __restrict__ flow, const float* __restrict__ motion,
               const float* __restrict__ intr, const float* __restrict__ baseline, const float* __restrict__ depth_in,
               const uint8_t* __restrict__ mask_in, const float* __restrict__ disp_th, int H, int W,
               float* __restrict__ z_out, uint8_t* __restrict__ mask_out, uint8_t* __restrict__ dmask_out,
               double* __restrict__ part /* [B][gridDim.x][SC_NSUM] */) {
    const int b = blockIdx.y, npix = H * W;
    const float fx = intr[4 * b], fy = intr[4 * b + 1], cx = intr[4 * b + 2], cy = intr[4 * b + 3], bl = baseline[b];
    // T.Inv(): R^T, -R^T t ; t_norm = normalize(t_inv)                                      dense_ba.py:144-146
    const float* mo = motion + 7 * (size_t)b;
    const float qi[4] = {-mo[3], -mo[4], -mo[5], mo[6]};
    float ti[3], tn[3];
    { float t[3] = {mo[0], mo[1], mo[2]}; quat_rot(qi, t, ti); ti[0] = -ti[0]; ti[1] = -ti[1]; ti[2] = -ti[2]; }
    const float nrm = fmaxf(sqrtf(ti[0] * ti[0] + ti[1] * ti[1] + ti[2] * ti[2]), 1e-12f);
    tn[0] = ti[0] / nrm; tn[1] = ti[1] / nrm; tn[2] = ti[2] / nrm;
    const float a0 = fx * tn[0] + cx * tn[2], a1 = fy * tn[1] + cy * tn[2], a2 = tn[2];     // a = K t_norm     :149
    const float th = disp_th[b];
    const float* dsp = disp + (size_t)b * npix;
    const float* f0 = flow + (size_t)b * 2 * npix;
    const float* f1 = f0 + npix;
    const float* dep = depth_in ? depth_in + (size_t)b * npix : nullptr;
    const uint8_t* min_ = mask_in ? mask_in + (size_t)b * npix : nullptr;
    double num = 0.0, den = 0.0, cnt = 0.0;
    double gs[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};      // d num / d a (3), d den / d a (3), d num / d delta_R (3)
#pragma unroll
    for (int k = 0; k < SC_PIX; ++k) {
        const int p = (blockIdx.x * SC_PIX + k) * SC_THREADS + threadIdx.x;
        if (p >= npix) continue;
        const int y = p / W, x = p - y * W;
        const float u = (float)x, v = (float)y, fl0 = f0[p], fl1 = f1[p];
        const float fu = fl0 + u, fv = fl1 + v;
        bool m = fu >= 0.f && fu <= (float)W && fv >= 0.f && fv <= (float)H && sqrtf(fl0 * fl0 + fl1 * fl1) > 0.f;   // :108-109
        if (min_) m = m && min_[p];
        bool dm;
        float z;
        if (!dep) {                                                                             // :115-124
            const float d = dsp[p];
            dm = (u - d) >= 0.f && (u - d) <= (float)W && d >= th;
            z = dm ? fx * bl / d : 0.f;
        } else {                                                                                // :126-132
            const float d = dep[p];
            dm = d <= fx * bl && d > 0.f;
            z = dm ? d : 0.f;
        }
        m = m && dm;
        z_out[(size_t)b * npix + p] = z;
        mask_out[(size_t)b * npix + p] = m;
        dmask_out[(size_t)b * npix + p] = dm;
        if (m) {
            const float P[3] = {(u - cx) / fx * z, (v - cy) / fy * z, z};                       // z K^-1 [u v 1]   :138-142
            float RP[3];
            quat_rot(qi, P, RP);
            const float b0 = fx * RP[0] + cx * RP[2], b1 = fy * RP[1] + cy * RP[2], b2 = RP[2];   // b = K R P        :150
            const float M1 = a2 * fu - a0, w1 = b0 - b2 * fu, M2 = a2 * fv - a1, w2 = b1 - b2 * fv;   // :153-156
            num += (double)M1 * w1 + (double)M2 * w2;
            den += (double)M1 * M1 + (double)M2 * M2;
            cnt += 1.0;
            if (GRAD) {
                // M = alpha a with alpha = [[-1, 0, fu], [0, -1, fv]];  w = beta b with beta = [[1, 0, -fu], [0, 1, -fv]]
                gs[0] -= (double)w1; gs[1] -= (double)w2; gs[2] += (double)fu * w1 + (double)fv * w2;
                gs[3] -= 2.0 * M1;   gs[4] -= 2.0 * M2;   gs[5] += 2.0 * ((double)fu * M1 + (double)fv * M2);
                // b = K (R P): R <- Exp(d) R moves R P by d x (R P), so d num / d d = (R P) x (K^T beta^T M)
                const float m0 = M1, m1 = M2, m2 = -fu * M1 - fv * M2;
                const float c0 = fx * m0, c1 = fy * m1, c2 = cx * m0 + cy * m1 + m2;
                gs[6] += (double)(RP[1] * c2 - RP[2] * c1);
                gs[7] += (double)(RP[2] * c0 - RP[0] * c2);
                gs[8] += (double)(RP[0] * c1 - RP[1] * c0);
            }
        }
    }
    __shared__ double sh[SC_THREADS / 32];
    num = block_sum<SC_THREADS>(num, sh);
    den = block_sum<SC_THREADS>(den, sh);
    cnt = block_sum<SC_THREADS>(cnt, sh);
    double* o = part + SC_NSUM * ((size_t)b * gridDim.x + blockIdx.x);
    if (threadIdx.x == 0) { o[0] = num; o[1] = den; o[2] = cnt; }
    if (GRAD) {
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const double v = block_sum<SC_THREADS>(gs[q], sh);
            if (threadIdx.x == 0) o[3 + q] = v;
        }
    }
}

__global__ void __launch_bounds__(SC_THREADS)
k_scale_finish(const double* __restrict__ part, int nblk, float* __restrict__ scale, int32_t* __restrict__ count,
               double* __restrict__ grad_sums /* [B][11] or nullptr */) {
    const int b = blockIdx.x;
    const int nsum = grad_sums ? SC_NSUM : 3;
    __shared__ double sh[SC_THREADS / 32];
    __shared__ double tot[SC_NSUM];
    for (int q = 0; q < nsum; ++q) {
        double v = 0.0;
        for (int k = threadIdx.x; k < nblk; k += SC_THREADS) v += part[SC_NSUM * ((size_t)b * nblk + k) + q];
        v = block_sum<SC_THREADS>(v, sh);
        if (threadIdx.x == 0) tot[q] = v;
    }
    if (threadIdx.x == 0) {
        scale[b] = (float)(tot[0] / tot[1]);                 // 1 / sum(M M) * M^T w   (dense_ba.py:166); 0/0 -> nan like torch
        if (count) count[b] = (int32_t)tot[2];               // the reference warns below 500 points (:134-135)
        if (grad_sums) {
            double* g = grad_sums + 11 * (size_t)b;
            g[0] = tot[0]; g[1] = tot[1];
            for (int q = 0; q < 9; ++q) g[2 + q] = tot[3 + q];
        }
    }
}

}  // namespace islam

extern "C" int64_t islam_scale_workspace_bytes(int32_t B, int32_t H, int32_t W) {
    const int64_t nblk = ((int64_t)H * W + islam::SC_THREADS * islam::SC_PIX - 1) / (islam::SC_THREADS * islam::SC_PIX);
    return B <= 0 || H <= 0 || W <= 0 ? -1 : islam::SC_NSUM * 8 * nblk * B;
}

extern "C" int islam_scale_from_disp_flow(const float* disp, const float* flow, const float* motion, const float* intr,
                                          const float* baseline, const float* depth, const uint8_t* mask_in,
                                          const float* disp_th, int32_t B, int32_t H, int32_t W, float* scale, float* z,
                                          uint8_t* mask, uint8_t* depth_mask, int32_t* mask_count, double* grad_sums,
                                          void* workspace, void* stream) {
    if (B <= 0 || H <= 0 || W <= 0 || !flow || !motion || !intr || !baseline || !disp_th || (!disp && !depth) || !scale || !z ||
        !mask || !depth_mask || !workspace)
        return -1;
    using namespace islam;
    const int nblk = (H * W + SC_THREADS * SC_PIX - 1) / (SC_THREADS * SC_PIX);
    cudaStream_t s = (cudaStream_t)stream;
    if (grad_sums)
        k_scale_pixels<true><<<dim3(nblk, B), SC_THREADS, 0, s>>>(disp, flow, motion, intr, baseline, depth, mask_in, disp_th, H, W, z,
                                                                  mask, depth_mask, (double*)workspace);
    else
        k_scale_pixels<false><<<dim3(nblk, B), SC_THREADS, 0, s>>>(disp, flow, motion, intr, baseline, depth, mask_in, disp_th, H, W, z,
                                                                   mask, depth_mask, (double*)workspace);
    k_scale_finish<<<B, SC_THREADS, 0, s>>>((const double*)workspace, nblk, scale, mask_count, grad_sums);
    return (int)cudaGetLastError();
}
