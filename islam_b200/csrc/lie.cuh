// SO(3)/SE(3) maps for the PVGO kernels (sm_100a).  Semantics follow PyPose's LieTensor as used at
// /root/reference/pvgo.py:36-39,45-48:
//   SE3 = [t(3), q = (x,y,z,w)],  se3 = [tau(3), phi(3)],  left perturbation X <- Exp(d) X.
// All 3x3 / 6x6 matrices are row-major.  Templated on the scalar so the same code serves the float32
// linearisation and the float64 retraction.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace islam {

template <typename T> struct Num;
template <> struct Num<float> {
    // below this angle the closed forms lose digits to cancellation -> 4-term Taylor series
    static __host__ __device__ constexpr float taylor() { return 0.5f; }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ void sincos_(float x, float* s, float* c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float atan_(float x) { return atanf(x); }
    static __device__ __forceinline__ float abs_(float x) { return fabsf(x); }
};
template <> struct Num<double> {
    static __host__ __device__ constexpr double taylor() { return 0.05; }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ void sincos_(double x, double* s, double* c) { sincos(x, s, c); }
    static __device__ __forceinline__ double atan_(double x) { return atan(x); }
    static __device__ __forceinline__ double abs_(double x) { return fabs(x); }
};

// ------------------------------------------------------------------------------------------- 3-vectors
template <typename T> __device__ __forceinline__ void cross3(const T* a, const T* b, T* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

// C = A * B (3x3 row-major)
template <typename T> __device__ __forceinline__ void mat3_mul(const T* A, const T* B, T* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

// ------------------------------------------------------------------------------------------- quaternions
template <typename T> __device__ __forceinline__ void q_mul(const T* a, const T* b, T* o) {
    T x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    T y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    T z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    T w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
template <typename T> __device__ __forceinline__ void q_inv(const T* a, T* o) {
    o[0] = -a[0]; o[1] = -a[1]; o[2] = -a[2]; o[3] = a[3];
}
// R(q) p = p + w t + v x t,  t = 2 v x p
template <typename T> __device__ __forceinline__ void q_rot(const T* q, const T* p, T* o) {
    T t[3], u[3];
    cross3(q, p, t);
    t[0] *= T(2); t[1] *= T(2); t[2] *= T(2);
    cross3(q, t, u);
    T r0 = p[0] + q[3] * t[0] + u[0], r1 = p[1] + q[3] * t[1] + u[1], r2 = p[2] + q[3] * t[2] + u[2];
    o[0] = r0; o[1] = r1; o[2] = r2;
}
template <typename T> __device__ __forceinline__ void q_matrix(const T* q, T* R) {
    T x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

// ------------------------------------------------------------------------------------------- SO3 maps
// so3.Exp: q = [sin(th/2)/th * phi, cos(th/2)]
template <typename T> __device__ __forceinline__ void so3_exp(const T* phi, T* q) {
    T th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
    T th = Num<T>::sqrt_(th2);
    T s, c;
    Num<T>::sincos_(T(0.5) * th, &s, &c);
    T k = (th < Num<T>::taylor())
              ? T(0.5) - th2 * (T(1.0 / 48.0) - th2 * (T(1.0 / 3840.0) - th2 * T(1.0 / 645120.0)))
              : s / th;
    q[0] = k * phi[0]; q[1] = k * phi[1]; q[2] = k * phi[2]; q[3] = c;
}
// SO3.Log: phi = 2 atan(|v|/w)/|v| * v   (q and -q map to the same phi, |phi| <= pi)
template <typename T> __device__ __forceinline__ void so3_log(const T* q, T* phi) {
    T n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    T n = Num<T>::sqrt_(n2);
    T w = q[3];
    T f;
    if (n < T(1e-6)) {
        f = T(2) / w - T(2.0 / 3.0) * n2 / (w * w * w);
    } else if (Num<T>::abs_(w) < T(1e-30)) {
        f = T(3.14159265358979323846) / n;
    } else {
        f = T(2) * Num<T>::atan_(n / w) / n;
    }
    phi[0] = f * q[0]; phi[1] = f * q[1]; phi[2] = f * q[2];
}

// coefficients of Jl = I + a K + b K^2 and of Jl^-1 = I - K/2 + c K^2
template <typename T> __device__ __forceinline__ void so3_coefs(T th2, T* a, T* b, T* c) {
    T th = Num<T>::sqrt_(th2);
    if (th < Num<T>::taylor()) {
        *a = T(0.5) - th2 * (T(1.0 / 24.0) - th2 * (T(1.0 / 720.0) - th2 * T(1.0 / 40320.0)));
        *b = T(1.0 / 6.0) - th2 * (T(1.0 / 120.0) - th2 * (T(1.0 / 5040.0) - th2 * T(1.0 / 362880.0)));
        *c = T(1.0 / 12.0) + th2 * (T(1.0 / 720.0) + th2 * (T(1.0 / 30240.0) + th2 * T(1.0 / 1209600.0)));
    } else {
        T s, co;
        Num<T>::sincos_(th, &s, &co);
        *a = (T(1) - co) / th2;
        *b = (th - s) / (th2 * th);
        *c = T(1) / th2 - (T(1) + co) / (T(2) * th * s);
    }
}
// M = I + a K + b K^2 with K = [phi]x   (row-major 3x3)
template <typename T> __device__ __forceinline__ void skew_poly(const T* p, T a, T b, T* M) {
    T x = p[0], y = p[1], z = p[2];
    T xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
    M[0] = T(1) - b * (yy + zz); M[1] = -a * z + b * xy;      M[2] = a * y + b * xz;
    M[3] = a * z + b * xy;       M[4] = T(1) - b * (xx + zz); M[5] = -a * x + b * yz;
    M[6] = -a * y + b * xz;      M[7] = a * x + b * yz;       M[8] = T(1) - b * (xx + yy);
}
template <typename T> __device__ __forceinline__ void so3_Jl(const T* phi, T* J) {
    T a, b, c;
    so3_coefs(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2], &a, &b, &c);
    skew_poly(phi, a, b, J);
}
template <typename T> __device__ __forceinline__ void so3_Jl_inv(const T* phi, T* J) {
    T a, b, c;
    so3_coefs(phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2], &a, &b, &c);
    skew_poly(phi, T(-0.5), c, J);
}

// ------------------------------------------------------------------------------------------- SE3 maps
template <typename T> __device__ __forceinline__ void se3_mul(const T* A, const T* B, T* O) {
    T r[3];
    q_rot(A + 3, B, r);
    T q[4];
    q_mul(A + 3, B + 3, q);
    O[0] = A[0] + r[0]; O[1] = A[1] + r[1]; O[2] = A[2] + r[2];
    O[3] = q[0]; O[4] = q[1]; O[5] = q[2]; O[6] = q[3];
}
template <typename T> __device__ __forceinline__ void se3_inv(const T* X, T* O) {
    T qi[4], r[3];
    q_inv(X + 3, qi);
    q_rot(qi, X, r);
    O[0] = -r[0]; O[1] = -r[1]; O[2] = -r[2];
    O[3] = qi[0]; O[4] = qi[1]; O[5] = qi[2]; O[6] = qi[3];
}
// se3.Exp: t = Jl(phi) tau ; q = so3.Exp(phi)
template <typename T> __device__ __forceinline__ void se3_exp(const T* xi, T* X) {
    T J[9];
    so3_Jl(xi + 3, J);
    X[0] = J[0] * xi[0] + J[1] * xi[1] + J[2] * xi[2];
    X[1] = J[3] * xi[0] + J[4] * xi[1] + J[5] * xi[2];
    X[2] = J[6] * xi[0] + J[7] * xi[1] + J[8] * xi[2];
    so3_exp(xi + 3, X + 3);
}
// SE3.Log: phi = Log(q) ; tau = Jl^-1(phi) t.   Also returns Jl^-1(phi) (needed by the 6x6 Jacobian).
template <typename T> __device__ __forceinline__ void se3_log(const T* X, T* xi, T* Ji) {
    so3_log(X + 3, xi + 3);
    so3_Jl_inv(xi + 3, Ji);
    xi[0] = Ji[0] * X[0] + Ji[1] * X[1] + Ji[2] * X[2];
    xi[1] = Ji[3] * X[0] + Ji[4] * X[1] + Ji[5] * X[2];
    xi[2] = Ji[6] * X[0] + Ji[7] * X[1] + Ji[8] * X[2];
}

// Barfoot's Q(xi) (upper-right block of the 6x6 left Jacobian), row-major 3x3
template <typename T> __device__ __forceinline__ void se3_Q(const T* xi, T* Q) {
    const T* tau = xi;
    const T* phi = xi + 3;
    T th2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
    T th = Num<T>::sqrt_(th2);
    T c1, c2, c3;
    if (th < Num<T>::taylor()) {
        c1 = T(1.0 / 6.0) - th2 * (T(1.0 / 120.0) - th2 * (T(1.0 / 5040.0) - th2 * T(1.0 / 362880.0)));
        c2 = T(1.0 / 24.0) - th2 * (T(1.0 / 720.0) - th2 * (T(1.0 / 40320.0) - th2 * T(1.0 / 3628800.0)));
        c3 = T(1.0 / 120.0) - th2 * (T(1.0 / 2520.0) - th2 * (T(1.0 / 120960.0) - th2 * T(1.0 / 9979200.0)));
    } else {
        T s, c;
        Num<T>::sincos_(th, &s, &c);
        T th4 = th2 * th2;
        c1 = (th - s) / (th2 * th);
        c2 = (th2 + T(2) * c - T(2)) / (T(2) * th4);
        c3 = (T(2) * th - T(3) * s + th * c) / (T(2) * th4 * th);
    }
    T Tm[9] = {0, -tau[2], tau[1], tau[2], 0, -tau[0], -tau[1], tau[0], 0};
    T K[9] = {0, -phi[2], phi[1], phi[2], 0, -phi[0], -phi[1], phi[0], 0};
    T KT[9], TK[9], KTK[9], KKT[9], TKK[9], KTKK[9], KKTK[9];
    mat3_mul(K, Tm, KT);
    mat3_mul(Tm, K, TK);
    mat3_mul(KT, K, KTK);
    mat3_mul(K, KT, KKT);
    mat3_mul(TK, K, TKK);
    mat3_mul(KTK, K, KTKK);
    mat3_mul(K, KTK, KKTK);
#pragma unroll
    for (int i = 0; i < 9; ++i)
        Q[i] = T(0.5) * Tm[i] + c1 * (KT[i] + TK[i] + KTK[i]) + c2 * (KKT[i] + TKK[i] - T(3) * KTK[i]) +
               c3 * (KTKK[i] + KKTK[i]);
}

}  // namespace islam
