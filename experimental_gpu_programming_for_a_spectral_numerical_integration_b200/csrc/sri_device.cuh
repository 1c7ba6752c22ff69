// sri_device.cuh -- device-side building blocks shared by the sm_100a kernels.
//
// Algebra used by stage 1 (see DESIGN.md "Stage 1 over the quaternions"):
//   The reference's 4x4 block A(K) (main.cpp:72-75) is right-multiplication by the pure quaternion (0,K):
//   A(K) Q = Q (x) (0,K).  After left-preconditioning with Dn_NN^-1 =: S (strain independent, cached) the
//   4M x 4M real collocation system of main.cpp:103-113 is the M x M system over the quaternions
//        sum_j Q_j (x) c_ij = g_i q0,     c_ij = delta_ij - 1/2 S_ij (0,K_j),   g = -S Dn_IN  (= 1 up to rounding)
//   which Gauss-Jordan elimination with row pivoting solves in 1/4 of the flops and 1/4 of the storage of the
//   dense real LU.  Same solution as the reference's A_NN.inverse()*(b-ivp) to ~1e-15.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sri {

struct quat {
    double w, x, y, z;
};

// c -= u (x) m   (Hamilton product, u on the left)
__device__ __forceinline__ void q_sub_mul(quat& c, const quat& u, const quat& m) {
    c.w = fma(-u.w, m.w, c.w); c.x = fma(-u.w, m.x, c.x); c.y = fma(-u.w, m.y, c.y); c.z = fma(-u.w, m.z, c.z);
    c.w = fma(u.x, m.x, c.w);  c.x = fma(-u.x, m.w, c.x); c.y = fma(u.x, m.z, c.y);  c.z = fma(-u.x, m.y, c.z);
    c.w = fma(u.y, m.y, c.w);  c.x = fma(-u.y, m.z, c.x); c.y = fma(-u.y, m.w, c.y); c.z = fma(u.y, m.x, c.z);
    c.w = fma(u.z, m.z, c.w);  c.x = fma(u.z, m.y, c.x);  c.y = fma(-u.z, m.x, c.y); c.z = fma(-u.z, m.w, c.z);
}

// One of the four FMA levels of c -= u (x) m (level l consumes component l of u); the levels of several
// independent updates can be interleaved to cover the FP64 pipe's dependent-issue latency.
template <int LEVEL>
__device__ __forceinline__ void q_sub_mul_level(quat& c, const quat& u, const quat& m) {
    if (LEVEL == 0) { c.w = fma(-u.w, m.w, c.w); c.x = fma(-u.w, m.x, c.x); c.y = fma(-u.w, m.y, c.y); c.z = fma(-u.w, m.z, c.z); }
    if (LEVEL == 1) { c.w = fma(u.x, m.x, c.w);  c.x = fma(-u.x, m.w, c.x); c.y = fma(u.x, m.z, c.y);  c.z = fma(-u.x, m.y, c.z); }
    if (LEVEL == 2) { c.w = fma(u.y, m.y, c.w);  c.x = fma(-u.y, m.z, c.x); c.y = fma(-u.y, m.w, c.y); c.z = fma(u.y, m.x, c.z); }
    if (LEVEL == 3) { c.w = fma(u.z, m.z, c.w);  c.x = fma(u.z, m.y, c.x);  c.y = fma(-u.z, m.x, c.y); c.z = fma(-u.z, m.w, c.z); }
}

// p (x) c
__device__ __forceinline__ quat q_mul(const quat& p, const quat& c) {
    quat r;
    r.w = p.w * c.w; r.x = p.w * c.x; r.y = p.w * c.y; r.z = p.w * c.z;
    r.w = fma(-p.x, c.x, r.w); r.x = fma(p.x, c.w, r.x); r.y = fma(-p.x, c.z, r.y); r.z = fma(p.x, c.y, r.z);
    r.w = fma(-p.y, c.y, r.w); r.x = fma(p.y, c.z, r.x); r.y = fma(p.y, c.w, r.y);  r.z = fma(-p.y, c.x, r.z);
    r.w = fma(-p.z, c.z, r.w); r.x = fma(-p.z, c.y, r.x); r.y = fma(p.z, c.x, r.y);  r.z = fma(p.z, c.w, r.z);
    return r;
}

__device__ __forceinline__ double q_norm2(const quat& c) {
    return fma(c.w, c.w, fma(c.x, c.x, fma(c.y, c.y, c.z * c.z)));
}

// R(q) * g with Eigen's un-normalised Quaterniond::toRotationMatrix formula (called at main.cpp:136).
__device__ __forceinline__ void q_rotate(const quat& q, double g0, double g1, double g2, double& o0, double& o1,
                                         double& o2) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    o0 = (1 - (tyy + tzz)) * g0 + (txy - twz) * g1 + (txz + twy) * g2;
    o1 = (txy + twz) * g0 + (1 - (txx + tzz)) * g1 + (tyz - twx) * g2;
    o2 = (txz - twy) * g0 + (tyz + twx) * g1 + (1 - (txx + tyy)) * g2;
}

// First column of R(q): R(q) * (1,0,0), the reference's hard-coded Gamma (main.cpp:136).
__device__ __forceinline__ void q_rotate_e1(const quat& q, double& o0, double& o1, double& o2) {
    const double ty = 2 * q.y, tz = 2 * q.z;
    o0 = 1 - (ty * q.y + tz * q.z);
    o1 = ty * q.x + tz * q.w;
    o2 = tz * q.x - ty * q.w;
}

// R(q)^T * v
__device__ __forceinline__ void q_rotate_T(const quat& q, double v0, double v1, double v2, double& o0, double& o1,
                                           double& o2) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    o0 = (1 - (tyy + tzz)) * v0 + (txy + twz) * v1 + (txz - twy) * v2;
    o1 = (txy - twz) * v0 + (1 - (txx + tzz)) * v1 + (tyz + twx) * v2;
    o2 = (txz + twy) * v0 + (tyz - twx) * v1 + (1 - (txx + tyy)) * v2;
}

__device__ __forceinline__ void st_quat(double* p, const quat& q) {
    reinterpret_cast<double2*>(p)[0] = make_double2(q.w, q.x);
    reinterpret_cast<double2*>(p)[1] = make_double2(q.y, q.z);
}
__device__ __forceinline__ quat ld_quat(const double* p) {
    const double2 a = reinterpret_cast<const double2*>(p)[0];
    const double2 b = reinterpret_cast<const double2*>(p)[1];
    quat q; q.w = a.x; q.x = a.y; q.y = b.x; q.z = b.y;
    return q;
}

// Philox4x32-10 (Salmon et al. 2011), counter = (rod index lo, hi, stream, 0), key = seed.
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in [0,1) from two 32-bit words
__host__ __device__ inline double u01_from_bits(uint32_t hi, uint32_t lo) {
    const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);
}

// Device-side state of one sri_newton_static_shape solve (reduce == NULL mode).
struct NewtonState {
    int done;             // set by newton_check_kernel once the tolerance is met: every later kernel of the solve exits at once
    int tested;           // number of convergence tests that have run
    unsigned long long singular;  // per-rod Newton systems with a zero pivot (their update was skipped), over all iterations
    double hist[64][2];   // [test] = sum g^2 and max |g| over all ranks
};

}  // namespace sri
