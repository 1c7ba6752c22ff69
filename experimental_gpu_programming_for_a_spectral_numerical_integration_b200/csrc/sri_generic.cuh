// sri_generic.cuh -- fused four-stage kernel for 17 <= N <= 64 Chebyshev nodes (M = N-1 <= 63 unknown nodes).
//
// One rod per CTA, the M x (M+1) quaternion system [C | b] resident in shared memory (M = 63: 130 KB, which is
// why the quaternion form matters: the dense real 252 x 252 operator of the reference, 508 KB, does not fit an SM).
// Thread (row i, column phase h), h = 0..3, updates the columns j = k+1+h, k+5+h, ... of row i in step k.
// Same algebra and pivoting as sri_fused16.cuh (Gauss-Jordan over the quaternions, implicit row pivoting).
//
// This kernel is LSU-bound (each quaternion update reads 64 B and writes 32 B of shared memory for 16 DFMA); the
// register-resident sliding-window form of sri_fused16.cuh is the template for the next optimisation round.
#pragma once
#include "sri_device.cuh"
#include "sri_fused16.cuh"  // FusedParams, fast_rcp, q_mul_tree

namespace sri {

// Packed operator tables for this kernel (doubles), stride R (32 or 64), zero padded.
struct OpsLayoutGeneric {
    int R;
    __host__ __device__ int Sp() const { return 0; }                 // [63][R]  Sp[j*R+i]  = (Dn_NN^-1)(i,j)
    __host__ __device__ int STt() const { return 63 * R; }           // [63][R]  STt[j*R+i] = (D_TT^-1)(i,j)
    __host__ __device__ int g() const { return 2 * 63 * R; }         // [R]
    __host__ __device__ int gT() const { return 2 * 63 * R + R; }    // [R]
    __host__ __device__ int DTI() const { return 2 * 63 * R + 2 * R; }
    __host__ __device__ int total() const { return 2 * 63 * R + 3 * R; }
};

__host__ __device__ inline int generic_row_doubles(int M) { return 4 * (M + 1) + 2; }  // +16 B skew: conflict-free columns

// shared memory footprint in doubles
__host__ __device__ inline size_t generic_smem_doubles(int M, int R) {
    OpsLayoutGeneric L{R};
    return (size_t)L.total() + (size_t)M * generic_row_doubles(M) + 4 * 64 /*ubuf*/ + 4 * 64 /*qnode*/ + 4 * 64 /*vec*/ +
           4 * 64 /*vec2*/ + 3 * 64 /*K*/ + 64 /*misc, keys, perm*/ + 64 + 64;
}

template <int R, bool SOLVE>
__global__ void __launch_bounds__(R * 4) generic_kernel(const FusedParams p) {
    extern __shared__ __align__(16) double smem[];
    const OpsLayoutGeneric L{R};
    const int M = p.M, N = p.N;
    const int rowd = generic_row_doubles(M);
    double* tab = smem;
    double* C = tab + L.total();              // [M][rowd]
    double* ubuf = C + (size_t)M * rowd;      // [64][4] pivot row copy
    double* qnode = ubuf + 256;               // [64][4] (spare)
    double* vec = qnode + 256;                // [64][4]
    double* vec2 = vec + 256;                 // [64][4]
    double* Kb = vec2 + 256;                  // [3][64]
    double* misc = Kb + 192;                  // F_tip 0..2, M_tip 3..5, r0 6..8, q0 9..12
    unsigned* keys = reinterpret_cast<unsigned*>(misc + 64);       // [R/32] warp maxima, [8] sing flag
    int* perm = reinterpret_cast<int*>(misc + 128);                // [64] pivot row of step k; [64..127] used flags

    const int t = threadIdx.x;
    const int i = t % R, h = t / R;
    const int lane = t & 31, warp = t >> 5;

    for (int idx = t; idx < L.total(); idx += blockDim.x) tab[idx] = p.ops[idx];
    __syncthreads();

    for (long long rod = blockIdx.x; rod < p.batch; rod += gridDim.x) {
        // ---- inputs ------------------------------------------------------------------------------------
        if (SOLVE && t < N) {
            const double* s = p.K + rod * 3 * N + t;
            Kb[t] = s[0]; Kb[64 + t] = s[N]; Kb[128 + t] = s[2 * N];
        }
        if (t < 3) {
            if (p.F_tip) misc[t] = p.F_tip[rod * 3 + t];
            if (p.M_tip) misc[3 + t] = p.M_tip[rod * 3 + t];
            misc[6 + t] = p.r0 ? p.r0[rod * 3 + t] : 0.0;
        }
        if (t < 4) misc[9 + t] = p.q0 ? p.q0[rod * 4 + t] : (t == 0 ? 1.0 : 0.0);
        if (t < 64) { perm[t] = 0; perm[64 + t] = (t >= M) ? 1 : 0; }
        if (t == 0) keys[8] = 0u;
        __syncthreads();
        quat q0; q0.w = misc[9]; q0.x = misc[10]; q0.y = misc[11]; q0.z = misc[12];

        quat q; q.w = 1.0; q.x = 0.0; q.y = 0.0; q.z = 0.0;
        if (SOLVE) {
        // ---- stage 1: assemble [C | b]: c_ij = delta_ij - 1/2 S_ij (0,K_j), b_i = g_i q0 -----------------
        if (i < M) {
            double* Ci = C + (size_t)i * rowd;
            for (int j = h; j <= M; j += 4) {
                quat c;
                if (j < M) {
                    const double s = -0.5 * tab[L.Sp() + j * R + i];
                    c.w = (j == i) ? 1.0 : 0.0; c.x = s * Kb[j]; c.y = s * Kb[64 + j]; c.z = s * Kb[128 + j];
                } else {
                    const double gi = tab[L.g() + i];
                    c.w = gi * q0.w; c.x = gi * q0.x; c.y = gi * q0.y; c.z = gi * q0.z;
                }
                st_quat(Ci + 4 * j, c);
            }
        }
        __syncthreads();

        // ---- Gauss-Jordan over the quaternions, implicit row pivoting ------------------------------------
        for (int k = 0; k < M; ++k) {
            // (1) pivot search among unused rows: column phase 0 threads hold one row each
            if (h == 0) {
                unsigned key = 0u;
                if (i < M && perm[64 + i] == 0) {
                    const quat c = ld_quat(C + (size_t)i * rowd + 4 * k);
                    const double nrm = fma(c.w, c.w, c.x * c.x) + fma(c.y, c.y, c.z * c.z);
                    key = (((unsigned)__double2hiint(nrm)) & 0xFFFFFFC0u) | (unsigned)(63 - i);
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const unsigned other = __shfl_xor_sync(0xffffffffu, key, off);
                    key = key > other ? key : other;
                }
                if (lane == 0) keys[warp] = key;
            }
            __syncthreads();
            unsigned key = keys[0];
            if (R == 64) { const unsigned k1 = keys[1]; key = key > k1 ? key : k1; }
            const int prow = 63 - (int)(key & 63u);
            const bool singular = (key >> 6) == 0u;
            // (2) copy the pivot row (columns k..M) out of the matrix
            if (!singular) {
                for (int j = k + t; j <= M; j += blockDim.x) st_quat(ubuf + 4 * j, ld_quat(C + (size_t)prow * rowd + 4 * j));
            } else if (t == 0 && keys[8] == 0u) {
                keys[8] = (unsigned)(k + 1);
            }
            __syncthreads();
            // (3) multipliers and rank-1 update
            if (!singular && i < M) {
                const quat piv = ld_quat(ubuf + 4 * k);
                const double inv = fast_rcp(fma(piv.w, piv.w, piv.x * piv.x) + fma(piv.y, piv.y, piv.z * piv.z));
                quat pinv; pinv.w = piv.w * inv; pinv.x = -piv.x * inv; pinv.y = -piv.y * inv; pinv.z = -piv.z * inv;
                double* Ci = C + (size_t)i * rowd;
                quat mlt;
                if (i == prow) { mlt.w = 1.0 - pinv.w; mlt.x = -pinv.x; mlt.y = -pinv.y; mlt.z = -pinv.z; }
                else mlt = q_mul_tree(pinv, ld_quat(Ci + 4 * k));
                for (int j = k + 1 + h; j <= M; j += 4) {
                    const quat u = ld_quat(ubuf + 4 * j);
                    quat c = ld_quat(Ci + 4 * j);
                    q_sub_mul(c, u, mlt);
                    st_quat(Ci + 4 * j, c);
                }
                if (i == prow && h == 0) { perm[k] = prow; perm[64 + prow] = 1; }
            }
            __syncthreads();
        }

        // ---- gather Q by node --------------------------------------------------------------------------
        if (t < M) q = ld_quat(C + (size_t)perm[t] * rowd + 4 * M);
        if (t == M) q = q0;
        if (p.info && t == 0) p.info[rod] = (int)keys[8];
        if (p.Q && t < M) {
            double* d = p.Q + rod * 4 * M + t;
            d[0] = q.w; d[M] = q.x; d[2 * M] = q.y; d[3 * M] = q.z;
        }
        } else {
            if (t == M) q = q0;
            if (p.Qin && t < M) {
                const double* s = p.Qin + rod * 4 * M + t;
                q.w = s[0]; q.x = s[M]; q.y = s[2 * M]; q.z = s[3 * M];
            }
        }
        if (p.r || p.n || p.m) {
            // ---- stage 2 (threads = nodes) -----------------------------------------------------------------
            if (t <= M) {
                double b0, b1, b2;
                if (p.Gamma) { const double* s = p.Gamma + rod * 3 * N + t; q_rotate(q, s[0], s[N], s[2 * N], b0, b1, b2); }
                else q_rotate_e1(q, b0, b1, b2);
                quat v; v.w = b0; v.x = b1; v.y = b2; v.z = 0.0;
                st_quat(vec + 4 * t, v);
            }
            const bool contract3 = (p.n || p.m) && p.fbar && !(!SOLVE && p.nin);
            double F0 = 0.0, F1 = 0.0, F2 = 0.0;
            if (p.n || p.m) { F0 = misc[0]; F1 = misc[1]; F2 = misc[2]; }
            if (contract3 && t < M) {
                const double dti = tab[L.DTI() + t];
                const double* s = p.fbar + rod * 3 * N + t + 1;
                quat v; v.w = -s[0] - dti * F0; v.x = -s[N] - dti * F1; v.y = -s[2 * N] - dti * F2; v.z = 0.0;
                st_quat(vec2 + 4 * t, v);
            }
            __syncthreads();
            if (p.r && t < M) {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                for (int j = 0; j < M; ++j) {
                    const double s = tab[L.Sp() + j * R + t];
                    a0 = fma(s, vec[4 * j], a0); a1 = fma(s, vec[4 * j + 1], a1); a2 = fma(s, vec[4 * j + 2], a2);
                }
                const double gi = tab[L.g() + t];
                a0 = fma(gi, misc[6], a0); a1 = fma(gi, misc[7], a1); a2 = fma(gi, misc[8], a2);
                double* d = p.r + rod * 3 * M + t;
                d[0] = a0; d[M] = a1; d[2 * M] = a2;
            }
            if (p.n || p.m) {
                // ---- stage 3: thread t = reduced row (node t+1) ---------------------------------------------
                double n0 = 0.0, n1 = 0.0, n2 = 0.0;
                if (t < M) {
                    if (!SOLVE && p.nin) {
                        const double* s = p.nin + rod * 3 * M + t; n0 = s[0]; n1 = s[M]; n2 = s[2 * M];
                    } else if (contract3) {
                        for (int j = 0; j < M; ++j) {
                            const double s = tab[L.STt() + j * R + t];
                            n0 = fma(s, vec2[4 * j], n0); n1 = fma(s, vec2[4 * j + 1], n1); n2 = fma(s, vec2[4 * j + 2], n2);
                        }
                    } else {
                        const double gi = tab[L.gT() + t];
                        n0 = gi * F0; n1 = gi * F1; n2 = gi * F2;
                    }
                    if (p.n) { double* d = p.n + rod * 3 * M + t; d[0] = n0; d[M] = n1; d[2 * M] = n2; }
                }
                if (p.m) {
                    // ---- stage 4 ---------------------------------------------------------------------------
                    __syncthreads();  // vec2 (stage 3 rhs) has been consumed
                    if (t < M) {
                        const double* rp = vec + 4 * (t + 1);
                        double l0 = 0.0, l1 = 0.0, l2 = 0.0;
                        if (p.lbar) { const double* s = p.lbar + rod * 3 * N + t + 1; l0 = s[0]; l1 = s[N]; l2 = s[2 * N]; }
                        const double dti = tab[L.DTI() + t];
                        const double c0 = rp[1] * n2 - rp[2] * n1, c1 = rp[2] * n0 - rp[0] * n2, c2 = rp[0] * n1 - rp[1] * n0;
                        quat v; v.w = -(c0 + l0) - dti * misc[3]; v.x = -(c1 + l1) - dti * misc[4]; v.y = -(c2 + l2) - dti * misc[5]; v.z = 0.0;
                        st_quat(vec2 + 4 * t, v);
                    }
                    __syncthreads();
                    if (t < M) {
                        double m0 = 0.0, m1 = 0.0, m2 = 0.0;
                        for (int j = 0; j < M; ++j) {
                            const double s = tab[L.STt() + j * R + t];
                            m0 = fma(s, vec2[4 * j], m0); m1 = fma(s, vec2[4 * j + 1], m1); m2 = fma(s, vec2[4 * j + 2], m2);
                        }
                        double* d = p.m + rod * 3 * M + t;
                        d[0] = m0; d[M] = m1; d[2 * M] = m2;
                    }
                }
            }
        }
        __syncthreads();  // shared buffers are reused by the next rod
    }
}

}  // namespace sri
