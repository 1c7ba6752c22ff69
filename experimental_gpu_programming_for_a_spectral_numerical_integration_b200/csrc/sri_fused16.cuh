// sri_fused16.cuh -- fused four-stage kernel for N <= 16 Chebyshev nodes (M = N-1 <= 15 unknown nodes).
//
// Mapping: TWO rods per warp.  Lanes 0..15 own rod A, lanes 16..31 rod B; lane `row` of a half-warp owns row
// `row` of that rod's M x M quaternion collocation operator (M quaternions + the right-hand side = 64 doubles,
// all in registers, every register index a compile-time constant).  Lane 15 of each half is padding.
//
// Replaces, per rod: updateA (main.cpp:55-88), A_NN.inverse()*(b-ivp) (main.cpp:113), updatePositionb
// (main.cpp:121-140), Dn_NN_inv*(b_NN-ivp) (main.cpp:172) and the two spec'd stages (rod_modeling.pdf 1.17-1.18).
#pragma once
#include "sri_device.cuh"

namespace sri {

constexpr int MP16 = 16;  // padded rows per rod in this kernel

// Packed operator tables (doubles), stride MP16, zero padded.  Built on the host by sri_api.cu.
struct OpsLayout16 {
    static constexpr int St = 0;                   // [15][16]  St[j*16+i]  = (Dn_NN^-1)(i,j)
    static constexpr int STt = St + 15 * MP16;     // [15][16]  STt[j*16+i] = (D_TT^-1)(i,j)
    static constexpr int g = STt + 15 * MP16;      // [16]  g  = -(Dn_NN^-1 Dn_IN)
    static constexpr int gT = g + MP16;            // [16]  gT = -(D_TT^-1 D_TI)
    static constexpr int DTI = gT + MP16;          // [16]  D_TI
    static constexpr int DnIN = DTI + MP16;        // [16]  Dn_IN
    static constexpr int total = DnIN + MP16;      // 544 doubles
};

struct FusedParams {
    long long batch;
    int N, M;
    const double* ops;  // OpsLayout16 (device)
    const double* K;
    const double* q0;
    const double* r0;
    const double* Gamma;
    const double* fbar;
    const double* lbar;
    const double* F_tip;
    const double* M_tip;
    const double* Qin;  // stage kernels only: quaternions computed by an earlier call
    const double* nin;  // stage kernels only: internal forces computed by an earlier call
    double* Q;
    double* r;
    double* n;
    double* m;
    int* info;
};

constexpr int kWarpScratch16 = 512;  // doubles of shared scratch per warp (4 KB)

// One Gauss-Jordan step over the quaternions with implicit row pivoting, one row per lane; K is a compile-time
// constant so that every access to c[] is a register.  pbuf: this half-warp's publish area, 2 x 16 slots x 4.
template <int K>
__device__ __forceinline__ void gj_step16(quat (&c)[15], quat& b, int row, double* pbuf, bool& used, int& mycol,
                                          int& sing) {
    // --- pivot search over the 16-lane segment: top bits of |c_ik|^2, row index in the low 4 bits
    const double nrm = q_norm2(c[K]);
    unsigned key = used ? 0u : ((((unsigned)__double2hiint(nrm)) & 0xFFFFFFF0u) | (unsigned)(15 - row));
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
        const unsigned other = __shfl_xor_sync(0xffffffffu, key, off);
        key = key > other ? key : other;
    }
    const int prow = 15 - (int)(key & 15u);
    const bool is_pivot = (row == prow) && !used;
    if ((key >> 4) == 0u && sing == 0) sing = K + 1;
    // --- the pivot lane publishes its row (columns K.. and the rhs)
    double* buf = pbuf + (K & 1) * (MP16 * 4);
    if (is_pivot) {
#pragma unroll
        for (int j = K; j < 15; ++j) st_quat(buf + 4 * j, c[j]);
        st_quat(buf + 4 * 15, b);
    }
    __syncwarp();
    // --- multiplier: m = c_pk^-1 (x) c_ik; the pivot lane takes 1 - c_pk^-1 so that the same update
    //     normalises its own row
    const quat piv = ld_quat(buf + 4 * K);
    const double inv = 1.0 / q_norm2(piv);
    quat pinv;
    pinv.w = piv.w * inv; pinv.x = -piv.x * inv; pinv.y = -piv.y * inv; pinv.z = -piv.z * inv;
    quat mlt = q_mul(pinv, c[K]);
    if (is_pivot) { mlt.w = 1.0 - pinv.w; mlt.x = -pinv.x; mlt.y = -pinv.y; mlt.z = -pinv.z; mycol = K; used = true; }
    // --- rank-1 update of the trailing columns and of the rhs
#pragma unroll
    for (int j = K + 1; j < 15; ++j) {
        const quat u = ld_quat(buf + 4 * j);
        q_sub_mul(c[j], u, mlt);
    }
    {
        const quat u = ld_quat(buf + 4 * 15);
        q_sub_mul(b, u, mlt);
    }
}

// Full elimination.  MS = 15: static size; MS = 0: runtime M <= 15 (steps K >= M are skipped; the padded rows and
// columns are zero).  On return b holds Q_{mycol}.
template <int MS>
__device__ __forceinline__ void gauss_jordan16(quat (&c)[15], quat& b, int M, int row, double* pbuf, int& mycol,
                                               int& sing) {
    bool used = (row >= M);
    mycol = row;
    sing = 0;
#define SRI_GJ_STEP(K) if (MS != 0 || K < M) gj_step16<K>(c, b, row, pbuf, used, mycol, sing);
    SRI_GJ_STEP(0) SRI_GJ_STEP(1) SRI_GJ_STEP(2) SRI_GJ_STEP(3) SRI_GJ_STEP(4)
    SRI_GJ_STEP(5) SRI_GJ_STEP(6) SRI_GJ_STEP(7) SRI_GJ_STEP(8) SRI_GJ_STEP(9)
    SRI_GJ_STEP(10) SRI_GJ_STEP(11) SRI_GJ_STEP(12) SRI_GJ_STEP(13) SRI_GJ_STEP(14)
#undef SRI_GJ_STEP
}

// SOLVE = true: all stages starting from the strain samples K.  SOLVE = false: the cached-operator stages only
// (position / stress / couple), reading Q (and optionally n) produced by an earlier call.
template <int MS, bool SOLVE>
__global__ void __launch_bounds__(128) fused16_kernel(const FusedParams p) {
    extern __shared__ __align__(16) double smem[];
    double* tab = smem;                                        // OpsLayout16::total doubles
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane >> 4, row = lane & 15;
    double* wscr = smem + OpsLayout16::total + warp * kWarpScratch16;
    double* pbuf = wscr + sub * 128;        // [2][16][4]   publish buffers of this half-warp
    double* qnode = wscr + 256 + sub * 64;  // [16][4]      quaternions by node (slot M = base node)
    double* vec = wscr + 384 + sub * 64;    // [16][4]      nodal 3-vectors
    double* vec2 = pbuf;                    // reused after the elimination
    double* vec3 = pbuf + 64;

    for (int i = threadIdx.x; i < OpsLayout16::total; i += blockDim.x) tab[i] = p.ops[i];
    __syncthreads();

    const int M = MS ? MS : p.M;
    const int N = M + 1;
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    const long long pairs = (p.batch + 1) >> 1;

    for (long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + warp; pair < pairs; pair += warps_total) {
        const long long rod = 2 * pair + sub;
        const bool live = rod < p.batch;

        // ---- inputs -----------------------------------------------------------------------------------
        quat q0; q0.w = 1.0; q0.x = 0.0; q0.y = 0.0; q0.z = 0.0;
        if (p.q0 && live) { const double* s = p.q0 + rod * 4; q0.w = s[0]; q0.x = s[1]; q0.y = s[2]; q0.z = s[3]; }
        quat q; q.w = 1.0; q.x = 0.0; q.y = 0.0; q.z = 0.0;
        if (SOLVE) {
            {
                quat ks; ks.w = 0.0; ks.x = 0.0; ks.y = 0.0; ks.z = 0.0;
                if (live && row < M) {
                    const double* s = p.K + rod * 3 * N + row;
                    ks.x = -0.5 * s[0]; ks.y = -0.5 * s[N]; ks.z = -0.5 * s[2 * N];
                }
                st_quat(vec + 4 * row, ks);
            }
            __syncwarp();

            // ---- stage 1: assemble c_ij = delta_ij - 1/2 S_ij (0,K_j) and eliminate ---------------------
            quat c[15], b;
#pragma unroll
            for (int j = 0; j < 15; ++j) {
                const double s = tab[OpsLayout16::St + j * MP16 + row];
                const quat kq = ld_quat(vec + 4 * j);
                c[j].w = (j == row) ? 1.0 : 0.0;
                c[j].x = s * kq.x; c[j].y = s * kq.y; c[j].z = s * kq.z;
            }
            {
                const double gi = tab[OpsLayout16::g + row];
                b.w = gi * q0.w; b.x = gi * q0.x; b.y = gi * q0.y; b.z = gi * q0.z;
            }
            int mycol, sing;
            gauss_jordan16<MS>(c, b, M, row, pbuf, mycol, sing);
            __syncwarp();
            if (row < M) st_quat(qnode + 4 * mycol, b);
            if (row == M) st_quat(qnode + 4 * M, q0);
            __syncwarp();
            if (p.info && live && row == 0) p.info[rod] = sing;

            if (row <= M) q = ld_quat(qnode + 4 * row);
            if (p.Q && live && row < M) {
                double* d = p.Q + rod * 4 * M + row;
                d[0] = q.w; d[M] = q.x; d[2 * M] = q.y; d[3 * M] = q.z;
            }
        } else {
            if (row == M) q = q0;
            if (p.Qin && live && row < M) {
                const double* s = p.Qin + rod * 4 * M + row;
                q.w = s[0]; q.x = s[M]; q.y = s[2 * M]; q.z = s[3 * M];
            }
        }
        if (!(p.r || p.n || p.m)) { __syncwarp(); continue; }

        // ---- stage 2: r = S (R(q) Gamma) + g r0 ---------------------------------------------------------
        double bv0 = 0.0, bv1 = 0.0, bv2 = 0.0;
        if (row <= M) {
            if (p.Gamma && live) {
                const double* s = p.Gamma + rod * 3 * N + row;
                q_rotate(q, s[0], s[N], s[2 * N], bv0, bv1, bv2);
            } else {
                q_rotate_e1(q, bv0, bv1, bv2);
            }
        }
        {
            quat t; t.w = bv0; t.x = bv1; t.y = bv2; t.z = 0.0;
            st_quat(vec + 4 * row, t);
        }
        __syncwarp();
        if (p.r) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
            for (int j = 0; j < 15; ++j) {
                if (MS == 0 && j >= M) break;
                const double s = tab[OpsLayout16::St + j * MP16 + row];
                const quat t = ld_quat(vec + 4 * j);
                a0 = fma(s, t.w, a0); a1 = fma(s, t.x, a1); a2 = fma(s, t.y, a2);
            }
            if (p.r0 && live) {
                const double gi = tab[OpsLayout16::g + row];
                const double* s = p.r0 + rod * 3;
                a0 = fma(gi, s[0], a0); a1 = fma(gi, s[1], a1); a2 = fma(gi, s[2], a2);
            }
            if (live && row < M) {
                double* d = p.r + rod * 3 * M + row;
                d[0] = a0; d[M] = a1; d[2 * M] = a2;
            }
        }
        if (!(p.n || p.m)) { __syncwarp(); continue; }

        // ---- stage 3: n = D_TT^-1 (-fbar - D_TI F_tip^T); lane `row` = reduced index (node row+1) -------
        double F0 = 0.0, F1 = 0.0, F2 = 0.0;
        if (live) { const double* s = p.F_tip + rod * 3; F0 = s[0]; F1 = s[1]; F2 = s[2]; }
        double n0, n1, n2;
        if (!SOLVE && p.nin) {
            n0 = 0.0; n1 = 0.0; n2 = 0.0;
            if (live && row < M) { const double* s = p.nin + rod * 3 * M + row; n0 = s[0]; n1 = s[M]; n2 = s[2 * M]; }
        } else if (p.fbar) {
            const double dti = tab[OpsLayout16::DTI + row];
            double f0 = 0.0, f1 = 0.0, f2 = 0.0;
            if (live && row < M) { const double* s = p.fbar + rod * 3 * N + row + 1; f0 = s[0]; f1 = s[N]; f2 = s[2 * N]; }
            quat t; t.w = -f0 - dti * F0; t.x = -f1 - dti * F1; t.y = -f2 - dti * F2; t.z = 0.0;
            st_quat(vec2 + 4 * row, t);
            __syncwarp();
            n0 = 0.0; n1 = 0.0; n2 = 0.0;
#pragma unroll
            for (int j = 0; j < 15; ++j) {
                if (MS == 0 && j >= M) break;
                const double s = tab[OpsLayout16::STt + j * MP16 + row];
                const quat u = ld_quat(vec2 + 4 * j);
                n0 = fma(s, u.w, n0); n1 = fma(s, u.x, n1); n2 = fma(s, u.y, n2);
            }
        } else {
            const double gi = tab[OpsLayout16::gT + row];
            n0 = gi * F0; n1 = gi * F1; n2 = gi * F2;
        }
        if (p.n && live && row < M) {
            double* d = p.n + rod * 3 * M + row;
            d[0] = n0; d[M] = n1; d[2 * M] = n2;
        }
        if (!p.m) { __syncwarp(); continue; }

        // ---- stage 4: m = D_TT^-1 (-(r' x n + lbar) - D_TI M_tip^T) ------------------------------------
        {
            double T0 = 0.0, T1 = 0.0, T2 = 0.0;
            if (live) { const double* s = p.M_tip + rod * 3; T0 = s[0]; T1 = s[1]; T2 = s[2]; }
            const int nb = (row < M) ? row + 1 : row;  // node of this reduced row
            const quat rp = ld_quat(vec + 4 * nb);     // r' at that node (w,x,y = components)
            double l0 = 0.0, l1 = 0.0, l2 = 0.0;
            if (p.lbar && live && row < M) { const double* s = p.lbar + rod * 3 * N + row + 1; l0 = s[0]; l1 = s[N]; l2 = s[2 * N]; }
            const double dti = tab[OpsLayout16::DTI + row];
            const double c0 = rp.x * n2 - rp.y * n1, c1 = rp.y * n0 - rp.w * n2, c2 = rp.w * n1 - rp.x * n0;
            quat t; t.w = -(c0 + l0) - dti * T0; t.x = -(c1 + l1) - dti * T1; t.y = -(c2 + l2) - dti * T2; t.z = 0.0;
            if (row >= M) { t.w = 0.0; t.x = 0.0; t.y = 0.0; }
            st_quat(vec3 + 4 * row, t);
            __syncwarp();
            double m0 = 0.0, m1 = 0.0, m2 = 0.0;
#pragma unroll
            for (int j = 0; j < 15; ++j) {
                if (MS == 0 && j >= M) break;
                const double s = tab[OpsLayout16::STt + j * MP16 + row];
                const quat u = ld_quat(vec3 + 4 * j);
                m0 = fma(s, u.w, m0); m1 = fma(s, u.x, m1); m2 = fma(s, u.y, m2);
            }
            if (live && row < M) {
                double* d = p.m + rod * 3 * M + row;
                d[0] = m0; d[M] = m1; d[2 * M] = m2;
            }
        }
        __syncwarp();
    }
}

}  // namespace sri
