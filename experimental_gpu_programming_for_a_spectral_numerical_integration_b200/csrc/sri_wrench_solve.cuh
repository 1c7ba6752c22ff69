// sri_wrench_solve.cuh -- local-frame statics solved directly (rod_modeling.pdf eqs. 1.29, 2.18), N <= 16.
//
//   N' = -K^ N - R^T fbar,   C' = -K^ C - Gamma^ N - R^T lbar,   N(1) = R(1)^T F_tip,  C(1) = R(1)^T M_tip
// collocated on the Chebyshev nodes with the tip node eliminated: the strain-DEPENDENT real operator
//   A = D_TT (x) I3 + blockdiag(K^_i),   3M x 3M  (45 x 45 at N = 16),
// one partial-pivot LU per rod, two solves (the couple's right-hand side needs N).  The 3 x 3 blocks K^ act on both
// sides of the quaternion algebra (v -> K x v is a commutator), so this system does not fit the one-sided quaternion
// elimination of the fused kernels; it is a plain real LU, one rod per warp, matrix in shared memory: the lanes own the
// columns of the trailing block (contiguous, conflict-free rows), the pivot search is a warp arg-max over the column.
// Unblocked right-looking elimination with the usual pivot choice (largest magnitude, first on ties), so that a
// sequential CPU restatement of the same steps agrees to a few ulp.
#pragma once
#include "sri_device.cuh"
#include "sri_fused16.cuh"  // FusedParams

namespace sri {

constexpr int kWrenchWarps = 4;
constexpr int kWrenchNmax = 45;              // 3 (N - 1), N <= 16
constexpr int kWrenchLD = 47;                // odd leading dimension: column accesses spread over the banks
struct WrenchScratch {                       // per warp, doubles
    static constexpr int A = 0;                              // [45][47]
    static constexpr int b = A + kWrenchNmax * kWrenchLD;    // [48] right-hand side / solution
    static constexpr int R = b + 48;                         // [16][9] rotation matrices by node (row-major)
    static constexpr int Nl = R + 144;                       // [48] local force (kept for the couple's right-hand side)
    static constexpr int piv = Nl + 48;                      // [48] ints in 24 doubles
    static constexpr int total = piv + 24;
};
constexpr size_t kWrenchSmem = (225 + 16 + (size_t)kWrenchWarps * WrenchScratch::total) * sizeof(double);

struct WrenchParams {
    long long batch;
    int N, M;
    const double* D_TT;  // [M][M] column-major (as sri_get_operator(4))
    const double* D_TI;  // [M]
    const double *K, *Q, *q0, *Gamma, *fbar, *lbar, *F_tip, *M_tip;
    double* Lambda;      // [batch][6][N]
    int* info;
};

__device__ __forceinline__ void quat_to_rot_rm(const quat& q, double* R) {  // Eigen toRotationMatrix, row-major
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

__global__ void __launch_bounds__(32 * kWrenchWarps) wrench_local_solve_kernel(const WrenchParams p) {
    extern __shared__ __align__(16) double wsm[];
    double* dtt = wsm;          // [15][15] column-major
    double* dti = wsm + 225;    // [15]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* scr = wsm + 241 + warp * WrenchScratch::total;
    double* A = scr + WrenchScratch::A;
    double* b = scr + WrenchScratch::b;
    double* Rm = scr + WrenchScratch::R;
    double* Nl = scr + WrenchScratch::Nl;
    int* piv = reinterpret_cast<int*>(scr + WrenchScratch::piv);
    const int N = p.N, M = p.M, n = 3 * M;
    constexpr int LD = kWrenchLD;
    for (int i = threadIdx.x; i < M * M; i += blockDim.x) dtt[i] = p.D_TT[i];
    for (int i = threadIdx.x; i < M; i += blockDim.x) dti[i] = p.D_TI[i];
    __syncthreads();

    // forward elimination + back substitution of b with the factors in A (row interchanges applied first)
    auto solve = [&]() {
        if (lane == 0)
            for (int k = 0; k < n; ++k) { const int pk = piv[k]; if (pk != k) { const double t = b[k]; b[k] = b[pk]; b[pk] = t; } }
        __syncwarp();
        for (int k = 0; k < n; ++k) {
            const double v = b[k];
            for (int i = k + 1 + lane; i < n; i += 32) b[i] -= A[i * LD + k] * v;
            __syncwarp();
        }
        for (int k = n - 1; k >= 0; --k) {
            if (lane == 0) b[k] /= A[k * LD + k];
            __syncwarp();
            const double v = b[k];
            for (int i = lane; i < k; i += 32) b[i] -= A[i * LD + k] * v;
            __syncwarp();
        }
    };

    const long long warps_total = (long long)gridDim.x * kWrenchWarps;
    for (long long rod = (long long)blockIdx.x * kWrenchWarps + warp; rod < p.batch; rod += warps_total) {
        // ---- rotations by node, operator -------------------------------------------------------------------------
        if (lane <= M) {
            quat q; q.w = 1.0; q.x = 0.0; q.y = 0.0; q.z = 0.0;
            if (lane < M) { const double* s = p.Q + rod * 4 * M + lane; q.w = s[0]; q.x = s[M]; q.y = s[2 * M]; q.z = s[3 * M]; }
            else if (p.q0) { const double* s = p.q0 + rod * 4; q.w = s[0]; q.x = s[1]; q.y = s[2]; q.z = s[3]; }
            quat_to_rot_rm(q, Rm + 9 * lane);
        }
        for (int e = lane; e < n * n; e += 32) {
            const int r = e / n, c = e - r * n;
            const int i = r / 3, a = r - 3 * i, j = c / 3, bb = c - 3 * j;
            double v = (a == bb) ? dtt[j * M + i] : 0.0;
            if (i == j && a != bb) {  // K^ of node i+1: [[0,-k2,k1],[k2,0,-k0],[-k1,k0,0]]
                const int comp = 3 - a - bb;
                const double kv = p.K[rod * 3 * N + comp * N + i + 1];
                const bool pos = (a == 0 && bb == 2) || (a == 1 && bb == 0) || (a == 2 && bb == 1);
                v += pos ? kv : -kv;
            }
            A[r * LD + c] = v;
        }
        __syncwarp();
        // ---- partial-pivot LU, in place ----------------------------------------------------------------------------
        int bad = 0;
        for (int k = 0; k < n; ++k) {
            double best = -1.0; int bi = k;
            for (int r = k + lane; r < n; r += 32) { const double v = fabs(A[r * LD + k]); if (v > best) { best = v; bi = r; } }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {  // arg-max; ties go to the smaller row index, as a sequential scan does
                const double ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) piv[k] = bi;
            if (best == 0.0) { if (!bad) bad = k + 1; __syncwarp(); continue; }
            if (bi != k)
                for (int c = lane; c < n; c += 32) { const double t = A[k * LD + c]; A[k * LD + c] = A[bi * LD + c]; A[bi * LD + c] = t; }
            __syncwarp();
            const double inv = 1.0 / A[k * LD + k];
            for (int r = k + 1 + lane; r < n; r += 32) A[r * LD + k] *= inv;
            __syncwarp();
            for (int c0 = k + 1; c0 < n; c0 += 32) {  // lanes own columns of the trailing block
                const int c = c0 + lane;
                if (c < n) {
                    const double u = A[k * LD + c];
                    int r = k + 1;
                    // four rows at a time, loads before stores: the rows are independent, but the compiler cannot prove
                    // that the store of one row does not alias the loads of the next (the plain loop ran 4 x slower)
                    for (; r + 3 < n; r += 4) {
                        double* a0 = A + r * LD;
                        const double l0 = a0[k], l1 = a0[LD + k], l2 = a0[2 * LD + k], l3 = a0[3 * LD + k];
                        const double v0 = a0[c], v1 = a0[LD + c], v2 = a0[2 * LD + c], v3 = a0[3 * LD + c];
                        a0[c] = v0 - l0 * u; a0[LD + c] = v1 - l1 * u; a0[2 * LD + c] = v2 - l2 * u; a0[3 * LD + c] = v3 - l3 * u;
                    }
                    for (; r < n; ++r) A[r * LD + c] -= A[r * LD + k] * u;
                }
            }
            __syncwarp();
        }
        // ---- internal force: b = -R_i^T fbar_i - D_TI N0 ---------------------------------------------------------------
        double N0[3], C0[3];
        {
            const double* F = p.F_tip + rod * 3; const double* T = p.M_tip + rod * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                N0[c] = Rm[0 * 3 + c] * F[0] + Rm[1 * 3 + c] * F[1] + Rm[2 * 3 + c] * F[2];
                C0[c] = Rm[0 * 3 + c] * T[0] + Rm[1 * 3 + c] * T[1] + Rm[2 * 3 + c] * T[2];
            }
        }
        for (int e = lane; e < n; e += 32) {
            const int i = e / 3, c = e - 3 * i;
            const double* Ri = Rm + 9 * (i + 1);
            double rf = 0.0;
            if (p.fbar) { const double* f = p.fbar + rod * 3 * N + i + 1; rf = Ri[0 * 3 + c] * f[0] + Ri[1 * 3 + c] * f[N] + Ri[2 * 3 + c] * f[2 * N]; }
            b[e] = -rf - dti[i] * N0[c];
        }
        __syncwarp();
        solve();
        for (int e = lane; e < n; e += 32) Nl[e] = b[e];
        __syncwarp();
        // ---- internal couple: b = -Gamma_i x N_i - R_i^T lbar_i - D_TI C0 -----------------------------------------------
        for (int e = lane; e < n; e += 32) {
            const int i = e / 3, c = e - 3 * i;
            const double* Ri = Rm + 9 * (i + 1);
            double g[3] = {1.0, 0.0, 0.0};
            if (p.Gamma) { const double* gm = p.Gamma + rod * 3 * N + i + 1; g[0] = gm[0]; g[1] = gm[N]; g[2] = gm[2 * N]; }
            const double n0 = Nl[3 * i], n1 = Nl[3 * i + 1], n2 = Nl[3 * i + 2];
            const double gx = (c == 0) ? g[1] * n2 - g[2] * n1 : (c == 1 ? g[2] * n0 - g[0] * n2 : g[0] * n1 - g[1] * n0);
            double rl = 0.0;
            if (p.lbar) { const double* l = p.lbar + rod * 3 * N + i + 1; rl = Ri[0 * 3 + c] * l[0] + Ri[1 * 3 + c] * l[N] + Ri[2 * 3 + c] * l[2 * N]; }
            b[e] = -gx - rl - dti[i] * C0[c];
        }
        __syncwarp();
        solve();
        // ---- Lambda [6][N]: couple first ------------------------------------------------------------------------------
        double* out = p.Lambda + rod * 6 * N;
        if (lane < 3) { out[lane * N] = C0[lane]; out[(3 + lane) * N] = N0[lane]; }
        for (int e = lane; e < n; e += 32) {
            const int i = e / 3, c = e - 3 * i;
            out[c * N + i + 1] = b[e];
            out[(3 + c) * N + i + 1] = Nl[e];
        }
        if (p.info && lane == 0) p.info[rod] = bad;
        __syncwarp();
    }
}

}  // namespace sri
