// sri_wrench_generic.cuh -- local-frame statics solved directly for 17 <= N <= 64 (rod_modeling.pdf eqs. 1.29,
// 2.18): the same problem as sri_wrench_solve.cuh (which see), one rod per CTA.
//
//   A = D_TT (x) I3 + blockdiag(K^_i),  n = 3 (N - 1) <= 189,  one partial-pivot LU per rod, two solves.
// The operator is a plain dense real matrix (the two-sided K^ blocks do not fit the quaternion elimination), 69 KB at
// N = 32 and 286 KB at N = 64: it lives in shared memory while it fits (N <= 55) and in a per-CTA scratch in global memory
// beyond that (a few hundred CTAs are resident, so the scratch stays in the 126 MB L2).  Right-looking LU, one column per step:
// CTA-wide arg-max (largest magnitude, smallest row on ties: the pivots of a sequential partial-pivot LU),
// whole-row exchange, rank-1 update with a warp per row; the triangular solves are column sweeps with one barrier per
// unknown.  Simple on purpose: this entry point is off the benchmark path, the N <= 16 kernel is the tuned one.
#pragma once
#include "sri_wrench_solve.cuh"

namespace sri {

constexpr int kWrenchGenThreads = 256;

struct WrenchGenLayout {  // shared memory, doubles (the matrix first when it is held in shared memory)
    int n, ld, N;
    bool in_smem;
    __host__ __device__ int A() const { return 0; }
    __host__ __device__ int y() const { return in_smem ? n * ld : 0; }          // [n] right-hand side / intermediate
    __host__ __device__ int x() const { return y() + n; }                        // [n] solution
    __host__ __device__ int Nl() const { return x() + n; }                       // [n] local force
    __host__ __device__ int R() const { return Nl() + n; }                       // [N][9]
    __host__ __device__ int kk() const { return R() + 9 * N; }                   // [3][N]
    __host__ __device__ int dinv() const { return kk() + 3 * N; }                // [n] reciprocal pivots
    __host__ __device__ int dti() const { return dinv() + n; }                   // [N]
    __host__ __device__ int red() const { return dti() + N; }                    // [16] arg-max partials (value, row) per warp
    __host__ __device__ int perm() const { return red() + 16; }                  // n ints
    __host__ __device__ int total() const { return perm() + (n + 1) / 2 + 2; }
};

__global__ void __launch_bounds__(kWrenchGenThreads) wrench_local_solve_generic_kernel(const WrenchParams p, double* __restrict__ gscratch,
                                                                                       const int in_smem) {
    extern __shared__ __align__(16) double wsm[];
    const int N = p.N, M = p.M, n = 3 * M, ld = n | 1;
    const WrenchGenLayout L{n, ld, N, in_smem != 0};
    double* A = in_smem ? wsm + L.A() : gscratch + (size_t)blockIdx.x * n * ld;
    double* y = wsm + L.y();
    double* x = wsm + L.x();
    double* Nl = wsm + L.Nl();
    double* Rm = wsm + L.R();
    double* kk = wsm + L.kk();
    double* dinv = wsm + L.dinv();
    double* dti = wsm + L.dti();
    double* red = wsm + L.red();
    int* perm = reinterpret_cast<int*>(wsm + L.perm());
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int T = kWrenchGenThreads, WARPS = T / 32;
    constexpr unsigned FULL = 0xffffffffu;
    for (int i = tid; i < M; i += T) dti[i] = p.D_TI[i];

    // y (original row order, in `y`) -> solution in `x`; uses the factors in A, dinv and perm
    auto solve = [&]() {
        __syncthreads();
        for (int i = tid; i < n; i += T) x[i] = y[perm[i]];
        __syncthreads();
        for (int i = tid; i < n; i += T) y[i] = x[i];
        __syncthreads();
        for (int k = 0; k < n - 1; ++k) {           // L y' = P y, unit lower triangle
            const double yk = y[k];
            for (int i = k + 1 + tid; i < n; i += T) y[i] = fma(-A[i * ld + k], yk, y[i]);
            __syncthreads();
        }
        for (int k = n - 1; k >= 0; --k) {          // U x = y'
            const double xk = y[k] * dinv[k];
            if (tid == 0) x[k] = xk;
            for (int i = tid; i < k; i += T) y[i] = fma(-A[i * ld + k], xk, y[i]);
            __syncthreads();
        }
    };

    for (long long rod = blockIdx.x; rod < p.batch; rod += gridDim.x) {
        __syncthreads();
        for (int e = tid; e < 3 * N; e += T) kk[e] = p.K[rod * 3 * N + e];
        if (tid <= M) {
            quat q; q.w = 1.0; q.x = 0.0; q.y = 0.0; q.z = 0.0;
            if (tid < M) { const double* s = p.Q + rod * 4 * M + tid; q.w = s[0]; q.x = s[M]; q.y = s[2 * M]; q.z = s[3 * M]; }
            else if (p.q0) { const double* s = p.q0 + rod * 4; q.w = s[0]; q.x = s[1]; q.y = s[2]; q.z = s[3]; }
            quat_to_rot_rm(q, Rm + 9 * tid);
        }
        for (int i = tid; i < n; i += T) perm[i] = i;
        __syncthreads();
        // ---- operator ---------------------------------------------------------------------------------------------------
        for (int e = tid; e < n * n; e += T) {
            const int r = e / n, c = e - r * n;
            const int i = r / 3, a = r - 3 * i, j = c / 3, b = c - 3 * j;
            double v = (a == b) ? __ldg(p.D_TT + (size_t)j * M + i) : 0.0;
            if (i == j && a != b) {   // K^ of node i+1: [[0,-k2,k1],[k2,0,-k0],[-k1,k0,0]]
                const double kv = kk[(3 - a - b) * N + i + 1];
                v += (b == a + 2 || a == b + 1) ? kv : -kv;
            }
            A[r * ld + c] = v;
        }
        __syncthreads();
        // ---- LU with partial pivoting, one column per step --------------------------------------------------------------------
        int bad = 0;
        for (int k = 0; k < n; ++k) {
            // arg-max of |A[k.., k]|: exact, smallest row on ties
            double best = -1.0;
            int brow = n;
            for (int i = k + tid; i < n; i += T) {
                const double v = fabs(A[i * ld + k]);
                if (v > best) { best = v; brow = i; }   // (a NaN never wins: the column then looks smaller than it is)
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const double ov = __shfl_xor_sync(FULL, best, off);
                const int orow = __shfl_xor_sync(FULL, brow, off);
                if (ov > best || (ov == best && orow < brow)) { best = ov; brow = orow; }
            }
            if (lane == 0) { red[2 * warp] = best; red[2 * warp + 1] = (double)brow; }
            __syncthreads();
            best = red[0]; brow = (int)red[1];
#pragma unroll
            for (int w2 = 1; w2 < WARPS; ++w2) {
                const double ov = red[2 * w2];
                const int orow = (int)red[2 * w2 + 1];
                if (ov > best || (ov == best && orow < brow)) { best = ov; brow = orow; }
            }
            const bool singular = !(best > 0.0) || !(best < INFINITY);
            if (singular && !bad) bad = k + 1;
            const int prow = (brow < n) ? brow : k;
            if (prow != k) {
                for (int j = tid; j < n; j += T) { const double t = A[k * ld + j]; A[k * ld + j] = A[prow * ld + j]; A[prow * ld + j] = t; }
                if (tid == 0) { const int t = perm[k]; perm[k] = perm[prow]; perm[prow] = t; }
            }
            __syncthreads();
            const double piv = A[k * ld + k];
            const double rp = 1.0 / piv;
            const double inv = singular ? 0.0 : rp;
            if (tid == 0) dinv[k] = rp;
            // rank-1 update, a warp per row: l = A[i][k] / pivot stays in place of the eliminated entry
            for (int i = k + 1 + warp; i < n; i += WARPS) {
                double* row = A + i * ld;
                const double l = row[k] * inv;
                __syncwarp();
                if (lane == 0) row[k] = l;
                const double* prw = A + k * ld;
                for (int j = k + 1 + lane; j < n; j += 32) row[j] = fma(-l, prw[j], row[j]);
            }
            __syncthreads();
        }
        // ---- internal force: y = -R_i^T fbar_i - D_TI N0 ----------------------------------------------------------------------
        double N0[3], C0[3];
        {
            const double* F = p.F_tip + rod * 3; const double* Tq = p.M_tip + rod * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                N0[c] = Rm[0 * 3 + c] * F[0] + Rm[1 * 3 + c] * F[1] + Rm[2 * 3 + c] * F[2];
                C0[c] = Rm[0 * 3 + c] * Tq[0] + Rm[1 * 3 + c] * Tq[1] + Rm[2 * 3 + c] * Tq[2];
            }
        }
        for (int e = tid; e < n; e += T) {
            const int i = e / 3, c = e - 3 * i;
            const double* Ri = Rm + 9 * (i + 1);
            double rf = 0.0;
            if (p.fbar) { const double* f = p.fbar + rod * 3 * N + i + 1; rf = Ri[0 * 3 + c] * f[0] + Ri[1 * 3 + c] * f[N] + Ri[2 * 3 + c] * f[2 * N]; }
            y[e] = -rf - dti[i] * (c == 0 ? N0[0] : (c == 1 ? N0[1] : N0[2]));
        }
        solve();
        for (int e = tid; e < n; e += T) Nl[e] = x[e];
        __syncthreads();
        // ---- internal couple: y = -Gamma_i x N_i - R_i^T lbar_i - D_TI C0 -------------------------------------------------------
        for (int e = tid; e < n; e += T) {
            const int i = e / 3, c = e - 3 * i;
            const double* Ri = Rm + 9 * (i + 1);
            double g[3] = {1.0, 0.0, 0.0};
            if (p.Gamma) { const double* gm = p.Gamma + rod * 3 * N + i + 1; g[0] = gm[0]; g[1] = gm[N]; g[2] = gm[2 * N]; }
            const double n0 = Nl[3 * i], n1 = Nl[3 * i + 1], n2 = Nl[3 * i + 2];
            const double gx = (c == 0) ? g[1] * n2 - g[2] * n1 : (c == 1 ? g[2] * n0 - g[0] * n2 : g[0] * n1 - g[1] * n0);
            double rl = 0.0;
            if (p.lbar) { const double* l = p.lbar + rod * 3 * N + i + 1; rl = Ri[0 * 3 + c] * l[0] + Ri[1 * 3 + c] * l[N] + Ri[2 * 3 + c] * l[2 * N]; }
            y[e] = -gx - rl - dti[i] * (c == 0 ? C0[0] : (c == 1 ? C0[1] : C0[2]));
        }
        solve();
        // ---- Lambda [6][N]: couple first ------------------------------------------------------------------------------------
        double* out = p.Lambda + rod * 6 * N;
        if (tid < 3) {
            out[tid * N] = tid == 0 ? C0[0] : (tid == 1 ? C0[1] : C0[2]);
            out[(3 + tid) * N] = tid == 0 ? N0[0] : (tid == 1 ? N0[1] : N0[2]);
        }
        for (int e = tid; e < n; e += T) {
            const int i = e / 3, c = e - 3 * i;
            out[c * N + i + 1] = x[e];
            out[(3 + c) * N + i + 1] = Nl[e];
        }
        if (p.info && tid == 0) p.info[rod] = bad;
    }
}

}  // namespace sri
