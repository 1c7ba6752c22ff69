// sri_wrench_gj_static.cuh -- local-frame statics solved directly, first pass for N = 17 and 23 <= N <= 33:
// the stage-1 recipe applied to the real 3M x 3M operator.
//
//   (D_TT (x) I3 + blockdiag(K^_j)) x = b      is left-multiplied by the cached S = D_TT^-1 (the integration matrix):
//   (I + (S (x) I3) blockdiag(K^_j)) x = (S (x) I3) b,      block (i, j) of the operator = delta_ij I3 + S_ij K^_j  (closed form).
// K^_j is skew, so like the quaternion operator of stage 1 this matrix is benign in STATIC pivot order on the strain range
// the reference targets (largest multiplier 0.3 on the benchmark rods, 2 at |K| = 20), which removes the arg-max, the vote
// and the pivot bookkeeping from every elimination step of sri_wrench_gj_multi.cuh -- what is left per step is: the owner
// (row k) publishes its window, one CTA barrier, every row updates.  One operator row per lane, one rod per CTA of NW warps,
// the rolled Gauss-Jordan on a sliding register window, multipliers [step][row] in shared memory for the second right-hand
// side (both right-hand sides are preconditioned by one M-term dot product per row).  Every multiplier is checked against a
// growth bound (integer compare of the high words; a zero or non-finite pivot trips it too); a rod that fails is appended to
// the hand-back list and re-solved with partial pivoting by the second pass (sri_wrench_gj.cuh / sri_wrench_gj_multi.cuh
// over that list) inside the same call.
// Measured against the row-pivoting kernels alone (run AA): +16 % at N = 17, +22..27 % at 23 <= N <= 33; no gain or a loss for
// N <= 16 and 18 <= N <= 22 (one row per lane is bound by the shared-memory return path that broadcasts the pivot row, 512 B
// per LDS.128, not by the arg-max this kernel removes: ncu digests under profiles/), so those N keep one pass.
#pragma once
#include "sri_wrench_gj_multi.cuh"

namespace sri {

template <int NW, int WMAX_>
struct WrenchGjStaticCfg {
    static_assert(WMAX_ % 8 == 0 && WMAX_ <= 32 * NW, "window: a multiple of 8 slots, at most one per row");
    static constexpr int ROWS = 32 * NW;
    static constexpr int WMAX = WMAX_;
    static constexpr int MC = WMAX / 3;                  // interior-node capacity (M <= MC)
    static constexpr int NODES = MC + 1;
    static constexpr int L = 0;                          // [WMAX][ROWS] multipliers by (step, row)
    static constexpr int urow = L + WMAX * ROWS;         // [2][WMAX + 4] published pivot row: window, (RHS, 1/pivot) at WMAX
    static constexpr int S = urow + 2 * (WMAX + 4);      // [MC][MC] D_TT^-1, column-major with leading dimension MC
    static constexpr int R = S + MC * MC + ((MC * MC) & 1);  // [NODES][9]
    static constexpr int kk = R + 9 * NODES + (NODES & 1);   // [3][NODES]
    static constexpr int sdti = kk + 3 * NODES + (NODES & 1);  // [MC] S D_TI
    static constexpr int g = sdti + MC + (MC & 1);       // [WMAX] un-preconditioned right-hand side by unknown index
    static constexpr int vec = g + WMAX;                 // [WMAX] couple solution
    static constexpr int Nl = vec + WMAX;                // [WMAX] force solution
    static constexpr int ysm = Nl + WMAX;                // [2]
    static constexpr int total = ysm + 2;
    static constexpr size_t smem_bytes = (size_t)total * sizeof(double);
};

template <int NW, int WMAX, int W>
__device__ __forceinline__ void wrench_gjs_body(double (&A)[WMAX], double& rhs, int& k, const int n, const int row, double& rc,
                                                int& gacc, double* __restrict__ sm) {
    using C = WrenchGjStaticCfg<NW, WMAX>;
#pragma unroll 1
    for (int s = 0; s < 8 && k < n; ++s, ++k) {
        const bool mine = row == k;
        double* ub = sm + C::urow + (k & 1) * (WMAX + 4);
        if (mine) {
#pragma unroll
            for (int j = 0; j < W; j += 2) *reinterpret_cast<double2*>(ub + j) = make_double2(A[j], A[j + 1]);
            *reinterpret_cast<double2*>(ub + WMAX) = make_double2(rhs, rc);
        }
        __syncthreads();
        const double2 tail = *reinterpret_cast<const double2*>(ub + WMAX);   // (right-hand side of the pivot row, 1 / pivot)
        const double ml = mine ? 1.0 - tail.y : A[0] * tail.y;   // the pivot row is normalised by the same update
        sm[C::L + k * C::ROWS + row] = ml;
        gacc = max(gacc, mine ? 0 : (__double2hiint(ml) & 0x7fffffff));
        {
            const double2 u = *reinterpret_cast<const double2*>(ub);
            A[0] = fma(-ml, u.y, A[1]);
        }
        rc = fast_rcp(A[0]);   // (only row k + 1 will use it: ready long before its turn)
#pragma unroll
        for (int j = 2; j < W; j += 2) {
            const double2 u = *reinterpret_cast<const double2*>(ub + j);
            A[j - 1] = fma(-ml, u.x, A[j]);
            A[j] = fma(-ml, u.y, A[j + 1]);
        }
        A[W - 1] = 0.0;
        rhs = fma(-ml, tail.x, rhs);
    }
}

template <int NW, int WMAX, int W>
__device__ __forceinline__ void wrench_gjs_bodies(double (&A)[WMAX], double& rhs, int& k, const int n, const int row, double& rc,
                                                  int& gacc, double* __restrict__ sm) {
    wrench_gjs_body<NW, WMAX, W>(A, rhs, k, n, row, rc, gacc, sm);
    if constexpr (W > 8) wrench_gjs_bodies<NW, WMAX, W - 8>(A, rhs, k, n, row, rc, gacc, sm);
}

// preconditioned right-hand side of row (ri, c): -(sum_j S[ri][j] g[3 j + c]) - (S D_TI)[ri] x0[c]
template <class C>
__device__ __forceinline__ double wrench_gjs_rhs(const double* __restrict__ sm, int M, int ri, int c, double x0c) {
    double acc = 0.0;
#pragma unroll 4
    for (int j = 0; j < M; ++j) acc = fma(sm[C::S + j * C::MC + ri], sm[C::g + 3 * j + c], acc);
    return -acc - sm[C::sdti + ri] * x0c;
}

template <int NW, int WMAX, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB) wrench_local_solve_gj_static_kernel(const WrenchParams p) {
    using C = WrenchGjStaticCfg<NW, WMAX>;
    extern __shared__ __align__(16) double sm[];
    double* Rm = sm + C::R;
    double* kk = sm + C::kk;
    double* gs = sm + C::g;
    double* vec = sm + C::vec;
    double* Nl = sm + C::Nl;
    double* ysm = sm + C::ysm;
    const int tid = threadIdx.x;
    const int N = p.N, M = p.M, n = 3 * M;
    const int row = tid;
    const int ri = row / 3, c = row - 3 * ri;
    const bool real = row < n;
    // ---- once per CTA: S = D_TT^-1 and S D_TI -------------------------------------------------------------------------------
    for (int e = tid; e < C::MC * C::MC; e += 32 * NW) {
        const int j = e / C::MC, i = e - j * C::MC;
        sm[C::S + e] = (i < M && j < M) ? p.S[(size_t)j * M + i] : 0.0;
    }
    __syncthreads();
    for (int i = tid; i < C::MC; i += 32 * NW) {
        double acc = 0.0;
        for (int j = 0; j < M; ++j) acc = fma(sm[C::S + j * C::MC + i], p.D_TI[j], acc);
        sm[C::sdti + i] = i < M ? acc : 0.0;
    }

    for (long long rod = blockIdx.x; rod < p.batch; rod += gridDim.x) {
        __syncthreads();
        for (int e = tid; e < 3 * C::NODES; e += 32 * NW) { const int cc = e / C::NODES, i = e - cc * C::NODES; kk[e] = i < N ? p.K[rod * 3 * N + cc * N + i] : 0.0; }
        if (tid <= M) {
            quat q; q.w = 1.0; q.x = 0.0; q.y = 0.0; q.z = 0.0;
            if (tid < M) { const double* s = p.Q + rod * 4 * M + tid; q.w = s[0]; q.x = s[M]; q.y = s[2 * M]; q.z = s[3 * M]; }
            else if (p.q0) { const double* s = p.q0 + rod * 4; q.w = s[0]; q.x = s[1]; q.y = s[2]; q.z = s[3]; }
            quat_to_rot_rm(q, Rm + 9 * tid);
        }
        __syncthreads();
        double N0[3], C0[3];
        {
            const double* F = p.F_tip + rod * 3; const double* T = p.M_tip + rod * 3;
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                N0[cc] = Rm[0 * 3 + cc] * F[0] + Rm[1 * 3 + cc] * F[1] + Rm[2 * 3 + cc] * F[2];
                C0[cc] = Rm[0 * 3 + cc] * T[0] + Rm[1 * 3 + cc] * T[1] + Rm[2 * 3 + cc] * T[2];
            }
        }
        const double* Ri = Rm + 9 * (ri + 1);
        if (row < WMAX) {
            double rf = 0.0;
            if (real && p.fbar) { const double* f = p.fbar + rod * 3 * N + ri + 1; rf = Ri[0 * 3 + c] * f[0] + Ri[1 * 3 + c] * f[N] + Ri[2 * 3 + c] * f[2 * N]; }
            gs[row] = rf;
        }
        // ---- this lane's operator row: delta + S[ri][j] K^_j[c][.] (identity on the padding rows) -----------------------------
        double A[WMAX];
#pragma unroll
        for (int jn = 0; jn < C::MC; ++jn) {
            const double s = real ? sm[C::S + jn * C::MC + ri] : 0.0;   // (zero for jn >= M)
            const double k0 = kk[jn + 1], k1 = kk[C::NODES + jn + 1], k2 = kk[2 * C::NODES + jn + 1];
            // K^ row c: [[0,-k2,k1],[k2,0,-k0],[-k1,k0,0]]
            A[3 * jn + 0] = s * (c == 0 ? 0.0 : (c == 1 ? k2 : -k1)) + (row == 3 * jn + 0 ? 1.0 : 0.0);
            A[3 * jn + 1] = s * (c == 0 ? -k2 : (c == 1 ? 0.0 : k0)) + (row == 3 * jn + 1 ? 1.0 : 0.0);
            A[3 * jn + 2] = s * (c == 0 ? k1 : (c == 1 ? -k0 : 0.0)) + (row == 3 * jn + 2 ? 1.0 : 0.0);
        }
#pragma unroll
        for (int j = 3 * C::MC; j < WMAX; ++j) A[j] = (j == row) ? 1.0 : 0.0;
        __syncthreads();
        double rhs = real ? wrench_gjs_rhs<C>(sm, M, ri, c, c == 0 ? N0[0] : (c == 1 ? N0[1] : N0[2])) : 0.0;
        // ---- Gauss-Jordan in static order -----------------------------------------------------------------------------------------
        int gacc = 0, k = 0;
        double rc = fast_rcp(A[0]);
        wrench_gjs_bodies<NW, WMAX, WMAX>(A, rhs, k, n, row, rc, gacc, sm);
        if (__syncthreads_or(gacc > p.growth_hi)) {   // hand the rod back to the row-pivoting pass
            if (tid == 0) p.rod_list[atomicAdd(p.rod_count, 1)] = (int)rod;
            continue;
        }
        if (real) Nl[row] = rhs;
        __syncthreads();
        // ---- internal couple through the stored multipliers ---------------------------------------------------------------------
        if (real) {
            double g0 = 1.0, g1 = 0.0, g2 = 0.0;
            if (p.Gamma) { const double* gm = p.Gamma + rod * 3 * N + ri + 1; g0 = gm[0]; g1 = gm[N]; g2 = gm[2 * N]; }
            const double n0 = Nl[3 * ri], n1 = Nl[3 * ri + 1], n2 = Nl[3 * ri + 2];
            const double gx = (c == 0) ? g1 * n2 - g2 * n1 : (c == 1 ? g2 * n0 - g0 * n2 : g0 * n1 - g1 * n0);
            double rl = 0.0;
            if (p.lbar) { const double* l = p.lbar + rod * 3 * N + ri + 1; rl = Ri[0 * 3 + c] * l[0] + Ri[1 * 3 + c] * l[N] + Ri[2 * 3 + c] * l[2 * N]; }
            gs[row] = gx + rl;
        }
        __syncthreads();
        double b = real ? wrench_gjs_rhs<C>(sm, M, ri, c, c == 0 ? C0[0] : (c == 1 ? C0[1] : C0[2])) : 0.0;
        if constexpr (NW == 1) {
#pragma unroll 1
            for (int kq = 0; kq < n; ++kq) b = fma(-sm[C::L + kq * C::ROWS + row], __shfl_sync(0xffffffffu, b, kq), b);
        } else {
#pragma unroll 1
            for (int kq = 0; kq < n; ++kq) {
                if (row == kq) ysm[kq & 1] = b;   // the pivot row of step kq posts its current value
                __syncthreads();
                b = fma(-sm[C::L + kq * C::ROWS + row], ysm[kq & 1], b);
            }
        }
        if (real) vec[row] = b;
        __syncthreads();
        // ---- Lambda [6][N]: couple first ---------------------------------------------------------------------------------------
        double* out = p.Lambda + rod * 6 * N;
        if (tid < 3) {
            out[tid * N] = tid == 0 ? C0[0] : (tid == 1 ? C0[1] : C0[2]);
            out[(3 + tid) * N] = tid == 0 ? N0[0] : (tid == 1 ? N0[1] : N0[2]);
        }
        for (int e = tid; e < n; e += 32 * NW) {
            const int i = e / 3, cc = e - 3 * i;
            out[cc * N + i + 1] = vec[e];
            out[(3 + cc) * N + i + 1] = Nl[e];
        }
        if (p.info && tid == 0) p.info[rod] = 0;
    }
}

}  // namespace sri
