// sri_wrench_gj_multi.cuh -- local-frame statics solved directly for 18 <= N <= 33: the register-resident rolled
// Gauss-Jordan elimination of sri_wrench_gj.cuh (which see for the algorithm) spread over NW = 2 or 3 warps, one rod per CTA.
//
//   n = 3 (N - 1) <= 32 NW rows, ONE row per lane (a sliding window of 32 NW registers, slot 0 = current pivot column),
//   implicit partial pivoting, multipliers of all rows and steps kept in shared memory ([step][row], 32 KB / 72 KB per rod)
//   for the second right-hand side.
// What the single-warp kernel does with warp votes and __syncwarp is done here with one small exchange per step: every warp
// reduces its own candidates (redux + ballot), lane 0 of each warp posts (|a| as two words, row), a CTA barrier, and every
// thread picks the best of the NW posts (exact, first row on ties: the pivots of a sequential partial-pivot elimination);
// the owner then publishes its window, its right-hand side and the reciprocal of the pivot, a second barrier, and all rows
// update.  The arg-max posts of step k + 1 are written before the bulk update of step k, so only the barriers themselves
// sit on the critical path.  Replaces the CTA-wide shared-memory LU of sri_wrench_generic.cuh for these N (that kernel
// keeps 34 <= N <= 64, where a row no longer fits a lane's registers).
#pragma once
#include "sri_wrench_gj.cuh"

namespace sri {

template <int NW, int WMAX_ = 32 * NW>
struct WrenchGjMultiCfg {
    static_assert(WMAX_ % 8 == 0 && WMAX_ <= 32 * NW, "window: a multiple of 8 slots, at most one per row");
    static constexpr int ROWS = 32 * NW;                 // rows (one per lane)
    static constexpr int WMAX = WMAX_;                   // window slots = unknowns the instantiation can hold
    static constexpr int NODES = WMAX / 3 + 1;           // node capacity (N <= NODES): 17 / 22 / 33
    // shared memory, doubles
    static constexpr int L = 0;                          // [WMAX][ROWS] multipliers by (step, row)
    static constexpr int urow = L + WMAX * ROWS;         // [2][WMAX + 4] published pivot row: window, (RHS, 1/pivot) at WMAX
    static constexpr int cand = urow + 2 * (WMAX + 4);   // [2][NW] arg-max posts: int4 (hi word, lo word, row, -)
    static constexpr int R = cand + 2 * NW * 2;          // [NODES][9]
    static constexpr int kk = R + 9 * NODES + (NODES & 1);  // [3][NODES]
    static constexpr int dti = kk + 3 * NODES + (NODES & 1);  // [NODES]
    static constexpr int vec = dti + NODES + (NODES & 1);  // [WMAX] couple solution by unknown index
    static constexpr int Nl = vec + WMAX;                // [WMAX] force solution by unknown index
    static constexpr int ysm = Nl + WMAX;                // [2] pivot value of the second sweep (double buffered)
    static constexpr int total = ysm + 2;
    static constexpr size_t smem_bytes = (size_t)total * sizeof(double);
};

// this warp's best candidate (|a| as two words, row; row < 0: none left) -> the post slot of the warp (lane 0 writes; a
// single-warp CTA needs no exchange and keeps it in registers)
template <int NW>
__device__ __forceinline__ void wrench_gjm_post(double a, bool used, int row, int lane, int* __restrict__ slot, unsigned& mh,
                                                unsigned& ml, int& best) {
    constexpr unsigned FULL = 0xffffffffu;
    const double v = used ? 0.0 : fabs(a);
    const unsigned h = (unsigned)__double2hiint(v), l = (unsigned)__double2loint(v);
    mh = __reduce_max_sync(FULL, h);
    ml = __reduce_max_sync(FULL, h == mh ? l : 0u);
    const unsigned m = __ballot_sync(FULL, !used && h == mh && l == ml);
    best = m ? (row & ~31) + __ffs(m) - 1 : -1;
    if (NW > 1 && lane == 0) *reinterpret_cast<int4*>(slot) = make_int4((int)mh, (int)ml, best, 0);
}

// best of the NW posts: exact comparison of (hi, lo), first (lowest) row on ties
template <int NW>
__device__ __forceinline__ void wrench_gjm_decide(const int* __restrict__ posts, int& prow, bool& singular) {
    unsigned bh = 0u, bl = 0u;
    prow = -1;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const int4 post = *reinterpret_cast<const int4*>(posts + 4 * w);
        const unsigned h = (unsigned)post.x, l = (unsigned)post.y;
        const int r = post.z;
        if (r >= 0 && (prow < 0 || h > bh || (h == bh && l > bl))) { bh = h; bl = l; prow = r; }
    }
    singular = prow < 0 || (bh | bl) == 0u || bh >= 0x7ff00000u;
}

template <int NW, int WMAX, int W>
__device__ __forceinline__ void wrench_gjm_body(double (&A)[WMAX], double& rhs, int& k, const int n, const int row, bool& used,
                                                int& srow, int& bad, int& prow, bool& singular, double& rc,
                                                double* __restrict__ sm, const int lane, const int warp) {
    using C = WrenchGjMultiCfg<NW, WMAX>;
#pragma unroll 1
    for (int s = 0; s < 8 && k < n; ++s, ++k) {
        if (singular && !bad) bad = k + 1;
        const bool mine = prow == row;
        double* ub = sm + C::urow + (k & 1) * (WMAX + 4);
        if (mine) {
#pragma unroll
            for (int j = 0; j < W; j += 2) *reinterpret_cast<double2*>(ub + j) = make_double2(A[j], A[j + 1]);
            *reinterpret_cast<double2*>(ub + WMAX) = make_double2(rhs, rc);
            used = true; srow = k;
        }
        if (prow < 0 && lane == 0 && warp == 0) *reinterpret_cast<double2*>(ub + WMAX) = make_double2(0.0, 0.0);
        __syncthreads();
        const double2 tail = *reinterpret_cast<const double2*>(ub + WMAX);   // (right-hand side of the pivot row, 1 / pivot)
        const double inv = singular ? 0.0 : tail.y;
        const double ml = mine ? 1.0 - inv : A[0] * inv;   // the pivot row is normalised by the same update
        sm[C::L + k * C::ROWS + row] = ml;
        // the next pivot column first: its candidates are posted before the bulk of the update
        {
            const double2 u = *reinterpret_cast<const double2*>(ub);
            A[0] = fma(-ml, u.y, A[1]);
        }
        unsigned ph = 0u, pl = 0u;
        int pbest = -1;
        if (k + 1 < n) {
            rc = fast_rcp(A[0]);
            wrench_gjm_post<NW>(A[0], used, row, lane, reinterpret_cast<int*>(sm + C::cand) + (((k + 1) & 1) * NW + warp) * 4, ph, pl, pbest);
        }
#pragma unroll
        for (int j = 2; j < W; j += 2) {
            const double2 u = *reinterpret_cast<const double2*>(ub + j);
            A[j - 1] = fma(-ml, u.x, A[j]);
            A[j] = fma(-ml, u.y, A[j + 1]);
        }
        A[W - 1] = 0.0;
        rhs = fma(-ml, tail.x, rhs);
        if constexpr (NW == 1) {
            if (k + 1 < n) { prow = pbest; singular = pbest < 0 || (ph | pl) == 0u || ph >= 0x7ff00000u; }
        } else {
            __syncthreads();
            if (k + 1 < n) wrench_gjm_decide<NW>(reinterpret_cast<const int*>(sm + C::cand) + ((k + 1) & 1) * NW * 4, prow, singular);
        }
    }
}

template <int NW, int WMAX, int W>
__device__ __forceinline__ void wrench_gjm_bodies(double (&A)[WMAX], double& rhs, int& k, const int n, const int row, bool& used,
                                                  int& srow, int& bad, int& prow, bool& singular, double& rc,
                                                  double* __restrict__ sm, const int lane, const int warp) {
    wrench_gjm_body<NW, WMAX, W>(A, rhs, k, n, row, used, srow, bad, prow, singular, rc, sm, lane, warp);
    if constexpr (W > 8) wrench_gjm_bodies<NW, WMAX, W - 8>(A, rhs, k, n, row, used, srow, bad, prow, singular, rc, sm, lane, warp);
}

template <int NW, int WMAX, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB) wrench_local_solve_gj_multi_kernel(const WrenchParams p) {
    using C = WrenchGjMultiCfg<NW, WMAX>;
    extern __shared__ __align__(16) double sm[];
    double* Rm = sm + C::R;
    double* kk = sm + C::kk;
    double* dti = sm + C::dti;
    double* vec = sm + C::vec;
    double* Nl = sm + C::Nl;
    double* ysm = sm + C::ysm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, M = p.M, n = 3 * M;
    const int row = tid;                       // row = (node row / 3 + 1, component row % 3)
    const int ri = row / 3, rc_comp = row - 3 * ri;
    const bool real = row < n;
    for (int i = tid; i < C::NODES; i += 32 * NW) dti[i] = i < M ? p.D_TI[i] : 0.0;

    const long long count = p.from_list ? (long long)*p.rod_count : p.batch;
    for (long long it = blockIdx.x; it < count; it += gridDim.x) {
        const long long rod = p.from_list ? (long long)p.rod_list[it] : it;
        __syncthreads();
        for (int e = tid; e < 3 * C::NODES; e += 32 * NW) { const int c = e / C::NODES, i = e - c * C::NODES; kk[e] = i < N ? p.K[rod * 3 * N + c * N + i] : 0.0; }
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
            for (int c = 0; c < 3; ++c) {
                N0[c] = Rm[0 * 3 + c] * F[0] + Rm[1 * 3 + c] * F[1] + Rm[2 * 3 + c] * F[2];
                C0[c] = Rm[0 * 3 + c] * T[0] + Rm[1 * 3 + c] * T[1] + Rm[2 * 3 + c] * T[2];
            }
        }
        // ---- this lane's operator row (identity on the padding rows >= n) and the first right-hand side -------------------------
        double A[WMAX], rhs = 0.0;
        {
            // K^ of node ri+1, row rc_comp: [[0,-k2,k1],[k2,0,-k0],[-k1,k0,0]]
            const double k0 = real ? kk[ri + 1] : 0.0, k1 = real ? kk[C::NODES + ri + 1] : 0.0, k2 = real ? kk[2 * C::NODES + ri + 1] : 0.0;
            const double kh0 = rc_comp == 0 ? 0.0 : (rc_comp == 1 ? k2 : -k1);
            const double kh1 = rc_comp == 0 ? -k2 : (rc_comp == 1 ? 0.0 : k0);
            const double kh2 = rc_comp == 0 ? k1 : (rc_comp == 1 ? -k0 : 0.0);
#pragma unroll
            for (int jn = 0; jn < WMAX / 3; ++jn) {
                const double d = (real && jn < M) ? __ldg(p.D_TT + (size_t)jn * M + ri) : 0.0;
                const bool diag = real && jn == ri;
                A[3 * jn + 0] = (rc_comp == 0 ? d : 0.0) + (diag ? kh0 : 0.0);
                A[3 * jn + 1] = (rc_comp == 1 ? d : 0.0) + (diag ? kh1 : 0.0);
                A[3 * jn + 2] = (rc_comp == 2 ? d : 0.0) + (diag ? kh2 : 0.0);
            }
#pragma unroll
            for (int j = 3 * (WMAX / 3); j < WMAX; ++j) A[j] = 0.0;
            if (!real) {
#pragma unroll
                for (int j = 0; j < WMAX; ++j) A[j] = (j == row) ? 1.0 : 0.0;
            } else {
                const double* Ri = Rm + 9 * (ri + 1);
                double rf = 0.0;
                if (p.fbar) { const double* f = p.fbar + rod * 3 * N + ri + 1; rf = Ri[0 * 3 + rc_comp] * f[0] + Ri[1 * 3 + rc_comp] * f[N] + Ri[2 * 3 + rc_comp] * f[2 * N]; }
                rhs = -rf - dti[ri] * (rc_comp == 0 ? N0[0] : (rc_comp == 1 ? N0[1] : N0[2]));
            }
        }
        // ---- Gauss-Jordan ----------------------------------------------------------------------------------------------------
        bool used = !real, singular = false;
        int srow = WMAX, bad = 0, k = 0, prow = -1;
        double rc = fast_rcp(A[0]);
        {
            unsigned ph, pl;
            int pbest;
            wrench_gjm_post<NW>(A[0], used, row, lane, reinterpret_cast<int*>(sm + C::cand) + warp * 4, ph, pl, pbest);
            if constexpr (NW == 1) {
                prow = pbest; singular = pbest < 0 || (ph | pl) == 0u || ph >= 0x7ff00000u;
            } else {
                __syncthreads();
                wrench_gjm_decide<NW>(reinterpret_cast<const int*>(sm + C::cand), prow, singular);
            }
        }
        wrench_gjm_bodies<NW, WMAX, WMAX>(A, rhs, k, n, row, used, srow, bad, prow, singular, rc, sm, lane, warp);
        if (srow < WMAX) Nl[srow] = rhs;   // the row that was pivot at step s holds unknown s
        __syncthreads();
        // ---- internal couple through the stored multipliers ---------------------------------------------------------------------
        double b = 0.0;
        if (real) {
            const double* Ri = Rm + 9 * (ri + 1);
            double g0 = 1.0, g1 = 0.0, g2 = 0.0;
            if (p.Gamma) { const double* gm = p.Gamma + rod * 3 * N + ri + 1; g0 = gm[0]; g1 = gm[N]; g2 = gm[2 * N]; }
            const double n0 = Nl[3 * ri], n1 = Nl[3 * ri + 1], n2 = Nl[3 * ri + 2];
            const double gx = (rc_comp == 0) ? g1 * n2 - g2 * n1 : (rc_comp == 1 ? g2 * n0 - g0 * n2 : g0 * n1 - g1 * n0);
            double rl = 0.0;
            if (p.lbar) { const double* l = p.lbar + rod * 3 * N + ri + 1; rl = Ri[0 * 3 + rc_comp] * l[0] + Ri[1 * 3 + rc_comp] * l[N] + Ri[2 * 3 + rc_comp] * l[2 * N]; }
            b = -gx - rl - dti[ri] * (rc_comp == 0 ? C0[0] : (rc_comp == 1 ? C0[1] : C0[2]));
        }
        if constexpr (NW == 1) {
#pragma unroll 1
            for (int kq = 0; kq < n; ++kq) {   // the pivot row of step kq hands its current value round by shuffle
                const unsigned own = __ballot_sync(0xffffffffu, srow == kq);
                const double y = own ? __shfl_sync(0xffffffffu, b, __ffs(own) - 1) : 0.0;
                b = fma(-sm[C::L + kq * C::ROWS + row], y, b);
            }
        } else {
#pragma unroll 1
            for (int kq = 0; kq < n; ++kq) {
                if (srow == kq) ysm[kq & 1] = b;   // the pivot row of step kq posts its current value
                __syncthreads();
                b = fma(-sm[C::L + kq * C::ROWS + row], ysm[kq & 1], b);
            }
        }
        if (srow < WMAX) vec[srow] = b;
        __syncthreads();
        // ---- Lambda [6][N]: couple first ---------------------------------------------------------------------------------------
        double* out = p.Lambda + rod * 6 * N;
        if (tid < 3) {
            out[tid * N] = tid == 0 ? C0[0] : (tid == 1 ? C0[1] : C0[2]);
            out[(3 + tid) * N] = tid == 0 ? N0[0] : (tid == 1 ? N0[1] : N0[2]);
        }
        for (int e = tid; e < n; e += 32 * NW) {
            const int i = e / 3, c = e - 3 * i;
            out[c * N + i + 1] = vec[e];
            out[(3 + c) * N + i + 1] = Nl[e];
        }
        if (p.info && tid == 0) p.info[rod] = bad;
    }
}

}  // namespace sri
