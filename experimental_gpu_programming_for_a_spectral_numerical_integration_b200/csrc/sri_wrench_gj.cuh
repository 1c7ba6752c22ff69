// sri_wrench_gj.cuh -- local-frame statics solved directly (rod_modeling.pdf eqs. 1.29, 2.18), N <= 16: one rod
// per warp, the 3M x 3M operator in REGISTERS, Gauss-Jordan with implicit partial pivoting in a ROLLED sliding window.
// The kernel sri_integrate_wrench_local ships for N <= 16 (round 2); sri_wrench_solve.cuh (blocked LU in shared memory with
// DMMA trailing updates; its header states the equations) stays selectable with SRI_WRENCH_IMPL=blocked.
//
//   A = D_TT (x) I3 + blockdiag(K^_i),  n = 3 (N - 1) <= 45 (padded to 48 with identity rows),  A N = b_N,  A C = b_C(N).
//
// Lane l owns rows l and l + 32 (the latter exists for l < 16).  A row is a WINDOW of registers whose slot 0 is always the
// current pivot column: step k takes  a[j-1] = a[j] - l u[j]  (u = pivot row, published once by its owner through shared
// memory and read back as broadcast 128-bit loads), so the same code serves every step and the elimination is a rolled loop
// with static register indices.  Six bodies (window 48, 40, ..., 8; eight steps each) keep the dead tail of the window short;
// together they are ~14 KB of code.  (The fully unrolled register-resident LU tried first was 363 KB of code and
// instruction-fetch bound: 2.0e7 rods/s)
// Gauss-Jordan rather than LU: in this layout every lane executes every row update whether its row is still in play or not,
// so eliminating above the pivot as well costs no extra instruction -- and it removes the U factor, the two backward sweeps
// and their storage.  The pivot row takes the multiplier 1 - 1/pivot, which normalises it with the same instruction stream
// (no divergence); every other row, retired or not, takes a_ik / pivot.  The multipliers of all rows and steps (the whole
// elimination as a product of rank-1 updates) are kept in shared memory, [step][row], 16 KB per rod: the second right-hand
// side -- the couple needs the force first -- is then ONE sweep of 45 fused multiply-adds per row, the pivot value of each
// step broadcast by shuffle.  Rows never move; the row that was pivot at step k ends up holding unknown k.
// Pivot choice: exact warp arg-max of |a_ik| over the rows not yet used, first row on ties -- the pivots of a sequential
// partial-pivot elimination.
#pragma once
#include "sri_fused16.cuh"  // fast_rcp
#include "sri_wrench_solve.cuh"

namespace sri {

constexpr int kWrenchGjWarps = 4;
constexpr int kWrenchGjShared = 256 + 16;         // D_TT row-major 16 x 16 (zero padded), D_TI [16]
struct WrenchGjScratch {                          // per warp, doubles
    static constexpr int L = 0;                   // [45][48] multipliers by (step, row)
    static constexpr int urow = L + 45 * 48;      // [2][52] published pivot row (double buffered): window slots 0..W-1, (RHS, 1/pivot) at 48
    static constexpr int R = urow + 104;          // [16][9] rotation matrices by node (row-major)
    static constexpr int kk = R + 144;            // [3][16] curvature samples
    static constexpr int vec = kk + 48;           // [48] couple solution by unknown index
    static constexpr int Nl = vec + 48;           // [48] force solution by unknown index
    static constexpr int piv = Nl + 48;           // 48 ints: pivot row of every step
    static constexpr int total = piv + 24;
};
static_assert(WrenchGjScratch::total % 2 == 0 && kWrenchGjShared % 2 == 0, "16-byte aligned pieces");
constexpr size_t kWrenchGjSmem = (kWrenchGjShared + (size_t)kWrenchGjWarps * WrenchGjScratch::total) * sizeof(double);

// The pivot decision of a step, taken one step ahead (software pipelining: the arg-max of step k + 1 and every lane's
// reciprocal of its own candidate run behind the bulk update of step k).
struct WrenchGjPivot {
    int prow;         // pivot row of the coming step (-1: no candidate left)
    bool singular;    // its pivot is zero / non-finite
    double rc0, rc1;  // 1 / a_r0,k and 1 / a_r1,k of this lane's rows: the owner publishes the one that was chosen
};

// exact arg-max of |a_i0| over the rows not yet used, first row on ties; every lane inverts its own candidates meanwhile
__device__ __forceinline__ WrenchGjPivot wrench_gj_search(double a0, double a1, bool used0, bool used1) {
    constexpr unsigned FULL = 0xffffffffu;
    WrenchGjPivot pv;
    const double v0 = used0 ? 0.0 : fabs(a0), v1 = used1 ? 0.0 : fabs(a1);
    const unsigned h0 = (unsigned)__double2hiint(v0), l0 = (unsigned)__double2loint(v0);
    const unsigned h1 = (unsigned)__double2hiint(v1), l1 = (unsigned)__double2loint(v1);
    const bool second = h1 > h0 || (h1 == h0 && l1 > l0);
    const unsigned hi = second ? h1 : h0, lo = second ? l1 : l0;
    const unsigned mh = __reduce_max_sync(FULL, hi);
    pv.rc0 = fast_rcp(a0);
    pv.rc1 = fast_rcp(a1);
    const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
    const unsigned m0 = __ballot_sync(FULL, !used0 && h0 == mh && l0 == ml);
    const unsigned m1 = __ballot_sync(FULL, !used1 && h1 == mh && l1 == ml);
    pv.prow = m0 ? __ffs(m0) - 1 : (m1 ? 31 + __ffs(m1) : -1);
    pv.singular = (mh | ml) == 0u || mh >= 0x7ff00000u || pv.prow < 0;
    return pv;
}

// Eight Gauss-Jordan steps (fewer when n is reached) on a window of W live slots; see the header.  `next` holds the pivot
// decision of step k on entry and of the step after the last one on exit.
template <int W>
__device__ __forceinline__ void wrench_gj_body(double (&A0)[48], double (&A1)[48], double& rhs0, double& rhs1, int& k, const int n,
                                               const int r0, const int r1, bool& used0, bool& used1, int& s0, int& s1, int& bad,
                                               WrenchGjPivot& next, double* __restrict__ urow, double* __restrict__ Lsm,
                                               int* __restrict__ piv, const int lane) {
#pragma unroll 1
    for (int s = 0; s < 8 && k < n; ++s, ++k) {
        const int prow = next.prow;
        const bool singular = next.singular;
        if (singular && !bad) bad = k + 1;
        const bool mine0 = prow == r0, mine1 = prow == r1;
        // the owner publishes its window (slots 0..W-1), its right-hand side and the reciprocal of the pivot
        double* ub = urow + (k & 1) * 52;
        if (mine0) {
#pragma unroll
            for (int j = 0; j < W; j += 2) *reinterpret_cast<double2*>(ub + j) = make_double2(A0[j], A0[j + 1]);
            *reinterpret_cast<double2*>(ub + 48) = make_double2(rhs0, next.rc0);
            piv[k] = prow;
            used0 = true; s0 = k;
        } else if (mine1) {
#pragma unroll
            for (int j = 0; j < W; j += 2) *reinterpret_cast<double2*>(ub + j) = make_double2(A1[j], A1[j + 1]);
            *reinterpret_cast<double2*>(ub + 48) = make_double2(rhs1, next.rc1);
            piv[k] = prow;
            used1 = true; s1 = k;
        }
        if (prow < 0 && lane == 0) { piv[k] = 0; *reinterpret_cast<double2*>(ub + 48) = make_double2(0.0, 0.0); }
        __syncwarp();
        const double2 tail = *reinterpret_cast<const double2*>(ub + 48);   // (right-hand side of the pivot row, 1 / pivot)
        const double inv = singular ? 0.0 : tail.y;
        const double m0l = mine0 ? 1.0 - inv : A0[0] * inv;   // the pivot row is normalised by the same update
        const double m1l = mine1 ? 1.0 - inv : A1[0] * inv;
        Lsm[k * 48 + r0] = m0l;
        if (r1 < 48) Lsm[k * 48 + r1] = m1l;
        // the next pivot column first, so that its arg-max and reciprocals hide behind the rest of the update
        {
            const double2 u = *reinterpret_cast<const double2*>(ub);
            A0[0] = fma(-m0l, u.y, A0[1]);
            A1[0] = fma(-m1l, u.y, A1[1]);
        }
        if (k + 1 < n) next = wrench_gj_search(A0[0], A1[0], used0, used1);
#pragma unroll
        for (int j = 2; j < W; j += 2) {
            const double2 u = *reinterpret_cast<const double2*>(ub + j);
            A0[j - 1] = fma(-m0l, u.x, A0[j]); A1[j - 1] = fma(-m1l, u.x, A1[j]);
            A0[j] = fma(-m0l, u.y, A0[j + 1]); A1[j] = fma(-m1l, u.y, A1[j + 1]);
        }
        A0[W - 1] = 0.0; A1[W - 1] = 0.0;
        rhs0 = fma(-m0l, tail.x, rhs0); rhs1 = fma(-m1l, tail.x, rhs1);
    }
}

__global__ void __launch_bounds__(32 * kWrenchGjWarps, 2) wrench_local_solve_gj_kernel(const WrenchParams p) {
    extern __shared__ __align__(16) double wsm[];
    double* dtt = wsm;          // [16][16] row-major
    double* dti = wsm + 256;    // [16]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* scr = wsm + kWrenchGjShared + warp * WrenchGjScratch::total;
    double* Lsm = scr + WrenchGjScratch::L;
    double* urow = scr + WrenchGjScratch::urow;
    double* Rm = scr + WrenchGjScratch::R;
    double* kk = scr + WrenchGjScratch::kk;
    double* vec = scr + WrenchGjScratch::vec;
    double* Nl = scr + WrenchGjScratch::Nl;
    int* piv = reinterpret_cast<int*>(scr + WrenchGjScratch::piv);
    const int N = p.N, M = p.M, n = 3 * M;
    constexpr unsigned FULL = 0xffffffffu;
    for (int e = threadIdx.x; e < 256; e += blockDim.x) {
        const int i = e >> 4, j = e & 15;
        dtt[e] = (i < M && j < M) ? p.D_TT[(size_t)j * M + i] : 0.0;
    }
    for (int e = threadIdx.x; e < 16; e += blockDim.x) dti[e] = e < M ? p.D_TI[e] : 0.0;
    __syncthreads();

    // rows of this lane: r0 = lane, r1 = lane + 32 (exists for lanes 0..15); row r = (node r / 3 + 1, component r % 3)
    const int r0 = lane, r1 = lane + 32;
    const bool has1 = lane < 16;
    const int i0 = r0 / 3, c0 = r0 - 3 * i0, i1 = r1 / 3, c1 = r1 - 3 * i1;

    const long long warps_total = (long long)gridDim.x * kWrenchGjWarps;
    const long long count = p.from_list ? (long long)*p.rod_count : p.batch;
    for (long long it = (long long)blockIdx.x * kWrenchGjWarps + warp; it < count; it += warps_total) {
        const long long rod = p.from_list ? (long long)p.rod_list[it] : it;
        // ---- this rod's curvature samples, rotations by node -----------------------------------------------------------------
        for (int e = lane; e < 48; e += 32) { const int c = e >> 4, i = e & 15; kk[e] = i < N ? p.K[rod * 3 * N + c * N + i] : 0.0; }
        if (lane < 16) {
            quat q; q.w = 1.0; q.x = 0.0; q.y = 0.0; q.z = 0.0;
            if (lane < M) { const double* s = p.Q + rod * 4 * M + lane; q.w = s[0]; q.x = s[M]; q.y = s[2 * M]; q.z = s[3 * M]; }
            else if (lane == M && p.q0) { const double* s = p.q0 + rod * 4; q.w = s[0]; q.x = s[1]; q.y = s[2]; q.z = s[3]; }
            quat_to_rot_rm(q, Rm + 9 * lane);
        }
        __syncwarp();
        double N0[3], C0[3];
        {
            const double* F = p.F_tip + rod * 3; const double* T = p.M_tip + rod * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                N0[c] = Rm[0 * 3 + c] * F[0] + Rm[1 * 3 + c] * F[1] + Rm[2 * 3 + c] * F[2];
                C0[c] = Rm[0 * 3 + c] * T[0] + Rm[1 * 3 + c] * T[1] + Rm[2 * 3 + c] * T[2];
            }
        }
        // ---- operator rows in registers (identity on the padding rows >= n) and the first right-hand side ---------------------
        double A0[48], A1[48], rhs0, rhs1;
        auto build_row = [&](double (&a)[48], double& rhs, int r, int i, int ac, bool exists) {
            const bool real = exists && r < n;
            // K^ of node i+1, row ac: [[0,-k2,k1],[k2,0,-k0],[-k1,k0,0]]
            const double k0 = real ? kk[i + 1] : 0.0, k1 = real ? kk[16 + i + 1] : 0.0, k2 = real ? kk[32 + i + 1] : 0.0;
            const double kh0 = ac == 0 ? 0.0 : (ac == 1 ? k2 : -k1);
            const double kh1 = ac == 0 ? -k2 : (ac == 1 ? 0.0 : k0);
            const double kh2 = ac == 0 ? k1 : (ac == 1 ? -k0 : 0.0);
#pragma unroll
            for (int jn = 0; jn < 16; ++jn) {
                const double d = real ? dtt[i * 16 + jn] : 0.0;
                const bool diag = real && jn == i;
                a[3 * jn + 0] = (ac == 0 ? d : 0.0) + (diag ? kh0 : 0.0);
                a[3 * jn + 1] = (ac == 1 ? d : 0.0) + (diag ? kh1 : 0.0);
                a[3 * jn + 2] = (ac == 2 ? d : 0.0) + (diag ? kh2 : 0.0);
            }
            if (!real) {
#pragma unroll
                for (int j = 0; j < 48; ++j) a[j] = (exists && j == r) ? 1.0 : 0.0;
            }
            // internal force: -R_i^T fbar_i - D_TI N0
            double v = 0.0;
            if (real) {
                const double* Ri = Rm + 9 * (i + 1);
                double rf = 0.0;
                if (p.fbar) { const double* f = p.fbar + rod * 3 * N + i + 1; rf = Ri[0 * 3 + ac] * f[0] + Ri[1 * 3 + ac] * f[N] + Ri[2 * 3 + ac] * f[2 * N]; }
                v = -rf - dti[i] * (ac == 0 ? N0[0] : (ac == 1 ? N0[1] : N0[2]));
            }
            rhs = v;
        };
        build_row(A0, rhs0, r0, i0, c0, true);
        build_row(A1, rhs1, r1, i1, c1, has1);

        // ---- Gauss-Jordan with implicit partial pivoting: six rolled bodies over a shrinking window ----------------------------
        bool used0 = r0 >= n, used1 = !has1 || r1 >= n;   // padding rows are never pivots (their columns are never eliminated)
        int s0 = 48, s1 = 48, bad = 0, k = 0;
        WrenchGjPivot next = wrench_gj_search(A0[0], A1[0], used0, used1);
        wrench_gj_body<48>(A0, A1, rhs0, rhs1, k, n, r0, r1, used0, used1, s0, s1, bad, next, urow, Lsm, piv, lane);
        wrench_gj_body<40>(A0, A1, rhs0, rhs1, k, n, r0, r1, used0, used1, s0, s1, bad, next, urow, Lsm, piv, lane);
        wrench_gj_body<32>(A0, A1, rhs0, rhs1, k, n, r0, r1, used0, used1, s0, s1, bad, next, urow, Lsm, piv, lane);
        wrench_gj_body<24>(A0, A1, rhs0, rhs1, k, n, r0, r1, used0, used1, s0, s1, bad, next, urow, Lsm, piv, lane);
        wrench_gj_body<16>(A0, A1, rhs0, rhs1, k, n, r0, r1, used0, used1, s0, s1, bad, next, urow, Lsm, piv, lane);
        wrench_gj_body<8>(A0, A1, rhs0, rhs1, k, n, r0, r1, used0, used1, s0, s1, bad, next, urow, Lsm, piv, lane);
        __syncwarp();
        // the row that was pivot at step s holds unknown s
        if (s0 < 48) Nl[s0] = rhs0;
        if (s1 < 48) Nl[s1] = rhs1;
        __syncwarp();
        // ---- internal couple: b = -Gamma_i x N_i - R_i^T lbar_i - D_TI C0, through the stored multipliers ----------------------
        auto couple_rhs = [&](int r, int i, int ac, bool exists) {
            if (!(exists && r < n)) return 0.0;
            const double* Ri = Rm + 9 * (i + 1);
            double g0 = 1.0, g1 = 0.0, g2 = 0.0;
            if (p.Gamma) { const double* gm = p.Gamma + rod * 3 * N + i + 1; g0 = gm[0]; g1 = gm[N]; g2 = gm[2 * N]; }
            const double n0 = Nl[3 * i], n1 = Nl[3 * i + 1], n2 = Nl[3 * i + 2];
            const double gx = (ac == 0) ? g1 * n2 - g2 * n1 : (ac == 1 ? g2 * n0 - g0 * n2 : g0 * n1 - g1 * n0);
            double rl = 0.0;
            if (p.lbar) { const double* l = p.lbar + rod * 3 * N + i + 1; rl = Ri[0 * 3 + ac] * l[0] + Ri[1 * 3 + ac] * l[N] + Ri[2 * 3 + ac] * l[2 * N]; }
            return -gx - rl - dti[i] * (ac == 0 ? C0[0] : (ac == 1 ? C0[1] : C0[2]));
        };
        double b0 = couple_rhs(r0, i0, c0, true), b1 = couple_rhs(r1, i1, c1, has1);
#pragma unroll 4
        for (int kq = 0; kq < n; ++kq) {
            const int prow = piv[kq];
            const double yk = __shfl_sync(FULL, (prow & 32) ? b1 : b0, prow & 31);
            b0 = fma(-Lsm[kq * 48 + r0], yk, b0);
            if (has1) b1 = fma(-Lsm[kq * 48 + r1], yk, b1);
        }
        if (s0 < 48) vec[s0] = b0;
        if (s1 < 48) vec[s1] = b1;
        __syncwarp();
        // ---- Lambda [6][N]: couple first ---------------------------------------------------------------------------------------
        double* out = p.Lambda + rod * 6 * N;
        if (lane < 3) {
            out[lane * N] = lane == 0 ? C0[0] : (lane == 1 ? C0[1] : C0[2]);
            out[(3 + lane) * N] = lane == 0 ? N0[0] : (lane == 1 ? N0[1] : N0[2]);
        }
        for (int e = lane; e < n; e += 32) {
            const int i = e / 3, c = e - 3 * i;
            out[c * N + i + 1] = vec[e];
            out[(3 + c) * N + i + 1] = Nl[e];
        }
        if (p.info && lane == 0) p.info[rod] = bad;
        __syncwarp();
    }
}

}  // namespace sri
