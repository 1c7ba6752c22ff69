// sri_wrench_solve.cuh -- local-frame statics solved directly (rod_modeling.pdf eqs. 1.29, 2.18), N <= 16.
//
//   N' = -K^ N - R^T fbar,   C' = -K^ C - Gamma^ N - R^T lbar,   N(1) = R(1)^T F_tip,  C(1) = R(1)^T M_tip
// collocated on the Chebyshev nodes with the tip node eliminated: the strain-DEPENDENT real operator
//   A = D_TT (x) I3 + blockdiag(K^_i),   3M x 3M  (45 x 45 at N = 16, padded with an identity block to 48 x 48),
// one partial-pivot LU per rod, two solves (the couple's right-hand side needs N).  The 3 x 3 blocks K^ act on both
// sides of the quaternion algebra (v -> K x v is a commutator), so this system does not fit the one-sided quaternion
// elimination of the fused kernels; it is a plain real LU, one rod per warp, matrix in shared memory.
//
// Blocked right-looking LU, block width 4 = the k extent of the FP64 tensor instruction:
//   panel   48 x 4 in registers, lanes own rows (lane, lane + 32); pivot search = warp arg-max (largest magnitude,
//           first on ties: the pivots of a sequential LU), the two rows are exchanged through shuffles;
//   swaps   applied to the other columns with the lanes owning columns; column 48 of the array carries the row index,
//           so after the factorisation it holds the permutation and a right-hand side is permuted by one gather;
//   U12     = L11^-1 A12, the lane that owns the column, 6 FMAs;
//   A22    -= L21 U12 on the tensor pipe: one DMMA m8n8k4 per 8 x 8 tile, the accumulator fragment is a 128-bit
//           shared-memory load/store (leading dimension 56 = 8 mod 16 doubles keeps those conflict-free).
// The triangular solves keep the right-hand side in registers (lanes own rows), four unknowns per step: the block's
// entries are broadcast by shuffle, every lane solves the 4 x 4 triangle redundantly and updates its own rows.
// A sequential CPU restatement (unblocked, same pivot choice) agrees to round-off (the blocked update only changes the
// order of the additions).
#pragma once
#include "sri_device.cuh"
#include "sri_fused16.cuh"     // FusedParams
#include "sri_stage_dmma.cuh"  // dmma_m8n8k4
#include <type_traits>

namespace sri {

constexpr int kWrenchWarps = 8;               // one CTA per SM
constexpr int kWrenchNP = 48;                 // 3 (N - 1) rounded up to whole 8 x 8 tiles, N <= 16
constexpr int kWrenchLD = 56;                 // doubles per row: 48 columns, the row-index column, padding
constexpr int kWrenchShared = 16 + kWrenchNP * kWrenchLD;  // D_TI [15]; the strain-independent part of A, shared by the warps
struct WrenchScratch {                        // per warp, doubles
    static constexpr int A = 0;                              // [48][56]
    static constexpr int b = A + kWrenchNP * kWrenchLD;      // [48] right-hand side / solution
    static constexpr int Nl = b + 48;                        // [48] local force (kept for the couple's right-hand side)
    static constexpr int R = Nl + 48;                        // [16][9] rotation matrices by node (row-major)
    static constexpr int kk = R + 144;                       // [3][N] curvature samples of this rod
    static constexpr int dinv = kk + 48;                     // [48] reciprocal pivots
    static constexpr int total = dinv + 48;
};
static_assert(WrenchScratch::total % 2 == 0 && kWrenchShared % 2 == 0, "16-byte aligned rows");
constexpr size_t kWrenchSmem = (kWrenchShared + (size_t)kWrenchWarps * WrenchScratch::total) * sizeof(double);

struct WrenchParams {
    long long batch;
    int N, M;
    const double* D_TT;  // [M][M] column-major (as sri_get_operator(4))
    const double* D_TI;  // [M]
    const double *K, *Q, *q0, *Gamma, *fbar, *lbar, *F_tip, *M_tip;
    double* Lambda;      // [batch][6][N]
    int* info;
    // static-order first pass + row-pivoting second pass (sri_wrench_gj_static.cuh); all null / zero otherwise
    const double* S;     // [M][M] D_TT^-1, column-major
    int* rod_count;      // hand-back list: number of rods, then (rod_list) their indices
    int* rod_list;
    int growth_hi;       // bound on |multiplier| as the high word of the double
    int from_list;       // 1: this launch processes the rods of the list instead of 0..batch-1
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

__global__ void __launch_bounds__(32 * kWrenchWarps, 1) wrench_local_solve_kernel(const WrenchParams p) {
    extern __shared__ __align__(16) double wsm[];
    double* dti = wsm;          // [15]
    double* tmpl = wsm + 16;    // [48][56]: D_TT (x) I3, identity on the padding rows, row index in column 48
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lr = lane >> 2, lk = lane & 3;
    double* scr = wsm + kWrenchShared + warp * WrenchScratch::total;
    double* A = scr + WrenchScratch::A;
    double* b = scr + WrenchScratch::b;
    double* Nl = scr + WrenchScratch::Nl;
    double* Rm = scr + WrenchScratch::R;
    double* kk = scr + WrenchScratch::kk;
    double* dinv = scr + WrenchScratch::dinv;
    const int N = p.N, M = p.M, n = 3 * M;
    const int NT = (n + 7) >> 3, NP = 8 * NT, NB = NP >> 2;  // tiles, padded order, blocks of 4 columns
    constexpr int LD = kWrenchLD;
    constexpr unsigned FULL = 0xffffffffu;
    const int row1 = lane + 32;
    const bool has0 = lane < NP, has1 = row1 < NP;           // the rows this lane owns
    for (int e = threadIdx.x; e < kWrenchNP * LD; e += blockDim.x) {
        const int r = e / LD, c = e - r * LD;
        double v = 0.0;
        if (r < n && c < n) { const int i = r / 3, j = c / 3; if (r - 3 * i == c - 3 * j) v = p.D_TT[j * M + i]; }
        else if (c == 48) v = (double)r;
        else if (r == c) v = 1.0;
        tmpl[e] = v;
    }
    for (int i = threadIdx.x; i < M; i += blockDim.x) dti[i] = p.D_TI[i];
    __syncthreads();

    // Row ownership: while rows 0..15 are still in play the lanes own two rows (lane, lane + 32); once the elimination has
    // passed row 16 (or, backwards, has left rows 32..47) one row per lane is enough -- row `rowB` below, or row `lane`.
    const int rowB = lane < 16 ? row1 : lane;                // rows 16..47 over the 32 lanes
    const bool hasB = rowB < NP;

    // b (original row order) -> solution in b, with the factors in A; the vector lives in registers in between
    auto solve = [&]() {
        double y0 = has0 ? b[(int)A[lane * LD + 48]] : 0.0;
        double y1 = has1 ? b[(int)A[row1 * LD + 48]] : 0.0;
        auto block_values = [&](double ys, int c0, double& t0, double& t1, double& t2, double& t3) {
            const int src = c0 & 31;
            t0 = __shfl_sync(FULL, ys, src); t1 = __shfl_sync(FULL, ys, src + 1);
            t2 = __shfl_sync(FULL, ys, src + 2); t3 = __shfl_sync(FULL, ys, src + 3);
        };
        // a row inside the block [c0, c0 + 4) takes the block's solved value: its offset is row & 3 = lane & 3 for both
        // rows of a lane (c0 and 32 are multiples of 4), so the selection is made once per block with loop-invariant predicates
        const int l3 = lane & 3;
        auto row_update = [&](int row, int c0, bool beyond, double t0, double t1, double t2, double t3, double& y) {
            const int d = row - c0;
            if (beyond) {
                const double2 la = *reinterpret_cast<const double2*>(A + row * LD + c0);
                const double2 lb = *reinterpret_cast<const double2*>(A + row * LD + c0 + 2);
                y = fma(-lb.y, t3, fma(-lb.x, t2, fma(-la.y, t1, fma(-la.x, t0, y))));
            } else if (d >= 0 && d < 4) {
                y = l3 == 0 ? t0 : (l3 == 1 ? t1 : (l3 == 2 ? t2 : t3));
            }
        };
        auto lower4 = [&](int c0, double t0, double& t1, double& t2, double& t3) {  // unit lower 4 x 4 triangle
            const double* Lb = A + c0 * LD + c0;
            t1 = fma(-Lb[LD], t0, t1);
            t2 = fma(-Lb[2 * LD + 1], t1, fma(-Lb[2 * LD], t0, t2));
            t3 = fma(-Lb[3 * LD + 2], t2, fma(-Lb[3 * LD + 1], t1, fma(-Lb[3 * LD], t0, t3)));
        };
        auto upper4 = [&](int c0, double& t0, double& t1, double& t2, double& t3) {
            const double* Ub = A + c0 * LD + c0;
            t3 *= dinv[c0 + 3];
            t2 = fma(-Ub[2 * LD + 3], t3, t2) * dinv[c0 + 2];
            t1 = fma(-Ub[LD + 3], t3, fma(-Ub[LD + 2], t2, t1)) * dinv[c0 + 1];
            t0 = fma(-Ub[3], t3, fma(-Ub[2], t2, fma(-Ub[1], t1, t0))) * dinv[c0];
        };
        double t0, t1, t2, t3;
        // ---- L y = P b ------------------------------------------------------------------------------------------------
        const int nbA = NB < 4 ? NB : 4;
        for (int kb = 0; kb < nbA; ++kb) {                   // columns 0..15: two rows per lane
            const int c0 = 4 * kb;
            block_values(y0, c0, t0, t1, t2, t3);
            lower4(c0, t0, t1, t2, t3);
            if (has0) row_update(lane, c0, lane >= c0 + 4, t0, t1, t2, t3, y0);
            if (has1) row_update(row1, c0, true, t0, t1, t2, t3, y1);
        }
        if (NB > 4) {                                        // columns 16..: rows 16..47, one per lane
            double ya = lane < 16 ? y1 : y0;
            for (int kb = 4; kb < NB; ++kb) {
                const int c0 = 4 * kb;
                block_values(ya, c0, t0, t1, t2, t3);
                lower4(c0, t0, t1, t2, t3);
                if (hasB) row_update(rowB, c0, rowB >= c0 + 4, t0, t1, t2, t3, ya);
            }
            if (lane < 16) y1 = ya; else y0 = ya;
        }
        // ---- U x = y ---------------------------------------------------------------------------------------------------
        for (int kb = NB - 1; kb >= 8; --kb) {               // columns 32..47: both rows of a lane
            const int c0 = 4 * kb;
            block_values(y1, c0, t0, t1, t2, t3);
            upper4(c0, t0, t1, t2, t3);
            if (has0) row_update(lane, c0, true, t0, t1, t2, t3, y0);
            if (has1) row_update(row1, c0, row1 < c0, t0, t1, t2, t3, y1);
        }
        for (int kb = (NB < 8 ? NB : 8) - 1; kb >= 0; --kb) { // columns 0..31: rows 0..31
            const int c0 = 4 * kb;
            block_values(y0, c0, t0, t1, t2, t3);
            upper4(c0, t0, t1, t2, t3);
            if (has0) row_update(lane, c0, lane < c0, t0, t1, t2, t3, y0);
        }
        __syncwarp();
        if (has0) b[lane] = y0;
        if (has1) b[row1] = y1;
        __syncwarp();
    };

    const long long warps_total = (long long)gridDim.x * kWrenchWarps;
    for (long long rod = (long long)blockIdx.x * kWrenchWarps + warp; rod < p.batch; rod += warps_total) {
        // ---- this rod's curvature samples, rotations by node ---------------------------------------------------------
        for (int e = lane; e < 3 * N; e += 32) kk[e] = p.K[rod * 3 * N + e];
        if (lane <= M) {
            quat q; q.w = 1.0; q.x = 0.0; q.y = 0.0; q.z = 0.0;
            if (lane < M) { const double* s = p.Q + rod * 4 * M + lane; q.w = s[0]; q.x = s[M]; q.y = s[2 * M]; q.z = s[3 * M]; }
            else if (p.q0) { const double* s = p.q0 + rod * 4; q.w = s[0]; q.x = s[1]; q.y = s[2]; q.z = s[3]; }
            quat_to_rot_rm(q, Rm + 9 * lane);
        }
        __syncwarp();
        // ---- operator: the CTA's template plus this rod's 6 M off-diagonal entries of blockdiag(K^) -----------------
        for (int e = lane; e < NP * (LD / 2); e += 32) reinterpret_cast<double2*>(A)[e] = reinterpret_cast<const double2*>(tmpl)[e];
        __syncwarp();
        for (int e = lane; e < 6 * M; e += 32) {  // K^ of node i+1: [[0,-k2,k1],[k2,0,-k0],[-k1,k0,0]]
            const int i = e / 6, k = e - 6 * i;
            const int a = k >> 1, bb = (a + 1 + (k & 1)) % 3;   // the two off-diagonal columns of row a
            const double kv = kk[(3 - a - bb) * N + i + 1];
            A[(3 * i + a) * LD + 3 * i + bb] = (bb == a + 2 || a == bb + 1) ? kv : -kv;
        }
        __syncwarp();
        // ---- blocked partial-pivot LU, in place ------------------------------------------------------------------------
        int bad = 0;
        for (int kb = 0; kb < NB; ++kb) {
            const int c0 = 4 * kb;
            int pr[4];          // pivot rows of this panel (positions before the panel's exchanges)
            double pv[4][4];    // their panel entries at the time they were chosen: L11 below the diagonal, U11 on and above
            // ---- panel: rows stay where they are, a row that has served as pivot retires (implicit pivoting) -----------
            auto panel = [&](auto two_rows) {
                constexpr bool TWO = decltype(two_rows)::value;
                const int ra = TWO ? lane : rowB;
                bool act0 = (TWO ? has0 : hasB) && ra >= c0, act1 = TWO && has1;
                double p0[4] = {0.0, 0.0, 0.0, 0.0}, p1[4] = {0.0, 0.0, 0.0, 0.0};
                if (act0) {
                    const double2 u = *reinterpret_cast<const double2*>(A + ra * LD + c0), w = *reinterpret_cast<const double2*>(A + ra * LD + c0 + 2);
                    p0[0] = u.x; p0[1] = u.y; p0[2] = w.x; p0[3] = w.y;
                }
                if (act1) {
                    const double2 u = *reinterpret_cast<const double2*>(A + row1 * LD + c0), w = *reinterpret_cast<const double2*>(A + row1 * LD + c0 + 2);
                    p1[0] = u.x; p1[1] = u.y; p1[2] = w.x; p1[3] = w.y;
                }
                const bool st0 = act0, st1 = act1;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // arg-max of |column| over the rows in play, exact, smallest row on ties: maximum of the high words,
                    // then of the low words among those, then the first lane in row order
                    const double v0 = act0 ? fabs(p0[j]) : 0.0, v1 = act1 ? fabs(p1[j]) : 0.0;
                    const unsigned h0 = (unsigned)__double2hiint(v0), l0 = (unsigned)__double2loint(v0);
                    const unsigned h1 = (unsigned)__double2hiint(v1), l1 = (unsigned)__double2loint(v1);
                    const bool second = TWO && (h1 > h0 || (h1 == h0 && l1 > l0));
                    const unsigned hi = second ? h1 : h0, lo = second ? l1 : l0;
                    const unsigned mh = __reduce_max_sync(FULL, hi);
                    const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
                    const unsigned m0 = __ballot_sync(FULL, act0 && h0 == mh && l0 == ml);
                    int bi;
                    if (TWO) {
                        const unsigned m1 = __ballot_sync(FULL, act1 && h1 == mh && l1 == ml);
                        bi = m0 ? __ffs(m0) - 1 : 31 + __ffs(m1);
                    } else {
                        const unsigned rot = (m0 >> 16) | (m0 << 16);  // row order: lanes 16..31 (rows 16..31), then 0..15
                        bi = (__ffs(rot) - 1 + 16) & 31;
                        bi = bi < 16 ? bi + 32 : bi;
                    }
                    pr[j] = bi;
                    const bool from1 = TWO && bi >= 32;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) pv[j][jj] = __shfl_sync(FULL, from1 ? p1[jj] : p0[jj], bi & 31);
                    if (lane == (bi & 31)) { if (from1) act1 = false; else act0 = false; }
                    const bool singular = (mh | ml) == 0u || mh >= 0x7ff00000u;
                    if (singular && !bad && c0 + j < n) bad = c0 + j + 1;
                    const double rp = __drcp_rn(pv[j][j]);
                    const double inv = singular ? 0.0 : rp;
                    if (lane == 0) dinv[c0 + j] = rp;
                    if (act0) {
                        const double l = p0[j] * inv; p0[j] = l;
#pragma unroll
                        for (int jj = j + 1; jj < 4; ++jj) p0[jj] = fma(-l, pv[j][jj], p0[jj]);
                    }
                    if (act1) {
                        const double l = p1[j] * inv; p1[j] = l;
#pragma unroll
                        for (int jj = j + 1; jj < 4; ++jj) p1[jj] = fma(-l, pv[j][jj], p1[jj]);
                    }
                }
                if (st0) {
                    *reinterpret_cast<double2*>(A + ra * LD + c0) = make_double2(p0[0], p0[1]);
                    *reinterpret_cast<double2*>(A + ra * LD + c0 + 2) = make_double2(p0[2], p0[3]);
                }
                if (st1) {
                    *reinterpret_cast<double2*>(A + row1 * LD + c0) = make_double2(p1[0], p1[1]);
                    *reinterpret_cast<double2*>(A + row1 * LD + c0 + 2) = make_double2(p1[2], p1[3]);
                }
            };
            if (c0 < 16) panel(std::true_type{}); else panel(std::false_type{});
            __syncwarp();
            // ---- the exchanges that bring pivot row j to position c0 + j, whole rows (factors to the left, panel, trailing
            //      block, row-index column), the lanes owning columns; U12 = L11^-1 A12 on the way.  Written as one gather
            //      (all loads, then all stores) of the arrangement that the sequence of exchanges j <-> q[j] produces:
            //      position c0 + j receives pivot row j, a displaced row c0 + j that is no pivot goes to dest[j] ---------------
            int dest[4];
            bool disp[4], moved = false;
            {
                int q[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {  // where pivot row j sits after the exchanges 0..j-1
                    int t = pr[j];
#pragma unroll
                    for (int jp = 0; jp < j; ++jp) if (t == c0 + jp) t = q[jp];
                    q[j] = t;
                    moved = moved || pr[j] != c0 + j;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    disp[j] = pr[0] != c0 + j && pr[1] != c0 + j && pr[2] != c0 + j && pr[3] != c0 + j;
                    int t = c0 + j;
#pragma unroll
                    for (int jj = j; jj < 4; ++jj) if (t == c0 + jj) t = q[jj];
                    dest[j] = t;
                }
            }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int c = lane + 32 * s;
                const bool trailing = c >= c0 + 4 && c < NP;
                if ((c < NP || c == 48) && (moved || trailing)) {
                    double u[4], d[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                    for (int j = 0; j < 4; ++j) u[j] = A[pr[j] * LD + c];
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (disp[j]) d[j] = A[(c0 + j) * LD + c];
                    if (trailing) {
                        u[1] = fma(-pv[1][0], u[0], u[1]);
                        u[2] = fma(-pv[2][1], u[1], fma(-pv[2][0], u[0], u[2]));
                        u[3] = fma(-pv[3][2], u[2], fma(-pv[3][1], u[1], fma(-pv[3][0], u[0], u[3])));
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) A[(c0 + j) * LD + c] = u[j];
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (disp[j]) A[dest[j] * LD + c] = d[j];
                }
            }
            __syncwarp();
            // trailing block A22 -= L21 U12: tiles from the one that holds row/column c0 + 4; entries of those tiles
            // outside the trailing block see a zero multiplier and keep their value
            const int t0 = (c0 + 4) >> 3;
            if (t0 < NT) {
                double bf[6];
#pragma unroll
                for (int nt = 0; nt < 6; ++nt) {
                    const int c = 8 * nt + lr;
                    bf[nt] = (nt >= t0 && nt < NT && c >= c0 + 4) ? A[(c0 + lk) * LD + c] : 0.0;
                }
                for (int mt = t0; mt < NT; ++mt) {
                    const int row = 8 * mt + lr;
                    const double af = (row >= c0 + 4) ? -A[row * LD + c0 + lk] : 0.0;
#pragma unroll
                    for (int nt = 0; nt < 6; ++nt) {
                        if (nt >= t0 && nt < NT) {
                            double2* cp = reinterpret_cast<double2*>(A + row * LD + 8 * nt + 2 * lk);
                            double2 acc = *cp;
                            dmma_m8n8k4(acc.x, acc.y, af, bf[nt]);
                            *cp = acc;
                        }
                    }
                }
                __syncwarp();
            }
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
        for (int e = lane; e < NP; e += 32) {
            double v = 0.0;
            if (e < n) {
                const int i = e / 3, c = e - 3 * i;
                const double* Ri = Rm + 9 * (i + 1);
                double rf = 0.0;
                if (p.fbar) { const double* f = p.fbar + rod * 3 * N + i + 1; rf = Ri[0 * 3 + c] * f[0] + Ri[1 * 3 + c] * f[N] + Ri[2 * 3 + c] * f[2 * N]; }
                v = -rf - dti[i] * (c == 0 ? N0[0] : (c == 1 ? N0[1] : N0[2]));
            }
            b[e] = v;
        }
        __syncwarp();
        solve();
        for (int e = lane; e < NP; e += 32) Nl[e] = b[e];
        __syncwarp();
        // ---- internal couple: b = -Gamma_i x N_i - R_i^T lbar_i - D_TI C0 -----------------------------------------------
        double bc[2] = {0.0, 0.0};
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int e = lane + 32 * s;
            if (e < n) {
                const int i = e / 3, c = e - 3 * i;
                const double* Ri = Rm + 9 * (i + 1);
                double g[3] = {1.0, 0.0, 0.0};
                if (p.Gamma) { const double* gm = p.Gamma + rod * 3 * N + i + 1; g[0] = gm[0]; g[1] = gm[N]; g[2] = gm[2 * N]; }
                const double n0 = Nl[3 * i], n1 = Nl[3 * i + 1], n2 = Nl[3 * i + 2];
                const double gx = (c == 0) ? g[1] * n2 - g[2] * n1 : (c == 1 ? g[2] * n0 - g[0] * n2 : g[0] * n1 - g[1] * n0);
                double rl = 0.0;
                if (p.lbar) { const double* l = p.lbar + rod * 3 * N + i + 1; rl = Ri[0 * 3 + c] * l[0] + Ri[1 * 3 + c] * l[N] + Ri[2 * 3 + c] * l[2 * N]; }
                bc[s] = -gx - rl - dti[i] * (c == 0 ? C0[0] : (c == 1 ? C0[1] : C0[2]));
            }
        }
        if (has0) b[lane] = bc[0];
        if (has1) b[row1] = bc[1];
        __syncwarp();
        solve();
        // ---- Lambda [6][N]: couple first ------------------------------------------------------------------------------
        double* out = p.Lambda + rod * 6 * N;
        if (lane < 3) {
            out[lane * N] = lane == 0 ? C0[0] : (lane == 1 ? C0[1] : C0[2]);
            out[(3 + lane) * N] = lane == 0 ? N0[0] : (lane == 1 ? N0[1] : N0[2]);
        }
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
