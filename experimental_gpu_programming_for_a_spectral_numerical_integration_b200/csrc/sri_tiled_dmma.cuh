// sri_tiled_dmma.cuh -- fused four-stage kernel for 17 <= N <= 64 with the quaternion elimination on the FP64 tensor
// cores: the multi-warp form of sri_fused16_dmma.cuh.  ONE rod per CTA.
//
// Same algebra and fragment layout as the N <= 16 kernel (see that file and tools/dmma_gj_emulator.py): the M x M
// quaternion system  sum_j Q_j (x) c_ij = b_i  is the real matrix  Cr[4 i + r][j]  (last column = b), held as 8 x 8 DMMA
// accumulator tiles; Gauss-Jordan step k is  Cr -= [Rmat(c_ik)]_i * (-(U (x) conj c_kk) / |c_kk|^2).
// What changes:
//   * the 8-row tiles are dealt CYCLICALLY to the W warps of the CTA (row tile T lives in warp T mod W, local index
//     T / W), RT tiles x CT column tiles x 2 doubles per lane (32 doubles for both instantiations:
//     <RT=4, CT=4, W=4> for N <= 32 and <RT=2, CT=8, W=16> for N <= 64);
//   * with that dealing the local tile index of pivot row k, (4 (k/8)) / W, is a compile-time constant inside the body
//     unrolled over the pivot's column tile k/8 and its parity k%2, so every register index is static while the
//     loop over (k%8)/2 stays a run-time loop (the code stays small);
//   * the warp that owns the pivot row normalises it and publishes the B fragments through shared memory (double
//     buffered), one CTA barrier per step; every warp gathers the pivot column for its own tiles by shuffles.
// Pivoting: static order + growth check + hand-back to the row-pivoting kernel (sri_tiled.cuh), as for N <= 16.
// Measured alternatives to the plain barrier per step: owner look-ahead (-2..7 %) and a release/acquire
// publication counter with a 16-slot ring instead of the barrier (-3 %; with the owner's sub-partition peers held back
// until its critical DMMAs are issued +3 % at N = 64, -22 % at N = 32), the same with mbarrier try_wait instead of polling
// (-10 %, -21 %), barrier + look-ahead + peers held back (-13 %), barrier + early pivot element so that the reciprocal
// chain hides behind the bulk DMMAs (-10 %).  bar.sync waits cost no issue slots and resume at once, and one instruction
// stream shared by all warps beats every scheme that gives the owner its own.
// Stages 2-4: [M x M] x [M x 3] DMMA contractions against fragment-ordered operator tables, one 8-row m-tile per warp,
// boundary terms as the last k index.
#pragma once
#include "sri_fused16_dmma.cuh"

namespace sri {

template <int RT, int CT, int W>
struct TiledDmmaCfg {
    static constexpr int QR = 2 * RT * W;   // quaternion row capacity (32 / 64)
    static constexpr int NC = 8 * CT;       // columns incl. the right-hand side (32 / 64)
    static_assert(QR == NC, "square capacity");
    static_assert(W % 4 == 0, "the compile-time local index of the pivot tile needs W to be a multiple of 4");
    static constexpr int RS = NC + 4;       // row stride of the [4][NC] right-hand sides (rows on disjoint banks)
    static constexpr int MT = QR / 8;       // m-tiles of the stage operators
    static constexpr int KT = NC / 4;       // k-tiles
    static constexpr int threads = 32 * W;
    // operator tables (global and shared): Stx [QR][NC] row-major | AS | AT fragment ordered [MT][KT][32]
    static constexpr int tab_doubles = 3 * QR * NC;
    // shared memory layout (doubles)
    static constexpr int stx = 0;
    static constexpr int AS = stx + QR * NC;
    static constexpr int AT = AS + QR * NC;
    static constexpr int kx = AT + QR * NC;           // [2][4][NC]
    static constexpr int ubuf = kx + 2 * 4 * NC;      // [2][CT][32]  published pivot-row fragments
    static constexpr int bs = ubuf + 2 * CT * 32;     // [4][RS]  r' at nodes 0..M-1 | r0 in the last slot
    static constexpr int fbs = bs + 4 * RS;           // [4][RS]  fbar at nodes 1..M | F_tip
    static constexpr int xs = fbs + 4 * RS;           // [4][RS]  r' x n + lbar | M_tip
    static constexpr int rps = xs + 4 * RS;           // [3][NC]  r' at all nodes
    static constexpr int lbs = rps + 3 * NC;          // [3][NC]  lbar at nodes 1..M (slot j = node j+1)
    static constexpr int gam = lbs + 3 * NC;          // [3][NC]
    static constexpr int ns = gam + 3 * NC;           // [3][NC]  n by reduced row
    static constexpr int qnode = ns + 3 * NC;         // [QR][4]
    static constexpr int ints = qnode + QR * 4;       // int thr[2]
    static constexpr int total = ints + 2;
    static constexpr size_t smem_bytes = (size_t)total * sizeof(double);
};

#ifndef SRI_T32_MINBLOCKS
#define SRI_T32_MINBLOCKS 5  // measured: 4 CTAs/SM 2.87e7, 5 (94 registers, no spills) 3.01e7, 6 (spills) 2.86e7 rods/s at N = 32
#endif
template <int RT, int CT, int W>
__global__ void __launch_bounds__(32 * W, (W == 4 ? SRI_T32_MINBLOCKS : 1)) tiled_dmma_kernel(const FusedParams p) {
    using C = TiledDmmaCfg<RT, CT, W>;
    constexpr int NC = C::NC, QR = C::QR, RS = C::RS, KT = C::KT, MT = C::MT;
    if (p.skip && *p.skip) return;  // Newton loop: the solve has already converged (device-side flag), nothing to do
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    if ((long long)blockIdx.x >= p.batch) return;
    for (int i = tid; i < C::tab_doubles; i += C::threads) smem[i] = p.ops2[i];
    for (int i = C::kx + tid; i < C::total; i += C::threads) smem[i] = 0.0;
    __syncthreads();
    const double* stx = smem + C::stx;
    const double* tabAS = smem + C::AS;
    const double* tabAT = smem + C::AT;
    double* ubuf = smem + C::ubuf;
    double* qnode = smem + C::qnode;
    int* thr_s = reinterpret_cast<int*>(smem + C::ints);

    const int M = p.M, N = M + 1;
    // fragment coordinates of this lane (identical to sri_fused16_dmma.cuh)
    const int rho = lane >> 2, cp = lane & 3;
    const int hi = rho >> 2, rr = rho & 3;
    const int srcU_base = 4 * cp + (lane >> 3);
    const bool odd_col = (rho & 1) != 0;
    const int srcL_base = 16 * hi + 4 * (rr ^ cp);
    const unsigned neg_tab = 0x428Eu;
    const unsigned sgL_mask = ((neg_tab >> (4 * rr + cp)) & 1u) << 31;
    const double dpiv0 = (hi == 0 && rr == cp) ? 1.0 : 0.0;
    const double dpiv1 = (hi == 1 && rr == cp) ? 1.0 : 0.0;
    const unsigned hi_mask = hi ? 0x7fffffffu : 0u;
    const int sp = rho >> 1;
    const int idxN = sp ^ cp;
    const unsigned sgN_mask = ((((neg_tab >> (4 * sp + cp)) & 1u) != 0) != (idxN != 0)) ? 0x80000000u : 0u;
    const int offb = (rho < 3 ? rho : 3) * RS + cp;
    const int growth_log = p.growth_log;

    // strain samples and q0 of rod `rod_` -> kx[slot]: row 0 = (0,..,0,q0w), rows 1..3 = (K_c[0..M-1], 0.., q0_c)
    auto prefetch_K = [&](long long rod_, int slot) {
        double* kb = smem + C::kx + slot * 4 * NC;
        if (tid < M) {
            const double* s = p.K + rod_ * 3 * N + tid;
            cp_async8(kb + NC + tid, s); cp_async8(kb + 2 * NC + tid, s + N); cp_async8(kb + 3 * NC + tid, s + 2 * N);
        }
        if (tid < 4) {
            if (p.q0) cp_async8(kb + NC * tid + NC - 1, p.q0 + rod_ * 4 + tid);
            else kb[NC * tid + NC - 1] = (tid == 0) ? 1.0 : 0.0;
        }
    };
    prefetch_K(blockIdx.x, 0);
    cp_async_commit();

    int it = 0;
    for (long long rod = blockIdx.x; rod < p.batch; rod += gridDim.x, ++it) {
        const int cur = it & 1;
        // ---- prefetch: this rod's late inputs (reduced-row shifted, boundary values in the last k slot) and the next
        //      rod's strain samples
        if (tid >= 1 && tid < N) {
            for (int c = 0; c < 3; ++c) {
                if (p.fbar) cp_async8(smem + C::fbs + RS * c + tid - 1, p.fbar + (rod * 3 + c) * N + tid);
                if (p.lbar) cp_async8(smem + C::lbs + NC * c + tid - 1, p.lbar + (rod * 3 + c) * N + tid);
            }
        }
        if (tid < N && p.Gamma)
            for (int c = 0; c < 3; ++c) cp_async8(smem + C::gam + NC * c + tid, p.Gamma + (rod * 3 + c) * N + tid);
        {
            const int t3 = C::threads - 1 - tid;  // the last threads of the CTA carry the three boundary vectors
            if (t3 >= 0 && t3 < 3) { if (p.F_tip) cp_async8(smem + C::fbs + RS * t3 + NC - 1, p.F_tip + rod * 3 + t3); }
            else if (t3 >= 4 && t3 < 7) { if (p.M_tip) cp_async8(smem + C::xs + RS * (t3 - 4) + NC - 1, p.M_tip + rod * 3 + t3 - 4); }
            else if (t3 >= 8 && t3 < 11) { if (p.r0) cp_async8(smem + C::bs + RS * (t3 - 8) + NC - 1, p.r0 + rod * 3 + t3 - 8); }
        }
        if (rod + gridDim.x < p.batch) prefetch_K(rod + gridDim.x, cur ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        const double* kb = smem + C::kx + cur * 4 * NC;
        // ---- stage 1: assemble Cr = delta - 1/2 S_ij (0,K_j) | g_i q0 in C-fragment layout -------------------------
        double c[RT][CT][2];
#pragma unroll
        for (int ct = 0; ct < CT; ++ct) {
            const double2 kq = *reinterpret_cast<const double2*>(kb + NC * rr + 8 * ct + 2 * cp);
#pragma unroll
            for (int tl = 0; tl < RT; ++tl) {
                const int i = 2 * (tl * W + w) + hi;
                const int j = 8 * ct + 2 * cp;
                const double2 s = *reinterpret_cast<const double2*>(stx + i * NC + j);
                c[tl][ct][0] = fma(s.x, kq.x, (rr == 0 && i == j) ? 1.0 : 0.0);
                c[tl][ct][1] = fma(s.y, kq.y, (rr == 0 && i == j + 1 && j + 1 < NC - 1) ? 1.0 : 0.0);
            }
        }

        double la[RT];
        unsigned mx = 0u;
        bool bad = false;

        // pivot row of step kn = 8 KCn + 2 pn + KEn: normalise and publish (owner warp only)
        auto prepare = [&](auto kcn_c, auto ken_c, int pn) {
            constexpr int KCn = decltype(kcn_c)::value, KEn = decltype(ken_c)::value;
            constexpr int tlo = (4 * KCn) / W;
            const int kn = 8 * KCn + 2 * pn + KEn;
            const int srcU = 16 * KEn + srcU_base;
            const double pcs = __shfl_sync(0xffffffffu, c[tlo][KCn][KEn], 16 * KEn + 4 * idxN + pn);
            const double bn0 = flip_sign(pcs, sgN_mask);
            double un0[CT];
#pragma unroll
            for (int ct = KCn; ct < CT; ++ct) {
                const double v0 = __shfl_sync(0xffffffffu, c[tlo][ct][0], srcU);
                const double v1 = __shfl_sync(0xffffffffu, c[tlo][ct][1], srcU);
                un0[ct] = dmma_zero(odd_col ? v1 : v0, bn0);  // U (x) conj(c_kk), B-fragment layout
            }
            const double nu = __shfl_sync(0xffffffffu, un0[KCn], 4 * (2 * pn + KEn));  // |c_kk|^2
            double r0;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(nu));
            const double e = fma(-nu, r0, 1.0);
            const double t2 = fma(e, e, e);
            const double r0n = flip_sign(r0, 0x80000000u);
            double* ub = ubuf + (kn & 1) * CT * 32 + lane;
#pragma unroll
            for (int ct = KCn; ct < CT; ++ct) { const double a = un0[ct] * r0n; ub[ct * 32] = fma(a, t2, a); }
            if (lane == 0) {
                const int hn = __double2hiint(nu);
                int thr = (int)(((long long)hn + 0x3ff00000LL + (long long)growth_log) >> 1);
                if (hn < 0x05000000 || hn >= 0x7ff00000) thr = -1;
                thr_s[kn & 1] = thr;
            }
        };
        // pivot column of step kn -> A fragments of this warp's tiles (+ growth maximum over the rows below the pivot)
        auto gatherL = [&](auto kcn_c, auto ken_c, int pn) {
            constexpr int KCn = decltype(kcn_c)::value, KEn = decltype(ken_c)::value;
            constexpr int tlo = (4 * KCn) / W;
            const int Tn = 4 * KCn + pn;  // row tile of the pivot
            const int srcL = srcL_base + pn;
            mx = 0u;
#pragma unroll
            for (int tl = 0; tl < RT; ++tl) {
                const int T = tl * W + w;
                const double v = __shfl_sync(0xffffffffu, c[tl][KCn][KEn], srcL);
                const unsigned h = (unsigned)__double2hiint(v) & 0x7fffffffu;
                if (T > Tn) mx = max(mx, h);
                else if (T == Tn && KEn == 0) mx = max(mx, h & hi_mask);
                la[tl] = flip_sign(v, sgL_mask);
            }
            if (w == (4 * KCn) % W + pn) la[tlo] -= KEn ? dpiv1 : dpiv0;  // owner warp only: pivot row, Rmat(c_kk - 1)
        };
        auto owner_of = [&](int kc, int pn) { return (4 * kc) % W + pn; };

        // prologue: step 0
        if (w == owner_of(0, 0)) prepare(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{}, 0);
        gatherL(std::integral_constant<int, 0>{}, std::integral_constant<int, 0>{}, 0);
        __syncthreads();

        // ---- Gauss-Jordan: unrolled over the pivot's column tile and parity, run-time loop over (k % 8) / 2 ---------
        bool done = false;
        auto step = [&](auto kc_c, auto ke_c, int pp) {
            constexpr int KC = decltype(kc_c)::value, KE = decltype(ke_c)::value;
            const int k = 8 * KC + 2 * pp + KE;
            // published fragments of the pivot row and the growth threshold
            double un[CT];
            const double* ub = ubuf + (k & 1) * CT * 32 + lane;
#pragma unroll
            for (int ct = KC; ct < CT; ++ct) un[ct] = ub[ct * 32];
            bad = bad || ((int)mx > thr_s[k & 1]);
            // rank-4 update of this warp's tiles
#pragma unroll
            for (int tl = 0; tl < RT; ++tl)
#pragma unroll
                for (int ct = KC; ct < CT; ++ct) dmma(c[tl][ct][0], c[tl][ct][1], la[tl], un[ct]);
            // next step
            if (k + 1 < M) {
                if (KE == 0) {
                    if (w == owner_of(KC, pp)) prepare(kc_c, std::integral_constant<int, 1>{}, pp);
                    gatherL(kc_c, std::integral_constant<int, 1>{}, pp);
                } else if (pp < 3) {
                    if (w == owner_of(KC, pp + 1)) prepare(kc_c, std::integral_constant<int, 0>{}, pp + 1);
                    gatherL(kc_c, std::integral_constant<int, 0>{}, pp + 1);
                } else if (KC + 1 < CT) {
                    constexpr int KN = (KC + 1 < CT) ? KC + 1 : KC;
                    if (w == owner_of(KN, 0)) prepare(std::integral_constant<int, KN>{}, std::integral_constant<int, 0>{}, 0);
                    gatherL(std::integral_constant<int, KN>{}, std::integral_constant<int, 0>{}, 0);
                }
            } else {
                done = true;
            }
            __syncthreads();
        };
        auto column_tile = [&](auto kc_c) {
            constexpr int KC = decltype(kc_c)::value;
            if (done || 8 * KC >= M) return;
#pragma unroll 1
            for (int pp = 0; pp < 4; ++pp) {
                step(kc_c, std::integral_constant<int, 0>{}, pp);
                if (done) break;
                step(kc_c, std::integral_constant<int, 1>{}, pp);
                if (done) break;
            }
        };
        column_tile(std::integral_constant<int, 0>{});
        if (CT > 1) column_tile(std::integral_constant<int, (1 < CT ? 1 : 0)>{});
        if (CT > 2) column_tile(std::integral_constant<int, (2 < CT ? 2 : 0)>{});
        if (CT > 3) column_tile(std::integral_constant<int, (3 < CT ? 3 : 0)>{});
        if (CT > 4) column_tile(std::integral_constant<int, (4 < CT ? 4 : 0)>{});
        if (CT > 5) column_tile(std::integral_constant<int, (5 < CT ? 5 : 0)>{});
        if (CT > 6) column_tile(std::integral_constant<int, (6 < CT ? 6 : 0)>{});
        if (CT > 7) column_tile(std::integral_constant<int, (7 < CT ? 7 : 0)>{});

        const bool flagged = __syncthreads_or(bad ? 1 : 0) != 0;
        cp_async_wait<0>();
        const bool keep = !(flagged && p.rod_list);
        if (!keep && tid == 0) p.rod_list[atomicAdd(p.rod_count, 1)] = (int)rod;
        if (p.info && keep && tid == 0) p.info[rod] = flagged ? -1 : 0;
        // ---- the solution is the last column: lanes cp == 3, register [tl][CT-1][1] -> qnode[i][r] -----------------
        if (cp == 3) {
#pragma unroll
            for (int tl = 0; tl < RT; ++tl) qnode[4 * (2 * (tl * W + w) + hi) + rr] = c[tl][CT - 1][1];
        }
        __syncthreads();
        if (tid < 4) qnode[4 * M + tid] = kb[NC * tid + NC - 1];  // base node: q0
        __syncthreads();
        // ---- pointwise work: thread `tid` = node tid --------------------------------------------------------------
        const int node = tid;
        if (node <= M) {
            const quat q = ld_quat(qnode + 4 * node);
            if (p.Q && keep && node < M) {
                double* d = p.Q + rod * 4 * M + node;
                d[0] = q.w; d[M] = q.x; d[2 * M] = q.y; d[3 * M] = q.z;
            }
            if (p.r || p.n || p.m) {
                double bv0, bv1, bv2;
                if (p.Gamma) {
                    const double* gm = smem + C::gam + node;
                    q_rotate(q, gm[0], gm[NC], gm[2 * NC], bv0, bv1, bv2);
                } else {
                    q_rotate_e1(q, bv0, bv1, bv2);
                }
                double* rp = smem + C::rps + node;
                rp[0] = bv0; rp[NC] = bv1; rp[2 * NC] = bv2;
                if (node < M) { double* b = smem + C::bs + node; b[0] = bv0; b[RS] = bv1; b[2 * RS] = bv2; }
            }
        }
        if (p.r || p.n || p.m) {
            __syncthreads();
            // m-tiles of the stage operators are dealt to the warps: mt = w, w + W, ...
            constexpr int MPW = (MT + W - 1) / W;
            double acc0[MPW], acc1[MPW];
            auto contract = [&](const double* at, const double* rhs, int kt0) {
                const double* b = rhs + offb;
#pragma unroll
                for (int q = 0; q < MPW; ++q) {
                    acc0[q] = 0.0; acc1[q] = 0.0;
                    const int mt = w + q * W;
                    if (mt < MT) {
                        const double* a = at + (mt * KT) * 32 + lane;
#pragma unroll 4
                        for (int kt = kt0; kt < KT; ++kt) dmma(acc0[q], acc1[q], a[kt * 32], b[4 * kt]);
                    }
                }
            };
            auto store = [&](double* out) {
#pragma unroll
                for (int q = 0; q < MPW; ++q) {
                    const int i = 8 * (w + q * W) + rho;
                    if (i < M) {
                        if (cp == 0) { out[i] = acc0[q]; out[M + i] = acc1[q]; }
                        else if (cp == 1) out[2 * M + i] = acc0[q];
                    }
                }
            };
            if (p.r) {
                contract(tabAS, smem + C::bs, 0);
                if (keep) store(p.r + rod * 3 * M);
            }
            if (p.n || p.m) {
                contract(tabAT, smem + C::fbs, p.fbar ? 0 : KT - 1);
                if (p.n && keep) store(p.n + rod * 3 * M);
                if (p.m) {
                    double* nsv = smem + C::ns;
#pragma unroll
                    for (int q = 0; q < MPW; ++q) {
                        const int i = 8 * (w + q * W) + rho;
                        if (i < QR) {
                            if (cp == 0) { nsv[i] = acc0[q]; nsv[NC + i] = acc1[q]; }
                            else if (cp == 1) nsv[2 * NC + i] = acc0[q];
                        }
                    }
                }
            }
            if (p.m) {
                __syncthreads();
                if (node < M) {
                    const double* nsv = smem + C::ns + node;
                    const double n0 = nsv[0], n1 = nsv[NC], n2 = nsv[2 * NC];
                    const double* rp = smem + C::rps + node + 1;  // node of reduced row `node`
                    const double rp0 = rp[0], rp1 = rp[NC], rp2 = rp[2 * NC];
                    double l0 = 0.0, l1 = 0.0, l2 = 0.0;
                    if (p.lbar) { const double* lb = smem + C::lbs + node; l0 = lb[0]; l1 = lb[NC]; l2 = lb[2 * NC]; }
                    double* x = smem + C::xs + node;
                    x[0] = fma(rp1, n2, fma(-rp2, n1, l0));
                    x[RS] = fma(rp2, n0, fma(-rp0, n2, l1));
                    x[2 * RS] = fma(rp0, n1, fma(-rp1, n0, l2));
                }
                __syncthreads();
                contract(tabAT, smem + C::xs, 0);
                if (keep) store(p.m + rod * 3 * M);
            }
        }
        __syncthreads();  // scratch is reused by the next rod
    }
    cp_async_wait<0>();
}

}  // namespace sri
