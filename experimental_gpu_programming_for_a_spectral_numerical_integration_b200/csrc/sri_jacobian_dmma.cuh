// sri_jacobian_dmma.cuh -- analytic Jacobian of the Galerkin shape residual (sri_shape_jacobian) on the FP64 tensor
// cores, N <= 16, one rod per warp.
//
// Math (restated on the CPU in oracle/tangent.py): for a direction d = (c, k), dK = P_k(t) e_c,
//     u_i   = P_k(t_i) R_i[:, c]                          i = 0..M-1 (the base node does not rotate)
//     dth   = S u                                         S   = Dn_NN^-1   (nodes 0..M-1; 0 at the base)
//     v_j   = -((dth_j x b_j) x n_j) = dth_j (b_j.n_j) - b_j (dth_j.n_j)     nodes j = 1..N-1, b = R Gamma
//     dm    = S_T v                                       S_T = D_TT^-1    (nodes 1..N-1; 0 at the tip)
//     drho_i= H dK_i - R_i^T (dm_i - dth_i x m_i)         all N nodes
//     J[(c',k')][d] = sum_i w_i P_k'(t_i) drho_i[c'].
// All 3 ne directions of a rod go through the two integration matrices at once: two [16 x 16] x [16 x 9 ne]
// contractions and one [ne x 16] x [16 x 9 ne] projection = (2 x 2 + 1) x 4 x NT DMMA m8n8k4, NT = ceil(9 ne / 8)
// (80 DMMAs per rod at ne = 3), where the scalar kernel it replaces (shape_jacobian_kernel, still used for N > 16) spent
// 2 x 15 x 45 x 9 FMAs per rod in one-lane dot products.  The pointwise work between the contractions (v, drho) is done
// once per (node, direction) item -- 16 x 3 ne items dealt to the 32 lanes -- and staged through the two field arrays of the
// warp's scratch, from which the lanes then read their B fragments with plain loads (evaluating every B fragment in the
// lane that owns it cost 3 x the FP64 work and 1.7 x the shared-memory loads and measured 447 us per 10^5 rods).
//
// Fragment layout of mma.sync.m8n8k4.f64 (rho = lane / 4, cp = lane % 4): A[row rho][k cp], B[k cp][col rho],
// C[row rho][cols 2 cp, 2 cp + 1].
#pragma once
#include "sri_device.cuh"
#include "sri_fused16.cuh"     // cp_async8
#include "sri_stage_dmma.cuh"  // dmma_m8n8k4, StageTables

namespace sri {

constexpr int kJacWarps = 4;
#ifndef SRI_JAC_MINBLOCKS
#define SRI_JAC_MINBLOCKS 2  // measured per 10^5 rods: 2 CTAs per SM (255 registers, no spills) 245 us, 3 (168 registers) 289 us, 4 (128) 315 us
#endif

template <int NE>
struct JacDmmaScratch {  // doubles per warp
    static constexpr int NT = (9 * NE + 7) / 8;       // column tiles of the 9 NE = 3 (components) x 3 NE (directions) columns
    static constexpr int LD = 8 * NT + 2;             // row stride of the two field arrays: = 2 mod 16 when NT is even, which keeps the
                                                      // pointwise passes (lane = (node, direction parity)) free of bank conflicts
    static constexpr int R = 0;                       // [16][9]  rotation matrices by node, row-major
    static constexpr int b = R + 144;                 // [16][3]  R Gamma
    static constexpr int n = b + 48;                  // [16][3]  internal force by node (node 0 unused)
    static constexpr int m = n + 48;                  // [16][3]  internal couple by node (node 0 = M_tip)
    static constexpr int bn = m + 48;                 // [16]     b . n
    static constexpr int th = bn + 16;                // [16][LD] dtheta by node, column 3 d + comp
    static constexpr int dm = th + 16 * LD;           // [17][LD] dm by node (row 0 = tip = 0, row 16 = padding)
    static constexpr int pre = dm + 17 * LD;          // [2][7][32] raw nodal inputs of this and the next rod (cp.async staging)
    static constexpr int total = pre + 2 * 7 * 32;
};

template <int NE>
__global__ void __launch_bounds__(32 * kJacWarps, NE <= 4 ? SRI_JAC_MINBLOCKS : 1) shape_jacobian_dmma_kernel(
    long long batch, int N, const double* __restrict__ ops, const double* __restrict__ ptab, const double* __restrict__ ccw,
    double h0, double h1, double h2, const double* __restrict__ Q, const double* __restrict__ q0,
    const double* __restrict__ Gamma, const double* __restrict__ nin, const double* __restrict__ min_,
    const double* __restrict__ M_tip, double* __restrict__ J, const int* __restrict__ skip) {
    if (skip && *skip) return;
    using SC = JacDmmaScratch<NE>;
    constexpr int NT = SC::NT, LD = SC::LD, ND = 3 * NE;  // ND directions, 3 ND columns
    constexpr int ITEMS = (ND + 1) / 2;                    // pointwise passes: lane = (node, parity), directions d = parity + 2 r
    extern __shared__ __align__(16) double jsm[];
    double* Ps = jsm;            // [8][16] Legendre table, zero beyond node N-1
    double* Pw = jsm + 128;      // [8][16] w_i P_k(t_i)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* scr = jsm + 256 + warp * SC::total;
    const int M = N - 1;
    const int rho = lane >> 2, cp = lane & 3;

    for (int e = threadIdx.x; e < 128; e += blockDim.x) {
        const int k = e >> 4, i = e & 15;
        const double pv = i < N ? ptab[k * N + i] : 0.0;
        Ps[e] = pv;
        Pw[e] = i < N ? ccw[i] * pv : 0.0;
    }
    for (int e = lane; e < SC::total; e += 32) scr[e] = 0.0;
    __syncthreads();

    // A fragments, loaded once: the two integration matrices (zero padded 16 x 16, row-major) and the projection
    double aS[2][4], aT[2][4], aP[4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            aS[mt][kt] = ops[StageTables::Srm + (8 * mt + rho) * 16 + 4 * kt + cp];
            aT[mt][kt] = ops[StageTables::STrm + (8 * mt + rho) * 16 + 4 * kt + cp];
        }
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) aP[kt] = rho < NE ? Pw[rho * 16 + 4 * kt + cp] : 0.0;

    // per column tile: what this lane's B column (8 nt + rho) means
    //   contractions: column = 3 d + comp          (projection: column = c' ND + d)
    int colD[NT], colComp[NT];
    bool colOk[NT];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int col = 8 * nt + rho;
        colOk[nt] = col < 3 * ND;
        const int cc = colOk[nt] ? col : 0;
        colD[nt] = cc / 3; colComp[nt] = cc - 3 * colD[nt];
    }

    // rod-independent factors in registers: P_k(t_j) of this lane's B entries in the first contraction (0 where the entry is
    // structurally zero), and P_k(t_node) of the node this lane serves in the pointwise passes
    double pk1[4][NT];
#pragma unroll
    for (int kt = 0; kt < 4; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int j = 4 * kt + cp, d = colD[nt], k = d - NE * (d / NE);
            pk1[kt][nt] = (colOk[nt] && j < M) ? Ps[k * 16 + j] : 0.0;
        }
    const int pnode = lane >> 1, dpar = lane & 1;
    double pkn[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) pkn[k] = Ps[k * 16 + pnode];

    double* Rs = scr + SC::R; double* bs = scr + SC::b; double* ns = scr + SC::n; double* ms = scr + SC::m;
    double* th = scr + SC::th; double* dmv = scr + SC::dm;

    const long long warps_total = (long long)gridDim.x * kJacWarps;
    // This lane's raw nodal inputs of one rod -- lanes 0..15 the quaternion (and Gamma) of node `lane`, lanes 16..31 the couple
    // and force of node `lane - 16` -- are copied one rod ahead into the staging slots with cp.async, so that the global-memory
    // latency hides behind the previous rod without holding registers.  Slot layout [7][32]: entry e of lane l at e * 32 + l.
    auto prefetch = [&](long long rod_, int slot) {
        double* st = scr + SC::pre + slot * 224 + lane;
        if (lane < 16) {
            const int i = lane;
            if (i < M) { const double* q = Q + rod_ * 4 * M + i; cp_async8(st, q); cp_async8(st + 32, q + M); cp_async8(st + 64, q + 2 * M); cp_async8(st + 96, q + 3 * M); }
            else if (i == M && q0) { const double* q = q0 + rod_ * 4; cp_async8(st, q); cp_async8(st + 32, q + 1); cp_async8(st + 64, q + 2); cp_async8(st + 96, q + 3); }
            if (Gamma && i < N) { const double* gm = Gamma + rod_ * 3 * N + i; cp_async8(st + 128, gm); cp_async8(st + 160, gm + N); cp_async8(st + 192, gm + 2 * N); }
        } else {
            const int i = lane - 16;
            if (i == 0) { const double* t = M_tip + rod_ * 3; cp_async8(st, t); cp_async8(st + 32, t + 1); cp_async8(st + 64, t + 2); }
            else if (i < N) {
                const double* t = min_ + rod_ * 3 * M + (i - 1); cp_async8(st, t); cp_async8(st + 32, t + M); cp_async8(st + 64, t + 2 * M);
                const double* f = nin + rod_ * 3 * M + (i - 1); cp_async8(st + 96, f); cp_async8(st + 128, f + M); cp_async8(st + 160, f + 2 * M);
            }
        }
    };
    const long long rod_first = (long long)blockIdx.x * kJacWarps + warp;
    if (rod_first < batch) prefetch(rod_first, 0);
    cp_async_commit();
    int it = 0;
    for (long long rod = rod_first; rod < batch; rod += warps_total, ++it) {
        if (rod + warps_total < batch) prefetch(rod + warps_total, (it + 1) & 1);
        cp_async_commit();
        cp_async_wait<1>();  // everything but the group just committed: this rod's inputs have landed
        __syncwarp();
        double cur[7];
        {
            const double* st = scr + SC::pre + (it & 1) * 224 + lane;
#pragma unroll
            for (int e = 0; e < 7; ++e) cur[e] = st[32 * e];
            if (lane < 16) {
                if (!(lane < M || (lane == M && q0))) { cur[0] = 1.0; cur[1] = 0.0; cur[2] = 0.0; cur[3] = 0.0; }  // identity (base node without q0)
                if (!Gamma) { cur[4] = 1.0; cur[5] = 0.0; cur[6] = 0.0; }
            }
        }
        // ---- nodal data -> the warp's scratch ---------------------------------------------------------------------------------
        if (lane < 16) {
            const int i = lane;
            if (i < N) {
                double Rl[9];
                {
                    const double qw = cur[0], qx = cur[1], qy = cur[2], qz = cur[3];
                    const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
                    const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
                    const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
                    const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
                    Rl[0] = 1 - (tyy + tzz); Rl[1] = txy - twz;       Rl[2] = txz + twy;
                    Rl[3] = txy + twz;       Rl[4] = 1 - (txx + tzz); Rl[5] = tyz - twx;
                    Rl[6] = txz - twy;       Rl[7] = tyz + twx;       Rl[8] = 1 - (txx + tyy);
                }
#pragma unroll
                for (int e = 0; e < 9; ++e) Rs[9 * i + e] = Rl[e];
                const double g0 = cur[4], g1 = cur[5], g2 = cur[6];
                bs[3 * i] = Rl[0] * g0 + Rl[1] * g1 + Rl[2] * g2;
                bs[3 * i + 1] = Rl[3] * g0 + Rl[4] * g1 + Rl[5] * g2;
                bs[3 * i + 2] = Rl[6] * g0 + Rl[7] * g1 + Rl[8] * g2;
            }
        } else {
            const int i = lane - 16;
            if (i < N) {
                ms[3 * i] = cur[0]; ms[3 * i + 1] = cur[1]; ms[3 * i + 2] = cur[2];
                if (i >= 1) { ns[3 * i] = cur[3]; ns[3 * i + 1] = cur[4]; ns[3 * i + 2] = cur[5]; }
            }
        }
        __syncwarp();

        __syncwarp();

        double acc[2][NT][2];
        // ---- dtheta = S u,  u[node j][3 d + comp] = P_k(t_j) R_j[comp][c] ----------------------------------------------------
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            const int j = 4 * kt + cp;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double bv = pk1[kt][nt] * Rs[9 * j + 3 * colComp[nt] + colD[nt] / NE];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], aS[mt][kt], bv);
            }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                *reinterpret_cast<double2*>(th + (8 * mt + rho) * LD + 8 * nt + 2 * cp) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        __syncwarp();
        // ---- v = dth (b.n) - b (dth.n) at nodes 1..M: lane (node, parity) loads its node's b and n once and walks over its
        //      directions; staged in the rows of `dmv` (row = node) that the contraction below then overwrites with dm ------------
        if (pnode >= 1 && pnode <= M) {
            const double n0 = ns[3 * pnode], n1 = ns[3 * pnode + 1], n2 = ns[3 * pnode + 2];
            const double b0 = bs[3 * pnode], b1 = bs[3 * pnode + 1], b2 = bs[3 * pnode + 2];
            const double bdn = b0 * n0 + b1 * n1 + b2 * n2;
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                const int d = dpar + 2 * r;
                if (d < ND) {
                    const double* a = th + pnode * LD + 3 * d;
                    const double a0 = a[0], a1 = a[1], a2 = a[2];
                    const double adn = a0 * n0 + a1 * n1 + a2 * n2;
                    double* vd = dmv + pnode * LD + 3 * d;
                    vd[0] = a0 * bdn - b0 * adn;
                    vd[1] = a1 * bdn - b1 * adn;
                    vd[2] = a2 * bdn - b2 * adn;
                }
            }
        }
        __syncwarp();
        // ---- dm = S_T v ------------------------------------------------------------------------------------------------------
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            const double* vrow = dmv + (4 * kt + cp + 1) * LD + rho;   // B[k = reduced row 4 kt + cp][column 8 nt + rho]
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double bv = vrow[8 * nt];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], aT[mt][kt], bv);
            }
        }
        __syncwarp();  // every lane has read its v before the rows are overwritten
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)   // reduced row 8 mt + rho is node 8 mt + rho + 1 (row 0 of dmv, the tip, stays zero)
                *reinterpret_cast<double2*>(dmv + (8 * mt + rho + 1) * LD + 8 * nt + 2 * cp) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
        __syncwarp();
        // ---- drho_i[c'][d] = H_c' P_k(t_i) [c' == c] - sum_r R_i[r][c'] (dm_i - dth_i x m_i)_r: lane (node, parity) loads its node's
        //      m and R once; results stay in registers until every lane has read dth, then go to `th` with column c' ND + d -------
        double dr[ITEMS][3];
        {
            const bool live = pnode < N;
            const int nd = live ? pnode : 0;
            const double m0 = ms[3 * nd], m1 = ms[3 * nd + 1], m2 = ms[3 * nd + 2];
            double Rl[9];
#pragma unroll
            for (int e = 0; e < 9; ++e) Rl[e] = Rs[9 * nd + e];
#pragma unroll
            for (int r = 0; r < ITEMS; ++r) {
                const int d = dpar + 2 * r;
                dr[r][0] = 0.0; dr[r][1] = 0.0; dr[r][2] = 0.0;
                if (d < ND && live) {
                    const int c = d / NE, k = d - NE * c;
                    const double* a = th + nd * LD + 3 * d;
                    const double* g = dmv + nd * LD + 3 * d;
                    const double a0 = a[0], a1 = a[1], a2 = a[2];
                    const double w0 = g[0] - (a1 * m2 - a2 * m1);
                    const double w1 = g[1] - (a2 * m0 - a0 * m2);
                    const double w2 = g[2] - (a0 * m1 - a1 * m0);
                    double pv = pkn[0];
#pragma unroll
                    for (int kk = 1; kk < NE; ++kk) pv = (k == kk) ? pkn[kk] : pv;
                    const double hk = (c == 0 ? h0 : (c == 1 ? h1 : h2)) * pv;
                    dr[r][0] = (c == 0 ? hk : 0.0) - (Rl[0] * w0 + Rl[3] * w1 + Rl[6] * w2);
                    dr[r][1] = (c == 1 ? hk : 0.0) - (Rl[1] * w0 + Rl[4] * w1 + Rl[7] * w2);
                    dr[r][2] = (c == 2 ? hk : 0.0) - (Rl[2] * w0 + Rl[5] * w1 + Rl[8] * w2);
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const int d = dpar + 2 * r;
            if (d < ND) { double* o = th + pnode * LD + d; o[0] = dr[r][0]; o[ND] = dr[r][1]; o[2 * ND] = dr[r][2]; }
        }
        __syncwarp();
        // ---- projection: J[(c',k')][d] = sum_i (w_i P_k'(t_i)) drho_i[c'][d] -------------------------------------------------------
        double pj[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { pj[nt][0] = 0.0; pj[nt][1] = 0.0; }
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            const double* drow = th + (4 * kt + cp) * LD + rho;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) dmma_m8n8k4(pj[nt][0], pj[nt][1], aP[kt], drow[8 * nt]);
        }
        // C fragment: row rho = k', columns 8 nt + 2 cp (+1) = c' ND + d  ->  J[(c' NE + k')][d], row-major ND x ND per rod
        if (rho < NE) {
            double* Jr = J + rod * ND * ND;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * nt + 2 * cp + e;
                    if (col < 3 * ND) {
                        const int cq = col / ND, d = col - ND * cq;
                        Jr[(cq * NE + rho) * ND + d] = pj[nt][e];
                    }
                }
        }
        __syncwarp();  // the scratch is rewritten by the next rod
    }
    cp_async_wait<0>();
}

}  // namespace sri
