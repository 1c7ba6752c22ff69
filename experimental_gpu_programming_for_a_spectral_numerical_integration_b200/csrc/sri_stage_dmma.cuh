// sri_stage_dmma.cuh -- the strain-independent stages as batched FP64 tensor-core contractions (N <= 16).
//
// Position (main.cpp:121-176), internal force and internal couple (rod_modeling.pdf eqs. 1.17-1.18) are all
//      Out[i, (rod,c)] = sum_j T[i,j] * Rhs[j, (rod,c)],   T = Dn_NN^-1 or D_TT^-1 (cached, 15 x 15 padded to 16 x 16),
// i.e. one small constant matrix applied to a very wide right-hand side: 2.7 kflop against ~0.8 KB per rod, so the
// roofline is HBM.  Mapping onto DMMA (mma.sync.m8n8k4.f64), one warp per tile of 8 rods:
//   * A fragments: T as 2 (m) x 4 (k) tiles of 8 x 4 -> 8 doubles per lane, loaded once per kernel;
//   * B fragments: lane l owns rod l/4 of the tile and the four nodes 4*kt + l%4.  It loads exactly the inputs of
//     those (rod, node) pairs (32-byte segments per rod, every byte of the tile read once, by one lane), evaluates the
//     pointwise right-hand side (R(q)Gamma, -fbar - D_TI F, -(r' x n + lbar) - D_TI M) in registers, and the three
//     components ARE its B fragments -- nothing is staged through shared memory;
//   * C fragments: lane l holds output nodes 8*mt + l/4 of rods 2*(l%4), 2*(l%4)+1 -> 64-byte segments per rod on the
//     way out.
// 24 DMMA per 8 rods; the kernel keeps ~4 KB of loads in flight per warp and many warps per SM to cover HBM latency.
#pragma once
#include "sri_device.cuh"
#include "sri_fused16.cuh"  // FusedParams, OpsLayout16

namespace sri {

enum StageKind { kStagePosition = 0, kStageStress = 1, kStageCouple = 2 };

// Row-major 16 x 16 zero-padded copies of the two cached inverses, appended to the OpsLayout16 table.
struct StageTables {
    static constexpr int Srm = OpsLayout16::total;        // [16][16]  Dn_NN^-1
    static constexpr int STrm = Srm + 256;                // [16][16]  D_TT^-1
    static constexpr int STsh = STrm + 256;               // [16][16]  D_TT^-1 with columns shifted by one: STsh[i][node] =
                                                          //           D_TT^-1[i][node-1], column 0 zero (k index = node)
    static constexpr int DTIsh = STsh + 256;              // [16]      D_TI shifted the same way (DTIsh[node] = D_TI[node-1])
    static constexpr int total = DTIsh + 16;
};

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 4 CTAs per SM (<= 128 registers): measured best on B200 -- with fewer registers (5-6 CTAs) the loads of a tile no longer
// all stay in flight (position -16..-29 %), with the compiler's free choice (96..160 registers) stress is 15 % slower.
#ifndef SRI_STAGE_MINBLOCKS
#define SRI_STAGE_MINBLOCKS 4
#endif
#define SRI_STAGE_BOUNDS __launch_bounds__(128, SRI_STAGE_MINBLOCKS)
template <int STAGE>
__global__ void SRI_STAGE_BOUNDS stage_dmma_kernel(const FusedParams p) {
    const int lane = threadIdx.x & 31;
    const int lr = lane >> 2, lk = lane & 3;  // tile-local rod (B/C row group) and k offset
    const int M = p.M, N = p.N;
    // Stress and couple contract over the reduced rows j' = node - 1.  Using the NODE as the k index (shifted table,
    // zero column for node 0) makes this lane's four k values the nodes 4*kt + l%4, so its loads of the nodal inputs
    // (fbar, lbar, Q, Gamma) are 32-byte aligned segments instead of segments straddling two sectors.
    const double* T = p.ops + (STAGE == kStagePosition ? StageTables::Srm : StageTables::STsh);

    // A fragments: a[mt][kt] = T[8*mt + lane/4][4*kt + lane%4]
    double a[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) a[mt][kt] = T[(8 * mt + lr) * 16 + 4 * kt + lk];
    double dti[4], gvec[2];
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) dti[kt] = p.ops[StageTables::DTIsh + 4 * kt + lk];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) gvec[mt] = p.ops[OpsLayout16::g + 8 * mt + lr];

    const long long tiles = (p.batch + 7) >> 3;
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long tile = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < tiles; tile += warps_total) {
        const long long rod = tile * 8 + lr;
        const bool live = rod < p.batch;

        // ---- pointwise right-hand side of this lane's four (rod, node) pairs = B fragments -----------------
        double bf[3][4];
        double w0 = 0.0, w1 = 0.0, w2 = 0.0;  // per-rod tip wrench (stress: F_tip, couple: M_tip)
        if (STAGE == kStageStress && live) { const double* s = p.F_tip + rod * 3; w0 = s[0]; w1 = s[1]; w2 = s[2]; }
        if (STAGE == kStageCouple && live) { const double* s = p.M_tip + rod * 3; w0 = s[0]; w1 = s[1]; w2 = s[2]; }
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            const int node = 4 * kt + lk;  // k index of this lane = node; stress / couple use reduced row node - 1
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
            const bool valid = (STAGE == kStagePosition) ? (node < M) : (node >= 1 && node <= M);
            if (live && valid) {
                if (STAGE == kStagePosition || STAGE == kStageCouple) {
                    quat q; q.w = 1.0; q.x = 0.0; q.y = 0.0; q.z = 0.0;
                    if (node < M) {
                        const double* s = p.Qin + rod * 4 * M + node;
                        q.w = s[0]; q.x = s[M]; q.y = s[2 * M]; q.z = s[3 * M];
                    } else if (p.q0) {
                        const double* s = p.q0 + rod * 4;
                        q.w = s[0]; q.x = s[1]; q.y = s[2]; q.z = s[3];
                    }
                    double b0, b1, b2;
                    if (p.Gamma) { const double* gm = p.Gamma + rod * 3 * N + node; q_rotate(q, gm[0], gm[N], gm[2 * N], b0, b1, b2); }
                    else q_rotate_e1(q, b0, b1, b2);
                    if (STAGE == kStagePosition) { r0 = b0; r1 = b1; r2 = b2; }
                    else {
                        const double* s = p.nin + rod * 3 * M + node - 1;
                        const double n0 = s[0], n1 = s[M], n2 = s[2 * M];
                        double l0 = 0.0, l1 = 0.0, l2 = 0.0;
                        if (p.lbar) { const double* lb = p.lbar + rod * 3 * N + node; l0 = lb[0]; l1 = lb[N]; l2 = lb[2 * N]; }
                        r0 = -((b1 * n2 - b2 * n1) + l0) - dti[kt] * w0;
                        r1 = -((b2 * n0 - b0 * n2) + l1) - dti[kt] * w1;
                        r2 = -((b0 * n1 - b1 * n0) + l2) - dti[kt] * w2;
                    }
                } else {  // stress
                    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
                    if (p.fbar) { const double* s = p.fbar + rod * 3 * N + node; f0 = s[0]; f1 = s[N]; f2 = s[2 * N]; }
                    r0 = -f0 - dti[kt] * w0; r1 = -f1 - dti[kt] * w1; r2 = -f2 - dti[kt] * w2;
                }
            }
            bf[0][kt] = r0; bf[1][kt] = r1; bf[2][kt] = r2;
        }

        // ---- Out = T * Rhs on the FP64 tensor cores ---------------------------------------------------------
        double acc[3][2][2];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                acc[c][mt][0] = 0.0; acc[c][mt][1] = 0.0;
#pragma unroll
                for (int kt = 0; kt < 4; ++kt) dmma_m8n8k4(acc[c][mt][0], acc[c][mt][1], a[mt][kt], bf[c][kt]);
            }

        // ---- epilogue: C fragment (node 8*mt + lane/4, rods 2*(lane%4) + {0,1}) -> [rod][c][node] ---------------
        double* out = (STAGE == kStagePosition) ? p.r : (STAGE == kStageStress ? p.n : p.m);
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const long long orod = tile * 8 + 2 * lk + w;
            if (orod >= p.batch) continue;
            double e0 = 0.0, e1 = 0.0, e2 = 0.0;
            if (STAGE == kStagePosition && p.r0) { const double* s = p.r0 + orod * 3; e0 = s[0]; e1 = s[1]; e2 = s[2]; }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int i = 8 * mt + lr;
                if (i >= M) continue;
                double* d = out + orod * 3 * M + i;
                d[0] = fma(gvec[mt], e0, acc[0][mt][w]);
                d[M] = fma(gvec[mt], e1, acc[1][mt][w]);
                d[2 * M] = fma(gvec[mt], e2, acc[2][mt][w]);
            }
        }
    }
}

// n_i = gT_i F_tip when there is no distributed load: pure streaming write, no contraction needed (any N <= 64).
// A warp produces one 2 KB tile of the [batch][3][M] output per pass as four 128-bit streaming stores per lane, each store
// instruction covering 512 contiguous bytes (whole sectors); one division per lane and tile locates the (rod, component) row,
// the other three positions follow by adding 64 = q64 M + r64.  (First version: 64 contiguous bytes per THREAD -- every store
// instruction then wrote half sectors -- 3.5 TB/s against the 7.3 TB/s a device memset of the same bytes reaches.)
__global__ void __launch_bounds__(256) stress_noload_kernel(long long total, int M, const double* __restrict__ gT,
                                                            const double* __restrict__ F_tip, double* __restrict__ n, int aligned16) {
    __shared__ double g[64];
    if (threadIdx.x < M) g[threadIdx.x] = gT[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int q64 = 64 / M, r64 = 64 - q64 * M;
    const long long tiles = (total + 255) >> 8;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const bool small = total <= 0x7fffffffLL;  // 32-bit index arithmetic covers 2^31 output doubles
    for (long long t = wid; t < tiles; t += nwarps) {
        const long long idx0 = (t << 8) + 2 * lane;
        long long u;
        int i;
        if (small) { const unsigned q = (unsigned)idx0 / (unsigned)M; u = q; i = (int)((unsigned)idx0 - q * (unsigned)M); }
        else { u = idx0 / M; i = (int)(idx0 - u * M); }
        double2 v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long idx = idx0 + 64 * e;
            int i1 = i + 1;
            long long u1 = u;
            if (i1 == M) { i1 = 0; ++u1; }
            const double f0 = idx < total ? __ldg(F_tip + u) : 0.0;
            const double f1 = idx + 1 < total ? __ldg(F_tip + u1) : 0.0;
            v[e] = make_double2(g[i] * f0, g[i1] * f1);
            u += q64; i += r64;
            if (i >= M) { i -= M; ++u; }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long idx = idx0 + 64 * e;
            if (idx + 1 < total && aligned16) __stcs(reinterpret_cast<double2*>(n + idx), v[e]);
            else { if (idx < total) n[idx] = v[e].x; if (idx + 1 < total) n[idx + 1] = v[e].y; }
        }
    }
}

}  // namespace sri
