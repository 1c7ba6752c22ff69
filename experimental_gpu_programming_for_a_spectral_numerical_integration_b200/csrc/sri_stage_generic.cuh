// sri_stage_generic.cuh -- the strain-independent stages for 17 <= N <= 64 as streaming DMMA contractions.
//
// Position (main.cpp:121-176), internal force and internal couple (rod_modeling.pdf eqs. 1.17-1.18) with the cached
// [M x M] operators: Out[i, (rod,c)] = sum_j At[i,j] Rhs[j, (rod,c)].  At N = 32 the stage is HBM-bound (6 kflop against
// 1.8 KB per rod), at N = 64 the FP64 tensor pipe and HBM are balanced (24 kflop against 3.5 KB: BASELINE configs[3]'s
// "DMMA-bound position/stress stages").  Same mapping as sri_stage_dmma.cuh -- one warp per tile of 8 rods, lane l owns
// rod l/4 and the k indices 4 kt + l%4, the pointwise right-hand side is evaluated in registers and IS the B fragment --
// with two differences: the A fragments (R/8 x R/4 of them) do not fit in registers and are streamed from the
// fragment-ordered tables in shared memory (one LDS.64 feeds three DMMAs), and the operators are the ones of the fused
// DMMA kernel (TiledDmmaCfg tables AS / AT: boundary term as the last k index, AT = -D_TT^-1), so the right-hand sides
// need no boundary arithmetic: k index j < M carries node j (position) or node j+1 (force, couple), k index R-1 carries
// r0 / F_tip / M_tip.
#pragma once
#include "sri_stage_dmma.cuh"

namespace sri {

template <int STAGE, int R>
__global__ void __launch_bounds__(128) stage_generic_kernel(const FusedParams p) {
    constexpr int KT = R / 4;
    extern __shared__ __align__(16) double gsm[];  // R*R doubles: AS (position) or AT (force, couple)
    const int lane = threadIdx.x & 31;
    const int lr = lane >> 2, lk = lane & 3;
    const int M = p.M, N = p.N;
    {
        const double* src = p.ops2 + (size_t)R * R * (STAGE == kStagePosition ? 1 : 2);
        for (int i = threadIdx.x; i < R * R; i += blockDim.x) gsm[i] = src[i];
    }
    __syncthreads();
    // k tiles that hold nodes (none for the force stage without a distributed load); the boundary term sits in tile KT-1
    const int kt_used = (STAGE == kStageStress && !p.fbar) ? 0 : (M + 3) >> 2;
    const int mt_used = (M + 7) >> 3;

    const long long tiles = (p.batch + 7) >> 3;
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long tile = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < tiles; tile += warps_total) {
        const long long rod = tile * 8 + lr;
        const bool live = rod < p.batch;
        // ---- B fragments: this lane's (rod, k index) right-hand sides, 3 components -------------------------------
        double bf[3][KT];
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            const int j = 4 * kt + lk;
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
            if (live && (kt < kt_used || kt == KT - 1)) {
                if (j < M) {
                    if (STAGE == kStageStress) {
                        if (p.fbar) { const double* s = p.fbar + rod * 3 * N + j + 1; r0 = s[0]; r1 = s[N]; r2 = s[2 * N]; }
                    } else {
                        const int node = (STAGE == kStagePosition) ? j : j + 1;
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
                            const double* s = p.nin + rod * 3 * M + j;
                            const double n0 = s[0], n1 = s[M], n2 = s[2 * M];
                            double l0 = 0.0, l1 = 0.0, l2 = 0.0;
                            if (p.lbar) { const double* lb = p.lbar + rod * 3 * N + node; l0 = lb[0]; l1 = lb[N]; l2 = lb[2 * N]; }
                            r0 = fma(b1, n2, fma(-b2, n1, l0));
                            r1 = fma(b2, n0, fma(-b0, n2, l1));
                            r2 = fma(b0, n1, fma(-b1, n0, l2));
                        }
                    }
                } else if (j == R - 1) {
                    const double* s = (STAGE == kStagePosition) ? p.r0 : (STAGE == kStageStress ? p.F_tip : p.M_tip);
                    if (s) { r0 = s[rod * 3]; r1 = s[rod * 3 + 1]; r2 = s[rod * 3 + 2]; }
                }
            }
            bf[0][kt] = r0; bf[1][kt] = r1; bf[2][kt] = r2;
        }
        // ---- Out = At * Rhs, one 8-row m-tile at a time; C fragment: node 8 mt + lane/4, rods 2 (lane%4) + {0,1} ---
        double* out = (STAGE == kStagePosition) ? p.r : (STAGE == kStageStress ? p.n : p.m);
#pragma unroll 1
        for (int mt = 0; mt < mt_used; ++mt) {
            double acc[3][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
            const double* a = gsm + (size_t)(mt * KT) * 32 + lane;
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                if (kt < kt_used || kt == KT - 1) {
                    const double av = a[kt * 32];
#pragma unroll
                    for (int c = 0; c < 3; ++c) dmma_m8n8k4(acc[c][0], acc[c][1], av, bf[c][kt]);
                }
            }
            const int i = 8 * mt + lr;
            if (i < M) {
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const long long orod = tile * 8 + 2 * lk + w;
                    if (orod < p.batch) {
                        double* d = out + orod * 3 * M + i;
                        d[0] = acc[0][w]; d[M] = acc[1][w]; d[2 * M] = acc[2][w];
                    }
                }
            }
        }
    }
}

}  // namespace sri
