// sri_stage_generic_tma.cuh -- the strain-independent stages for 17 <= N <= 64 with TMA-staged inputs.
//
// Contraction and operator tables as in sri_stage_generic.cuh; data movement as in sri_stage_tma.cuh: a tile's
// operands are contiguous in global memory (8 rods x 4 M / 3 M / 3 N doubles), so one cp.async.bulk per array brings
// them into this warp's shared-memory buffer (completion on an mbarrier).  One buffer per warp: the lanes first turn the
// buffer into their B fragments (registers), then the copies of the NEXT tile are issued into the same buffer and fly
// while this tile's R/8 x R/4 x 3 DMMAs run -- at N = 64 that is 384 DMMAs, long enough to cover the copy.  Needs whole
// tiles of 8 rods and 16-byte aligned base pointers; the host sends the ragged tail and unaligned calls through
// sri_stage_generic.cuh.
#pragma once
#include "sri_stage_generic.cuh"
#include "sri_stage_tma.cuh"

namespace sri {

template <int STAGE, int R>
__global__ void __launch_bounds__(128) stage_generic_tma_kernel(const FusedParams p, const StageTmaLayout L, long long tiles) {
    constexpr int KT = R / 4;
    extern __shared__ __align__(128) unsigned char gts[];  // R*R doubles of operator table, then one buffer per warp
    __shared__ __align__(8) unsigned long long bars[4];
    double* gsm = reinterpret_cast<double*>(gts);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lr = lane >> 2, lk = lane & 3;
    const int M = p.M, N = p.N;
    {
        const double* src = p.ops2 + (size_t)R * R * (STAGE == kStagePosition ? 1 : 2);
        for (int i = threadIdx.x; i < R * R; i += blockDim.x) gsm[i] = src[i];
    }
    unsigned char* wbuf = gts + (size_t)R * R * sizeof(double) + (size_t)warp * L.in_bytes;
    const unsigned bar = smem_u32(&bars[warp]);
    if (lane == 0) mbar_init(bar, 1);
    fence_proxy_async();
    __syncthreads();
    const int kt_used = (STAGE == kStageStress && L.load < 0) ? 0 : (M + 3) >> 2;
    const int mt_used = (M + 7) >> 3;

    const unsigned qb = 8u * 4 * M * 8, vb = 8u * 3 * M * 8, nb = 8u * 3 * N * 8;
    const double* tip_src = (STAGE == kStagePosition) ? p.r0 : (STAGE == kStageStress ? p.F_tip : p.M_tip);
    const double* load_src = (STAGE == kStageStress) ? p.fbar : p.lbar;
    auto issue = [&](long long tile) {  // lane 0 only
        const unsigned dst = smem_u32(wbuf);
        mbar_expect_tx(bar, (unsigned)L.in_bytes);
        if (L.q >= 0) tma_load_1d(dst + L.q, p.Qin + tile * (qb / 8), qb, bar);
        if (L.nin >= 0) tma_load_1d(dst + L.nin, p.nin + tile * (vb / 8), vb, bar);
        if (L.gam >= 0) tma_load_1d(dst + L.gam, p.Gamma + tile * (nb / 8), nb, bar);
        if (L.load >= 0) tma_load_1d(dst + L.load, load_src + tile * (nb / 8), nb, bar);
        if (L.tip >= 0) tma_load_1d(dst + L.tip, tip_src + tile * 24, 192, bar);
        if (L.q0 >= 0) tma_load_1d(dst + L.q0, p.q0 + tile * 32, 256, bar);
    };

    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    const long long tile0 = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (tile0 < tiles && lane == 0) issue(tile0);
    double* out = (STAGE == kStagePosition) ? p.r : (STAGE == kStageStress ? p.n : p.m);

    int it = 0;
    for (long long tile = tile0; tile < tiles; tile += warps_total, ++it) {
        mbar_wait(bar, (unsigned)(it & 1));
        const double* Qs = reinterpret_cast<const double*>(wbuf + (L.q >= 0 ? L.q : 0)) + lr * 4 * M;
        const double* ns = reinterpret_cast<const double*>(wbuf + (L.nin >= 0 ? L.nin : 0)) + lr * 3 * M;
        const double* gs = reinterpret_cast<const double*>(wbuf + (L.gam >= 0 ? L.gam : 0)) + lr * 3 * N;
        const double* ls = reinterpret_cast<const double*>(wbuf + (L.load >= 0 ? L.load : 0)) + lr * 3 * N;
        const double* tips = reinterpret_cast<const double*>(wbuf + (L.tip >= 0 ? L.tip : 0)) + lr * 3;
        const double* q0s = reinterpret_cast<const double*>(wbuf + (L.q0 >= 0 ? L.q0 : 0)) + lr * 4;
        // ---- B fragments from the staged tile ---------------------------------------------------------------------
        double bf[3][KT];
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            const int j = 4 * kt + lk;
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
            if (kt < kt_used || kt == KT - 1) {
                if (j < M) {
                    if (STAGE == kStageStress) {
                        if (L.load >= 0) { r0 = ls[j + 1]; r1 = ls[N + j + 1]; r2 = ls[2 * N + j + 1]; }
                    } else {
                        const int node = (STAGE == kStagePosition) ? j : j + 1;
                        quat q; q.w = 1.0; q.x = 0.0; q.y = 0.0; q.z = 0.0;
                        if (node < M) { q.w = Qs[node]; q.x = Qs[M + node]; q.y = Qs[2 * M + node]; q.z = Qs[3 * M + node]; }
                        else if (L.q0 >= 0) { q.w = q0s[0]; q.x = q0s[1]; q.y = q0s[2]; q.z = q0s[3]; }
                        double b0, b1, b2;
                        if (L.gam >= 0) q_rotate(q, gs[node], gs[N + node], gs[2 * N + node], b0, b1, b2);
                        else q_rotate_e1(q, b0, b1, b2);
                        if (STAGE == kStagePosition) { r0 = b0; r1 = b1; r2 = b2; }
                        else {
                            const double n0 = ns[j], n1 = ns[M + j], n2 = ns[2 * M + j];
                            double l0 = 0.0, l1 = 0.0, l2 = 0.0;
                            if (L.load >= 0) { l0 = ls[node]; l1 = ls[N + node]; l2 = ls[2 * N + node]; }
                            r0 = fma(b1, n2, fma(-b2, n1, l0));
                            r1 = fma(b2, n0, fma(-b0, n2, l1));
                            r2 = fma(b0, n1, fma(-b1, n0, l2));
                        }
                    }
                } else if (j == R - 1 && L.tip >= 0) {
                    r0 = tips[0]; r1 = tips[1]; r2 = tips[2];
                }
            }
            bf[0][kt] = r0; bf[1][kt] = r1; bf[2][kt] = r2;
        }
        __syncwarp();  // every lane has its operands in registers: the buffer may be refilled
        if (tile + warps_total < tiles && lane == 0) issue(tile + warps_total);
        // ---- Out = At * Rhs while the next tile's copies are in flight ----------------------------------------------
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
                    double* d = out + (tile * 8 + 2 * lk + w) * 3 * M + i;
                    d[0] = acc[0][w]; d[M] = acc[1][w]; d[2 * M] = acc[2][w];
                }
            }
        }
    }
}

}  // namespace sri
