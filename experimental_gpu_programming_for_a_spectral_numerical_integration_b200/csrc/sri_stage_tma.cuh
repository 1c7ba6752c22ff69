// sri_stage_tma.cuh -- the strain-independent stages (N <= 16) with TMA-staged streaming.
//
// Same contraction as sri_stage_dmma.cuh (one warp per tile of 8 rods, the cached 16 x 16 operator in 8 A-fragment
// registers, DMMA m8n8k4), different data movement.  There every lane loads its own (rod, node) operands straight
// from global memory as 32-byte segments; because a rod's stacks are 15 doubles per component, most of those segments
// straddle two 32-byte sectors, and the L2 -> L1 sector traffic is ~1.7 x the DRAM traffic (couple: 62 % of HBM).
// Here a tile's operands -- which are CONTIGUOUS in global memory, 8 rods x 480 / 360 / 384 bytes -- are brought into
// shared memory by one bulk-TMA copy per array (cp.async.bulk, completion on an mbarrier, two stages per warp), the
// lanes pick their operands out of shared memory at any alignment, and the result tile leaves the same way
// (cp.async.bulk shared -> global).  Requirements of the bulk copies (16-byte aligned addresses and sizes) hold for whole
// tiles of 8 rods when the base pointers are 16-byte aligned; the host routes the ragged tail (batch % 8 rods) and
// unaligned calls to sri_stage_dmma.cuh.
#pragma once
#include "sri_stage_dmma.cuh"

namespace sri {

constexpr int kStageTmaWarps = 4;

// byte offsets of one pipeline stage of one warp (16-byte aligned), computed once on the host
struct StageTmaLayout {
    int q, nin, gam, load, tip, q0, r0;  // inputs (-1: absent)
    int in_bytes;                        // bytes of one stage
    int out;                             // offset of the result tile behind the two stages
    int warp_bytes;                      // 2 * in_bytes + out tile
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* dst, unsigned src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int STAGE>
__global__ void __launch_bounds__(32 * kStageTmaWarps) stage_tma_kernel(const FusedParams p, const StageTmaLayout L, long long tiles) {
    extern __shared__ __align__(128) unsigned char tsm[];
    __shared__ __align__(8) unsigned long long bars[kStageTmaWarps][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lr = lane >> 2, lk = lane & 3;
    const int M = p.M, N = p.N;
    unsigned char* wbase = tsm + (size_t)warp * L.warp_bytes;
    const unsigned bar0 = smem_u32(&bars[warp][0]), bar1 = smem_u32(&bars[warp][1]);
    if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar1, 1); }
    fence_proxy_async();
    __syncwarp();

    const double* T = p.ops + (STAGE == kStagePosition ? StageTables::Srm : StageTables::STsh);
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

    const unsigned qb = 8u * 4 * M * 8, vb = 8u * 3 * M * 8, nb = 8u * 3 * N * 8;  // tile bytes of Q, [3][M] and [3][N] stacks
    const double* tip_src = (STAGE == kStageStress) ? p.F_tip : p.M_tip;
    const double* load_src = (STAGE == kStageStress) ? p.fbar : p.lbar;
    auto issue = [&](long long tile, int s) {  // lane 0 only
        const unsigned bar = s ? bar1 : bar0;
        const unsigned dst = smem_u32(wbase + (size_t)s * L.in_bytes);
        mbar_expect_tx(bar, (unsigned)L.in_bytes);
        if (L.q >= 0) tma_load_1d(dst + L.q, p.Qin + tile * (qb / 8), qb, bar);
        if (L.nin >= 0) tma_load_1d(dst + L.nin, p.nin + tile * (vb / 8), vb, bar);
        if (L.gam >= 0) tma_load_1d(dst + L.gam, p.Gamma + tile * (nb / 8), nb, bar);
        if (L.load >= 0) tma_load_1d(dst + L.load, load_src + tile * (nb / 8), nb, bar);
        if (L.tip >= 0) tma_load_1d(dst + L.tip, tip_src + tile * 24, 192, bar);
        if (L.q0 >= 0) tma_load_1d(dst + L.q0, p.q0 + tile * 32, 256, bar);
        if (L.r0 >= 0) tma_load_1d(dst + L.r0, p.r0 + tile * 24, 192, bar);
    };

    const long long warps_total = (long long)gridDim.x * kStageTmaWarps;
    const long long tile0 = (long long)blockIdx.x * kStageTmaWarps + warp;
    if (tile0 < tiles && lane == 0) issue(tile0, 0);
    double* outs = reinterpret_cast<double*>(wbase + L.out);
    double* out_g = (STAGE == kStagePosition) ? p.r : (STAGE == kStageStress ? p.n : p.m);

    int it = 0;
    for (long long tile = tile0; tile < tiles; tile += warps_total, ++it) {
        const int s = it & 1;
        if (tile + warps_total < tiles && lane == 0) issue(tile + warps_total, s ^ 1);
        mbar_wait(s ? bar1 : bar0, (unsigned)((it >> 1) & 1));
        const unsigned char* sb = wbase + (size_t)s * L.in_bytes;
        const double* Qs = reinterpret_cast<const double*>(sb + (L.q >= 0 ? L.q : 0)) + lr * 4 * M;
        const double* ns = reinterpret_cast<const double*>(sb + (L.nin >= 0 ? L.nin : 0)) + lr * 3 * M;
        const double* gs = reinterpret_cast<const double*>(sb + (L.gam >= 0 ? L.gam : 0)) + lr * 3 * N;
        const double* ls = reinterpret_cast<const double*>(sb + (L.load >= 0 ? L.load : 0)) + lr * 3 * N;
        const double* tips = reinterpret_cast<const double*>(sb + (L.tip >= 0 ? L.tip : 0));
        const double* q0s = reinterpret_cast<const double*>(sb + (L.q0 >= 0 ? L.q0 : 0)) + lr * 4;
        const double* r0s = reinterpret_cast<const double*>(sb + (L.r0 >= 0 ? L.r0 : 0));

        // ---- pointwise right-hand side of this lane's four (rod, node) pairs = B fragments -----------------
        double bf[3][4];
        double w0 = 0.0, w1 = 0.0, w2 = 0.0;
        if (STAGE != kStagePosition) { w0 = tips[lr * 3]; w1 = tips[lr * 3 + 1]; w2 = tips[lr * 3 + 2]; }
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            const int node = 4 * kt + lk;
            double r0 = 0.0, r1 = 0.0, r2 = 0.0;
            const bool valid = (STAGE == kStagePosition) ? (node < M) : (node >= 1 && node <= M);
            if (valid) {
                if (STAGE == kStagePosition || STAGE == kStageCouple) {
                    quat q; q.w = 1.0; q.x = 0.0; q.y = 0.0; q.z = 0.0;
                    if (node < M) { q.w = Qs[node]; q.x = Qs[M + node]; q.y = Qs[2 * M + node]; q.z = Qs[3 * M + node]; }
                    else if (L.q0 >= 0) { q.w = q0s[0]; q.x = q0s[1]; q.y = q0s[2]; q.z = q0s[3]; }
                    double b0, b1, b2;
                    if (L.gam >= 0) q_rotate(q, gs[node], gs[N + node], gs[2 * N + node], b0, b1, b2);
                    else q_rotate_e1(q, b0, b1, b2);
                    if (STAGE == kStagePosition) { r0 = b0; r1 = b1; r2 = b2; }
                    else {
                        const double n0 = ns[node - 1], n1 = ns[M + node - 1], n2 = ns[2 * M + node - 1];
                        double l0 = 0.0, l1 = 0.0, l2 = 0.0;
                        if (L.load >= 0) { l0 = ls[node]; l1 = ls[N + node]; l2 = ls[2 * N + node]; }
                        r0 = -((b1 * n2 - b2 * n1) + l0) - dti[kt] * w0;
                        r1 = -((b2 * n0 - b0 * n2) + l1) - dti[kt] * w1;
                        r2 = -((b0 * n1 - b1 * n0) + l2) - dti[kt] * w2;
                    }
                } else {  // stress
                    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
                    if (L.load >= 0) { f0 = ls[node]; f1 = ls[N + node]; f2 = ls[2 * N + node]; }
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
        // ---- epilogue: C fragment (node 8*mt + lane/4, rods 2*(lane%4) + {0,1}) -> result tile in shared memory ----
        if (lane == 0) tma_store_wait_read();  // the previous tile's store has finished reading `outs`
        __syncwarp();
#pragma unroll
        for (int w = 0; w < 2; ++w) {
            const int orod = 2 * lk + w;
            double e0 = 0.0, e1 = 0.0, e2 = 0.0;
            if (STAGE == kStagePosition && L.r0 >= 0) { e0 = r0s[orod * 3]; e1 = r0s[orod * 3 + 1]; e2 = r0s[orod * 3 + 2]; }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int i = 8 * mt + lr;
                if (i >= M) continue;
                double* d = outs + orod * 3 * M + i;
                d[0] = fma(gvec[mt], e0, acc[0][mt][w]);
                d[M] = fma(gvec[mt], e1, acc[1][mt][w]);
                d[2 * M] = fma(gvec[mt], e2, acc[2][mt][w]);
            }
        }
        fence_proxy_async();  // generic-proxy writes of `outs` before the async-proxy read of the bulk store
        __syncwarp();         // also: every lane has finished reading stage s before it is refilled next iteration
        if (lane == 0) { tma_store_1d(out_g + tile * (vb / 8), smem_u32(outs), vb); tma_store_commit(); }
    }
    if (lane == 0) tma_store_wait_all();
}

}  // namespace sri
