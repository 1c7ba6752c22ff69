// Second round of LSU probes: independent SHFL throughput, and 2-lane publishes (one lane in each half-warp),
// conflict-free addresses.  SM-cycles per warp-instruction with 8/16 warps per SM issuing back to back.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int OP>
__global__ void k_probe(unsigned* out, int iters, int srcA, int srcB) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned sbase = (unsigned)__cvta_generic_to_shared(smraw) + warp * 1024 + (lane >> 4) * 512;
    const int src = (lane < 16) ? srcA : srcB;
    const bool act = (lane == srcA) || (lane == srcB);
    unsigned v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 8 + i;
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        if (OP == 0) {  // 8 independent SHFL.IDX (source lane differs per half-warp)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += __shfl_sync(0xffffffffu, v[i] + it, src);
        } else if (OP == 1) {  // 2 x STS.128 from two lanes
            if (act) {
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(sbase), "r"(v[0] + it), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(sbase + 16), "r"(v[4] + it), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
            }
        } else if (OP == 2) {  // 4 x STS.64
            if (act) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(sbase + 8 * i), "r"(v[2 * i] + it), "r"(v[2 * i + 1]) : "memory");
            }
        } else if (OP == 3) {  // 8 x STS.32
            if (act) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("st.shared.u32 [%0], %1;" :: "r"(sbase + 4 * i), "r"(v[i] + it) : "memory");
            }
        } else if (OP == 4) {  // 2 x LDS.128, two distinct addresses (one per half-warp), integer consumer
            unsigned a, b, c, d;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(sbase + (it & 15) * 32) : "memory");
            acc += a ^ b ^ c ^ d;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(sbase + (it & 15) * 32 + 16) : "memory");
            acc += a ^ b ^ c ^ d;
        } else if (OP == 5) {  // 2 x LDS.128, single address for the whole warp
            unsigned a, b, c, d;
            unsigned s1 = (unsigned)__cvta_generic_to_shared(smraw) + warp * 1024;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(s1 + (it & 15) * 32) : "memory");
            acc += a ^ b ^ c ^ d;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(s1 + (it & 15) * 32 + 16) : "memory");
            acc += a ^ b ^ c ^ d;
        } else if (OP == 6) {  // SHFL + STS mix: do they share a pipe?  8 SHFL + 2 STS.128
#pragma unroll
            for (int i = 0; i < 8; ++i) acc += __shfl_sync(0xffffffffu, v[i] + it, src);
            if (act) {
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(sbase), "r"(v[0] + it), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(sbase + 16), "r"(v[4] + it), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
            }
        } else if (OP == 7) {  // 8 lanes (4 per half) x 2 STS.128
            if ((lane & 15) < 4) {
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(sbase + (lane & 15) * 32), "r"(v[0] + it), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(sbase + (lane & 15) * 32 + 16), "r"(v[4] + it), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
            }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

static const char* names[] = {"8xSHFL.IDX(2 src lanes)", "2xSTS.128 (2 lanes)", "4xSTS.64 (2 lanes)", "8xSTS.32 (2 lanes)",
                              "2xLDS.128 (2 addr)", "2xLDS.128 (1 addr)", "8xSHFL + 2xSTS.128", "2xSTS.128 (8 lanes)"};
static const int ninstr[] = {8, 2, 4, 8, 2, 2, 10, 2};

template <int OP>
static void run(int sms, double clk, int wps) {
    unsigned* out; CK(cudaMalloc(&out, 8));
    const int iters = 4096;
    dim3 g(sms), b(32 * wps);
    size_t sh = wps * 1024;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_probe<OP><<<g, b, sh>>>(out, iters, 3, 21); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); k_probe<OP><<<g, b, sh>>>(out, iters, 3, 21); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double cyc_per_iter = best * 1e-3 * clk / ((double)iters * wps);
    printf("{\"probe\": \"%s\", \"warps_per_sm\": %d, \"sm_cycles_per_group\": %.2f, \"sm_cycles_per_instr\": %.2f}\n", names[OP], wps,
           cyc_per_iter, cyc_per_iter / ninstr[OP]);
    CK(cudaFree(out));
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount; const double clk = p.clockRate * 1e3;
    for (int wps : {8, 16, 32}) {
        run<0>(sms, clk, wps); run<1>(sms, clk, wps); run<2>(sms, clk, wps); run<3>(sms, clk, wps);
        run<4>(sms, clk, wps); run<5>(sms, clk, wps); run<6>(sms, clk, wps); run<7>(sms, clk, wps);
    }
    return 0;
}
