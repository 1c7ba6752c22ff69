// Shared-memory / shuffle instruction cost probes for B200 (sm_100a).
// Reports SM-cycles per warp-instruction when 8 or 16 warps per SM issue the op back to back.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

enum { STS128 = 0, STS64 = 1, STS32 = 2, LDS128 = 3, LDS64 = 4, SHFL = 5, REDUX = 6 };

// mask: which lanes take part in stores; naddr: number of distinct addresses for loads (1 = broadcast)
template <int OP>
__global__ void k_probe(double* out, int iters, unsigned mask, int naddr) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* base = reinterpret_cast<double2*>(smraw) + warp * 64;
    base[lane] = make_double2(lane, 1.0);
    base[lane + 32] = make_double2(lane, 2.0);
    __syncwarp();
    const bool active = (mask >> lane) & 1u;
    const int slot = (naddr >= 32) ? lane : (lane % naddr);
    double2 acc = make_double2(0.0, 0.0);
    double2 v = make_double2(lane, 3.0);
    unsigned ia = lane;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == STS128) { if (active) asm volatile("st.shared.v2.f64 [%0], {%1,%2};" :: "r"((unsigned)__cvta_generic_to_shared(base + ((u * 4 + lane) & 63))), "d"(v.x), "d"(v.y) : "memory"); }
            if (OP == STS64)  { if (active) asm volatile("st.shared.f64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(base + ((u * 4 + lane) & 63))), "d"(v.x) : "memory"); }
            if (OP == STS32)  { if (active) asm volatile("st.shared.u32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(base + ((u * 4 + lane) & 63))), "r"(ia) : "memory"); }
            if (OP == LDS128) { double2 t; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(t.x), "=d"(t.y) : "r"((unsigned)__cvta_generic_to_shared(base + slot + u)) : "memory"); acc.x += t.x; acc.y += t.y; }
            if (OP == LDS64)  { double t; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"((unsigned)__cvta_generic_to_shared(base + slot + u)) : "memory"); acc.x += t; }
            if (OP == SHFL)   { ia = __shfl_sync(0xffffffffu, ia * 3 + u, (ia + u) & 31); }
            if (OP == REDUX)  { ia = __reduce_max_sync(0xffffffffu, ia + lane + u); }
        }
    }
    if (acc.x + acc.y == 123.456 || ia == 0x12345678u) out[0] = acc.x + ia;
}

template <int OP>
static void run(const char* name, int sms, double clk_hz, unsigned mask, int naddr, int wps) {
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 2048;
    dim3 g(sms), b(32 * wps);
    size_t sh = wps * 64 * sizeof(double2);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_probe<OP><<<g, b, sh>>>(out, iters, mask, naddr);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); k_probe<OP><<<g, b, sh>>>(out, iters, mask, naddr); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double ops_per_sm = 8.0 * iters * wps;
    printf("{\"probe\": \"%s\", \"mask\": \"0x%08x\", \"naddr\": %d, \"warps_per_sm\": %d, \"sm_cycles_per_warp_instr\": %.3f}\n",
           name, mask, naddr, wps, best * 1e-3 * clk_hz / ops_per_sm);
    CK(cudaFree(out));
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount; const double clk = p.clockRate * 1e3;
    for (int wps : {8, 16}) {
        for (unsigned mask : {0x1u, 0x3u, 0xFu, 0xFFu, 0x11111111u, 0x01010101u, 0x0000FFFFu, 0xFFFFFFFFu}) {
            run<STS128>("sts128", sms, clk, mask, 0, wps);
            run<STS64>("sts64", sms, clk, mask, 0, wps);
            run<STS32>("sts32", sms, clk, mask, 0, wps);
        }
        for (int naddr : {1, 2, 4, 8, 16, 32}) {
            run<LDS128>("lds128", sms, clk, 0, naddr, wps);
            run<LDS64>("lds64", sms, clk, 0, naddr, wps);
        }
        run<SHFL>("shfl_idx", sms, clk, 0, 0, wps);
        run<REDUX>("redux_max_u32", sms, clk, 0, 0, wps);
    }
    return 0;
}
