// Instruction-supply probe: straight-line DFMA streams of different code sizes, warps de-synchronised.
// If throughput drops once the body exceeds the instruction cache, fully unrolled elimination code is fetch-bound.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NI>  // NI DFMA instructions in the loop body (16 B each)
__global__ void __launch_bounds__(128) k_stream(double* out, int iters, double s, int stagger) {
    double a[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double b = s, c = 1.0 - s;
    // de-synchronise the warps of an SM
    const int w = (threadIdx.x >> 5) + 4 * (blockIdx.x % 8);
    double d = 0;
    for (int t = 0; t < stagger * w; ++t) d = fma(d, b, c);
    a[0] += d * 1e-30;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NI; ++i) a[i % 32] = fma(a[i % 32], b, c);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) r += a[i];
    if (r == 123.456) out[0] = r;
}

template <int NI>
static void run(int sms, int bps, int stagger) {
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = (1 << 22) / NI;
    dim3 g(sms * bps), b(128);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_stream<NI><<<g, b>>>(out, iters, 0.5, stagger); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); k_stream<NI><<<g, b>>>(out, iters, 0.5, stagger); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double fl = 2.0 * NI * (double)iters * g.x * b.x;
    printf("{\"body_kb\": %.0f, \"blocks_per_sm\": %d, \"stagger\": %d, \"tflops\": %.2f}\n", NI * 16.0 / 1024, bps, stagger, fl / best * 1e-9);
    CK(cudaFree(out));
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    for (int stagger : {0, 997}) {
        for (int bps : {1, 3}) {
            run<256>(sms, bps, stagger); run<1024>(sms, bps, stagger); run<2048>(sms, bps, stagger);
            run<4096>(sms, bps, stagger); run<8192>(sms, bps, stagger);
        }
    }
    return 0;
}
