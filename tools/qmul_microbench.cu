// Register-only throughput of the quaternion rank-1 update c[j] -= u[j] (x) m (16 DFMA each), no memory traffic.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../experimental_gpu_programming_for_a_spectral_numerical_integration_b200/csrc/sri_device.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
using namespace sri;

template <int NQ>
__global__ void __launch_bounds__(128) k_qupd(double* out, int iters, double s) {
    quat c[NQ], u[4];
#pragma unroll
    for (int j = 0; j < NQ; ++j) { c[j].w = threadIdx.x + j; c[j].x = j; c[j].y = 1.0; c[j].z = s; }
#pragma unroll
    for (int j = 0; j < 4; ++j) { u[j].w = s * j; u[j].x = 1e-3 * threadIdx.x; u[j].y = s; u[j].z = j; }
    quat m; m.w = s; m.x = 0.5 * s; m.y = 0.25 * s; m.z = 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NQ; ++j) q_sub_mul(c[j], u[j & 3], m);
        m.w += 1e-9;  // keep the multiplier loop-variant
    }
    double r = 0;
#pragma unroll
    for (int j = 0; j < NQ; ++j) r += c[j].w + c[j].x + c[j].y + c[j].z;
    if (r == 123.456) out[0] = r;
}

template <int NQ>
static void run(int sms, int wps) {
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 2048;
    dim3 g(sms * (wps / 4)), b(128);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_qupd<NQ><<<g, b>>>(out, iters, 0.5); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0)); k_qupd<NQ><<<g, b>>>(out, iters, 0.5); CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    double fl = 2.0 * 16 * NQ * iters * (double)g.x * b.x;
    printf("{\"bench\": \"q_sub_mul x%d\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", NQ, wps, fl / best * 1e-9);
    CK(cudaFree(out));
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    for (int wps : {4, 8, 12}) { run<4>(p.multiProcessorCount, wps); run<8>(p.multiProcessorCount, wps); run<15>(p.multiProcessorCount, wps); }
    return 0;
}
