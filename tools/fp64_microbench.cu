// FP64 pipe micro-benchmarks for B200 (sm_100a): establishes the roofline denominators that
// MEASURED_PEAKS.json lacks (it holds only HBM and bf16 figures).
//   dfma      : register-only dependent-chain-free DFMA stream
//   dmma      : mma.sync.m8n8k4.f64 stream
//   dfma_lds  : 4 DFMA per broadcast LDS.128 (the Gauss-Jordan inner loop's instruction mix)
//   dfma_lds_sts : same plus one single-lane STS.128 per LDS
//   shfl      : SHFL.IDX stream
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double s) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-9 + i;
    double b = s, c = 1.0 - s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += a[i];
    if (r == 123.456) out[0] = r;
}

template <int TILES>
__global__ void k_dmma(double* out, int iters, double s) {
    double c0[TILES], c1[TILES];
#pragma unroll
    for (int i = 0; i < TILES; ++i) { c0[i] = threadIdx.x; c1[i] = i; }
    double a = s, b = 1.0 - s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < TILES; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < TILES; ++i) r += c0[i] + c1[i];
    if (r == 123.456) out[0] = r;
}

// 16 complex accumulators per lane; each iteration: 16 x (LDS.128 broadcast + 4 DFMA) [+ single-lane STS.128]
template <bool WITH_STS>
__global__ void k_dfma_lds(double* out, int iters, double s) {
    extern __shared__ double2 sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2* row = sm + warp * 32;
    if (lane < 32) row[lane] = make_double2(s * lane, 1.0 - s);
    __syncwarp();
    double2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_double2(lane, i);
    double2 f = make_double2(s, 0.5 * s);
    for (int it = 0; it < iters; ++it) {
        if (WITH_STS) {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (lane == (it & 31)) row[i] = acc[i];
            __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            double2 u = row[i];
            acc[i].x = fma(-f.x, u.x, acc[i].x);
            acc[i].x = fma(f.y, u.y, acc[i].x);
            acc[i].y = fma(-f.x, u.y, acc[i].y);
            acc[i].y = fma(-f.y, u.x, acc[i].y);
        }
        if (WITH_STS) __syncwarp();
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += acc[i].x + acc[i].y;
    if (r == 123.456) out[0] = r;
}

__global__ void k_shfl(double* out, int iters) {
    int v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __shfl_sync(0xffffffffu, v[i], (it + i) & 31);
    }
    int r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += v[i];
    if (r == 123456789) out[0] = r;
}

template <typename F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, 8));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    const int iters = 4096;
    // sweep warps per SM
    for (int wps : {4, 8, 12, 16, 32}) {
        int threads = 32 * wps > 1024 ? 1024 : 32 * wps;
        int bpsm = (32 * wps + threads - 1) / threads;
        dim3 g(sms * bpsm), b(threads);
        {
            float ms = time_ms([&] { k_dfma<16><<<g, b>>>(out, iters, 0.5); }, 5);
            double fl = 2.0 * 16 * iters * (double)g.x * b.x;
            printf("{\"bench\": \"dfma\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", wps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_dmma<8><<<g, b>>>(out, iters, 0.5); }, 5);
            double fl = 2.0 * 256 * 8 * iters * (double)g.x * (b.x / 32);
            printf("{\"bench\": \"dmma_m8n8k4\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", wps, ms, fl / ms * 1e-9);
        }
        {
            size_t sh = (threads / 32) * 32 * sizeof(double2);
            float ms = time_ms([&] { k_dfma_lds<false><<<g, b, sh>>>(out, iters, 0.5); }, 5);
            double fl = 2.0 * 64 * iters * (double)g.x * b.x;
            printf("{\"bench\": \"dfma_lds128\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", wps, ms, fl / ms * 1e-9);
            ms = time_ms([&] { k_dfma_lds<true><<<g, b, sh>>>(out, iters, 0.5); }, 5);
            printf("{\"bench\": \"dfma_lds128_sts1\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}\n", wps, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_shfl<<<g, b>>>(out, iters); }, 5);
            double n = 8.0 * iters * (double)g.x * (b.x / 32);
            printf("{\"bench\": \"shfl\", \"warps_per_sm\": %d, \"ms\": %.4f, \"warp_shfl_per_clk_per_sm\": %.3f}\n", wps, ms,
                   n / (ms * 1e-3) / sms / (p.clockRate * 1e3));
        }
    }
    // sustained DFMA (about 2 s) to see the power-capped clock
    {
        dim3 g(sms * 2), b(512);
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0;
        for (; n < 400; ++n) k_dfma<16><<<g, b>>>(out, iters * 4, 0.5);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * 16 * iters * 4 * (double)g.x * b.x * n;
        printf("{\"bench\": \"dfma_sustained\", \"seconds\": %.3f, \"tflops\": %.3f}\n", ms * 1e-3, fl / ms * 1e-9);
    }
    CK(cudaFree(out));
    return 0;
}
