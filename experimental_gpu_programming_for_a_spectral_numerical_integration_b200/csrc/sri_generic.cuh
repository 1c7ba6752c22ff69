// sri_generic.cuh -- operator-table layout shared by the kernels for 17 <= N <= 64 (sri_tiled.cuh).
#pragma once

namespace sri {

// Packed operator tables (doubles), row stride R (32 or 64 = row capacity of the kernel), zero padded.
struct OpsLayoutGeneric {
    int R;
    __host__ __device__ int Sp() const { return 0; }                 // [63][R]  Sp[j*R+i]  = (Dn_NN^-1)(i,j)
    __host__ __device__ int STt() const { return 63 * R; }           // [63][R]  STt[j*R+i] = (D_TT^-1)(i,j)
    __host__ __device__ int g() const { return 2 * 63 * R; }         // [R]  g  = -(Dn_NN^-1 Dn_IN)
    __host__ __device__ int gT() const { return 2 * 63 * R + R; }    // [R]  gT = -(D_TT^-1 D_TI)
    __host__ __device__ int DTI() const { return 2 * 63 * R + 2 * R; }
    __host__ __device__ int total() const { return 2 * 63 * R + 3 * R; }
};

}  // namespace sri
