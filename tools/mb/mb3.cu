// Probe (informational): B200 keeps full-rate FP64.  How does a 52-bit-limb DFMA product compare with
// the 8x32 IMAD.WIDE formulation the north star prescribes?
//   dfma_peak : dependency-light DFMA chains (the FP64 pipe's issue rate)
//   dfma_row  : one 260x260-bit schoolbook product in 5 x 52-bit limbs held in doubles: every limb product is taken as
//               hi = fma_rz(a, b, 2^104) - 2^104 and lo = fma_rz(a, b, -hi) (two DFMA), the 25 (hi, lo) pairs are summed
//               per column as 64-bit integers (the bit patterns of exactly representable doubles), no carry resolution.
//               That is the multiply core of a DFMA Montgomery multiplication - 50 DFMA + 2 DADD-class ops per pair.
//   imad_row  : fp_mul_wide of fp.cuh (64 IMAD.WIDE + 22 adds) for the same 256-bit product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/mb/mb3 tools/mb/mb3.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../plonky2_bn254_pairing_b200/csrc/fp.cuh"

__global__ void __launch_bounds__(256) dfma_peak(double* out, u32 iters, double seed) {
    double a = seed + threadIdx.x, b = 1.0000001;
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; c++) acc[c] = a + c;
#pragma unroll 1
    for (u32 it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 8; c++) acc[c] = __fma_rn(acc[c], b, a);
    }
    double x = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) x += acc[c];
    if (x == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

__global__ void __launch_bounds__(128) dfma_row(u64* out, u32 iters, u32 seed) {
    double a[5], b[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        a[i] = (double)((u64)(seed * (i + 1) + threadIdx.x) & ((1ull << 52) - 1));
        b[i] = (double)((u64)(seed * (i + 7) + 3 * threadIdx.x) & ((1ull << 52) - 1));
    }
    const double C = 20282409603651670423947251286016.0;  // 2^104
    u64 col[10];
#pragma unroll
    for (int i = 0; i < 10; i++) col[i] = 0;
#pragma unroll 1
    for (u32 it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 5; i++)
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const double hi = __fma_rz(a[i], b[j], C);       // 2^104 + floor(ab / 2^52) * 2^52
                const double lo = __fma_rz(a[i], b[j], C - hi);  // ab - hi part, exact
                col[i + j + 1] += (u64)__double_as_longlong(hi); // bit patterns of same-exponent doubles add as integers
                col[i + j] += (u64)__double_as_longlong(lo);
            }
        a[0] = (double)(col[3] & ((1ull << 52) - 1));  // keep a data dependence between iterations
    }
    u64 x = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) x ^= col[i];
    if (x == 0x1234567812345678ull) out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

__global__ void __launch_bounds__(128) imad_row(u32* out, u32 iters, u32 seed) {
    u32 a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed * (i + 1) + threadIdx.x; b[i] = a[i] ^ 0x9e3779b9u * (i + 3); }
    a[7] &= 0x1fffffffu; b[7] &= 0x1fffffffu;
#pragma unroll 1
    for (u32 it = 0; it < iters; it++) {
        u32 T[16];
        fp_mul_wide(T, a, b);
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = T[i] ^ T[i + 8];
        a[7] &= 0x1fffffffu;
    }
    u32 x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= a[i];
    if (x == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <typename F>
static float best_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    return best;
}

int main() {
    void* d;
    cudaMalloc(&d, 1 << 26);
    const int sms = 148;
    {
        const u32 iters = 1 << 14;
        float ms = best_ms([&] { dfma_peak<<<sms * 8, 256>>>((double*)d, iters, 3.0); });
        double n = (double)sms * 8 * 256 * iters * 32;
        printf("dfma_peak   %8.3f ms  %.3e DFMA/s  (%.1f per clk per SM at 1.965 GHz)\n", ms, n / (ms * 1e-3), n / (ms * 1e-3) / 148 / 1.965e9);
    }
    for (int bps = 4; bps <= 16; bps *= 2) {
        const u32 iters = 4096;
        float ms = best_ms([&] { dfma_row<<<sms * bps, 128>>>((u64*)d, iters, 77u); });
        double n = (double)sms * bps * 128 * iters;
        printf("dfma_row    warps/SM %3d  %8.3f ms  %.3e products/s  (%.1f SM-cycles per warp-product)\n", bps * 4, ms, n / (ms * 1e-3),
               ms * 1e-3 * 1.965e9 / (n / 32 / sms));
        ms = best_ms([&] { imad_row<<<sms * bps, 128>>>((u32*)d, iters, 77u); });
        printf("imad_row    warps/SM %3d  %8.3f ms  %.3e products/s  (%.1f SM-cycles per warp-product)\n", bps * 4, ms, n / (ms * 1e-3),
               ms * 1e-3 * 1.965e9 / (n / 32 / sms));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
