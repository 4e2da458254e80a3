// Development microbenchmarks: achievable IMAD.WIDE rate with the real operand patterns of fp.cuh.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/mb/mb tools/mb/mb.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../plonky2_bn254_pairing_b200/csrc/fp2.cuh"

template <int MODE>
__global__ void __launch_bounds__(128) k(u32* out, u32 iters, u32 seed) {
    u32 a[8], b[8], c[8], d[8], e[8], f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = seed * (i + 1) + threadIdx.x;
        b[i] = a[i] ^ 0x9e3779b9u * (i + 3);
        c[i] = a[i] + 0x7f4a7c15u * (i + 5);
        d[i] = b[i] + 0x12345u * (i + 7);
        e[i] = a[i] * 3u + 1u;
        f[i] = b[i] * 5u + 7u;
    }
    a[7] &= 0x1fffffffu; b[7] &= 0x1fffffffu; c[7] &= 0x1fffffffu; d[7] &= 0x1fffffffu; e[7] &= 0x1fffffffu; f[7] &= 0x1fffffffu;
#pragma unroll 1
    for (u32 it = 0; it < iters; it++) {
        if (MODE == 0) {  // wide product only: 64 MAC
            u32 T[16];
            fp_mul_wide(T, a, b);
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = T[i] ^ T[i + 8];
            a[7] &= 0x1fffffffu;
        } else if (MODE == 1) {  // fp_mul: 64 + 72
            u32 r[8];
            fp_mul(r, a, b);
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = r[i];
        } else if (MODE == 2) {  // fp2_mul: 336
            Fp2 x, y, r;
#pragma unroll
            for (int i = 0; i < 8; i++) { x.c0[i] = a[i]; x.c1[i] = b[i]; y.c0[i] = c[i]; y.c1[i] = d[i]; }
            fp2_mul(r, x, y);
#pragma unroll
            for (int i = 0; i < 8; i++) { a[i] = r.c0[i]; b[i] = r.c1[i]; }
        } else if (MODE == 3) {  // fp2_sqr: 272
            Fp2 x, r;
#pragma unroll
            for (int i = 0; i < 8; i++) { x.c0[i] = a[i]; x.c1[i] = b[i]; }
            fp2_sqr(r, x);
#pragma unroll
            for (int i = 0; i < 8; i++) { a[i] = r.c0[i]; b[i] = r.c1[i]; }
        } else if (MODE == 4) {  // two independent wide products
            u32 T[16], U[16];
            fp_mul_wide(T, a, b);
            fp_mul_wide(U, c, d);
#pragma unroll
            for (int i = 0; i < 8; i++) { a[i] = T[i] ^ T[i + 8]; c[i] = U[i] ^ U[i + 8]; }
            a[7] &= 0x1fffffffu; c[7] &= 0x1fffffffu;
        } else if (MODE == 6) {  // two independent fp2_mul in one basic block: 672 MAC
            Fp2 x, y, r, x2, y2, r2;
#pragma unroll
            for (int i = 0; i < 8; i++) { x.c0[i] = a[i]; x.c1[i] = b[i]; y.c0[i] = c[i]; y.c1[i] = d[i];
                                          x2.c0[i] = e[i]; x2.c1[i] = f[i]; y2.c0[i] = c[i] ^ 5; y2.c1[i] = d[i] ^ 3; }
            fp2_mul(r, x, y);
            fp2_mul(r2, x2, y2);
#pragma unroll
            for (int i = 0; i < 8; i++) { a[i] = r.c0[i]; b[i] = r.c1[i]; e[i] = r2.c0[i]; f[i] = r2.c1[i]; }
        } else if (MODE == 7) {  // fp2_mul plus four independent canonical fp2 additions
            Fp2 x, y, r, x2, y2, r2;
#pragma unroll
            for (int i = 0; i < 8; i++) { x.c0[i] = a[i]; x.c1[i] = b[i]; y.c0[i] = c[i]; y.c1[i] = d[i];
                                          x2.c0[i] = e[i]; x2.c1[i] = f[i]; y2.c0[i] = c[i]; y2.c1[i] = d[i]; }
            fp2_mul(r, x, y);
            fp2_add(r2, x2, y2);
            fp2_add(r2, r2, y2);
            fp2_sub(r2, r2, x2);
            fp2_add(r2, r2, y2);
#pragma unroll
            for (int i = 0; i < 8; i++) { a[i] = r.c0[i]; b[i] = r.c1[i]; e[i] = r2.c0[i]; f[i] = r2.c1[i]; }
        } else if (MODE == 5) {  // fp2 add (canonical)
            Fp2 x, y, r;
#pragma unroll
            for (int i = 0; i < 8; i++) { x.c0[i] = a[i]; x.c1[i] = b[i]; y.c0[i] = c[i]; y.c1[i] = d[i]; }
            fp2_add(r, x, y);
#pragma unroll
            for (int i = 0; i < 8; i++) { a[i] = r.c0[i]; b[i] = r.c1[i]; }
        }
    }
    u32 x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= a[i] ^ b[i] ^ c[i] ^ d[i] ^ e[i] ^ f[i];
    if (x == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <int MODE>
void run(const char* name, double macs_per_iter, int blocks_per_sm) {
    u32* d;
    cudaMalloc(&d, 1 << 24);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const u32 iters = 4096;
    int sms = 148;
    for (int bps = 1; bps <= blocks_per_sm; bps *= 2) {
        float best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            k<MODE><<<sms * bps, 128>>>(d, iters, 77u + rep);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = ms < best ? ms : best;
        }
        double warps_per_smsp = bps * 4 / 4.0;
        double ops = (double)sms * bps * 128 * iters;
        double cyc_per_warp_op_smsp = best * 1e-3 * 1.965e9 / (ops / 32 / (sms * 4));
        printf("%-10s warps/SMSP %4.1f  %8.3f ms  %.3e MAC/s  %8.1f SMSP-cycles/warp-op (pipe min %.0f)\n", name, warps_per_smsp, best,
               ops * macs_per_iter / (best * 1e-3), cyc_per_warp_op_smsp, macs_per_iter * 4);
    }
    cudaFree(d);
}

int main() {
    run<0>("mulwide", 64, 8);
    run<4>("mulwide2", 128, 8);
    run<1>("fp_mul", 136, 8);
    run<2>("fp2_mul", 336, 4);
    run<3>("fp2_sqr", 272, 4);
    run<5>("fp2_add", 0, 8);
    run<6>("fp2_mul x2", 672, 2);
    run<7>("mul+4add", 336, 2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
