// Design probe for round 2 (NOT part of the product): a register-resident, lane-cooperative Fq6 multiplication.
// Three lanes hold one Fq6 element each way: lane k of a triple owns the Fq2 coefficient a_k (and b_k) in
// REGISTERS; the schoolbook product
//     c_0 = a0 b0 + xi (a1 b2 + a2 b1),   c_1 = a0 b1 + a1 b0 + xi a2 b2,   c_2 = a0 b2 + a1 b1 + a2 b0
// fetches the partners' coefficients with warp shuffles, accumulates the three wide (unreduced) Karatsuba Fq2
// products of a lane and reduces once (2 Montgomery reductions per lane).  No shared-memory slots, no decode, no
// canonicalisation between products.  Question answered: what fraction of the IMAD.WIDE pipe does this organisation
// reach (the sequencer reaches 73 % on a pure MUL stream and 51 % on a whole pairing)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/mb/coop tools/mb/coop.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../plonky2_bn254_pairing_b200/csrc/fp2.cuh"

__device__ __forceinline__ void shfl_fp2(Fp2& r, const Fp2& v, int src) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.c0[i] = __shfl_sync(0xffffffffu, v.c0[i], src);
        r.c1[i] = __shfl_sync(0xffffffffu, v.c1[i], src);
    }
}

__device__ __forceinline__ void add16w(u32* r, const u32* a) {  // r += a (512-bit)
    asm("add.cc.u32  %0, %0, %16;\n\taddc.cc.u32 %1, %1, %17;\n\taddc.cc.u32 %2, %2, %18;\n\taddc.cc.u32 %3, %3, %19;\n\t"
        "addc.cc.u32 %4, %4, %20;\n\taddc.cc.u32 %5, %5, %21;\n\taddc.cc.u32 %6, %6, %22;\n\taddc.cc.u32 %7, %7, %23;\n\t"
        "addc.cc.u32 %8, %8, %24;\n\taddc.cc.u32 %9, %9, %25;\n\taddc.cc.u32 %10, %10, %26;\n\taddc.cc.u32 %11, %11, %27;\n\t"
        "addc.cc.u32 %12, %12, %28;\n\taddc.cc.u32 %13, %13, %29;\n\taddc.cc.u32 %14, %14, %30;\n\taddc.u32 %15, %15, %31;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
          "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]),
          "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]));
}

// MODE 0: cooperative Fq6 product (3 lanes per element);  MODE 1: same arithmetic without the shuffles (upper bound);
// MODE 2: Fq6 SQUARING-shaped use (b = a), MODE 0 otherwise
template <int MODE>
__global__ void __launch_bounds__(128) k(u32* out, u32 iters, u32 seed) {
    const int lane = threadIdx.x & 31;
    const int kk = lane % 3, base = lane - kk;   // lanes 30, 31 form an incomplete triple: they compute garbage
    Fp2 a, b;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a.c0[i] = seed * (i + 1) + threadIdx.x;
        a.c1[i] = a.c0[i] ^ 0x9e3779b9u * (i + 3);
        b.c0[i] = a.c0[i] + 0x7f4a7c15u * (i + 5);
        b.c1[i] = a.c1[i] + 0x12345u * (i + 7);
    }
    a.c0[7] &= 0x0fffffffu; a.c1[7] &= 0x0fffffffu; b.c0[7] &= 0x0fffffffu; b.c1[7] &= 0x0fffffffu;
#pragma unroll 1
    for (u32 it = 0; it < iters; it++) {
        u32 T0[16], T1[16];
#pragma unroll
        for (int t = 0; t < 3; t++) {            // term t: a_i * b_j with i = t, j = (k - t) mod 3, xi when i > k
            Fp2 x, y;
            const int i = t, j = (kk - t + 3) % 3;
            if (MODE == 1) {
                x = a; y = b;
            } else {
                shfl_fp2(x, a, min(base + i, 31));
                shfl_fp2(y, b, min(base + j, 31));
            }
            if (t > 0) {                          // xi factor on the a side when i > k (select, not a branch)
                Fp2 xx;
                fp2_mul_xi(xx, x);
                const bool need = i > kk;
#pragma unroll
                for (int q = 0; q < 8; q++) { x.c0[q] = need ? xx.c0[q] : x.c0[q]; x.c1[q] = need ? xx.c1[q] : x.c1[q]; }
            }
            if (t == 0) {
                fp2_mul_wide(T0, T1, x, y);
            } else {
                u32 U0[16], U1[16];
                fp2_mul_wide(U0, U1, x, y);
                add16w(T0, U0);
                add16w(T1, U1);
            }
        }
        Fp2 r;
        fp_redc_lazy(r.c0, T0);
        fp_redc_lazy(r.c1, T1);
        fp_canon(r.c0, 2u);
        fp_canon(r.c1, 2u);
        a = r;                                     // dependent chain, like f <- f * g in a Miller loop
    }
    u32 x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= a.c0[i] ^ a.c1[i];
    if (x == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <int MODE>
void run(const char* name) {
    u32* d;
    cudaMalloc(&d, 1 << 24);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const u32 iters = 2048;
    const int sms = 148;
    // per lane and iteration: 3 Karatsuba Fq2 products (9 x 64) + 2 reductions (2 x 72) + 2 xi (2 x (16 + 16)) MACs issued;
    // ALGORITHMIC work of one Fq6 product (Karatsuba, 18 products + 12 reductions = 2016 MACs) is shared by 3 lanes
    const double issued = 9 * 64 + 2 * 72 + 2 * 32, algorithmic = 2016.0 / 3.0;
    for (int bps = 1; bps <= 4; bps *= 2) {
        float best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            k<MODE><<<sms * bps, 128>>>(d, iters, 77u + rep);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = ms < best ? ms : best;
        }
        const double lanes = (double)sms * bps * 128 * (30.0 / 32.0);   // lanes doing useful work
        const double ops = lanes * iters;
        printf("%-22s warps/SMSP %3d  %8.3f ms  issued %.3e MAC/s  algorithmic-equivalent %.3e MAC/s  (Fq6 products/s %.3e)\n", name, bps,
               best, ops * issued / (best * 1e-3) * (32.0 / 30.0), ops * algorithmic / (best * 1e-3), ops / 3.0 / (best * 1e-3));
    }
    cudaFree(d);
}

int main() {
    run<0>("coop fq6 mul (shuffles)");
    run<1>("same, no shuffles");
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
