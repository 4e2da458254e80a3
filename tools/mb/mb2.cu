// Design probe for round 2: a miniature Fq2 sequencer with ONE THREAD PER FULL Fq2 OPERATION (Karatsuba per lane),
// run at the occupancies the two candidate slot layouts allow:
//   8 warps/SM, 14 slots of 64 B per thread  (one thread = one pairing)
//  12 warps/SM,  9 slots per thread           (two lanes share an 18-slot file)
//  16 warps/SM,  7 slots per thread           (two lanes share a 13/14-slot file)
// Measures SMSP-cycles per warp-instruction for MUL / SQR / ADD streams and a program-like mix.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/mb/mb2 tools/mb/mb2.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../plonky2_bn254_pairing_b200/csrc/fp2.cuh"

template <int T>
struct Slots {
    uint4* base;
    __device__ __forceinline__ void load(Fp2& r, u32 s) const {
        const uint4* p = base + s * (4 * T);
        uint4 q0 = p[0], q1 = p[T], q2 = p[2 * T], q3 = p[3 * T];
        r.c0[0] = q0.x; r.c0[1] = q0.y; r.c0[2] = q0.z; r.c0[3] = q0.w;
        r.c0[4] = q1.x; r.c0[5] = q1.y; r.c0[6] = q1.z; r.c0[7] = q1.w;
        r.c1[0] = q2.x; r.c1[1] = q2.y; r.c1[2] = q2.z; r.c1[3] = q2.w;
        r.c1[4] = q3.x; r.c1[5] = q3.y; r.c1[6] = q3.z; r.c1[7] = q3.w;
    }
    __device__ __forceinline__ void store(u32 s, const Fp2& r) const {
        uint4* p = base + s * (4 * T);
        p[0] = make_uint4(r.c0[0], r.c0[1], r.c0[2], r.c0[3]);
        p[T] = make_uint4(r.c0[4], r.c0[5], r.c0[6], r.c0[7]);
        p[2 * T] = make_uint4(r.c1[0], r.c1[1], r.c1[2], r.c1[3]);
        p[3 * T] = make_uint4(r.c1[4], r.c1[5], r.c1[6], r.c1[7]);
    }
};

// Karatsuba with the three products' rows interleaved in program order (a hint for ptxas: six independent chains)
__device__ __forceinline__ void fp2_mul_wide_il(u32* T0, u32* T1, const Fp2& a, const Fp2& b) {
    u32 sa[8], sb[8];
    add8(sa, a.c0, a.c1);
    add8(sb, b.c0, b.c1);
    u32 E0[16], O0[14], E1[16], O1[14], E2[16], O2[14];
#define ROW0(E, O, x, y) chain_fresh<0>(E, x[0], y[0], y[2], y[4], y[6]); chain_fresh<0>(O, x[0], y[1], y[3], y[5], y[7]);
#define ROW(A, C, B_, D_, i, x, y) mul_row<B_, D_>(A, C, x[i], y[0], y[2], y[4], y[6], y[1], y[3], y[5], y[7]);
    ROW0(E0, O0, a.c0, b.c0) ROW0(E1, O1, a.c1, b.c1) ROW0(E2, O2, sa, sb)
    ROW(O0, E0, 0, 2, 1, a.c0, b.c0) ROW(O1, E1, 0, 2, 1, a.c1, b.c1) ROW(O2, E2, 0, 2, 1, sa, sb)
    ROW(E0, O0, 2, 2, 2, a.c0, b.c0) ROW(E1, O1, 2, 2, 2, a.c1, b.c1) ROW(E2, O2, 2, 2, 2, sa, sb)
    ROW(O0, E0, 2, 4, 3, a.c0, b.c0) ROW(O1, E1, 2, 4, 3, a.c1, b.c1) ROW(O2, E2, 2, 4, 3, sa, sb)
    ROW(E0, O0, 4, 4, 4, a.c0, b.c0) ROW(E1, O1, 4, 4, 4, a.c1, b.c1) ROW(E2, O2, 4, 4, 4, sa, sb)
    ROW(O0, E0, 4, 6, 5, a.c0, b.c0) ROW(O1, E1, 4, 6, 5, a.c1, b.c1) ROW(O2, E2, 4, 6, 5, sa, sb)
    ROW(E0, O0, 6, 6, 6, a.c0, b.c0) ROW(E1, O1, 6, 6, 6, a.c1, b.c1) ROW(E2, O2, 6, 6, 6, sa, sb)
    ROW(O0, E0, 6, 8, 7, a.c0, b.c0) ROW(O1, E1, 6, 8, 7, a.c1, b.c1) ROW(O2, E2, 6, 8, 7, sa, sb)
    u32 P1[16];
    wide_merge(T0, E0, O0);
    wide_merge(P1, E1, O1);
    wide_merge(T1, E2, O2);
    sub16(T1, T1, T0);
    sub16(T1, T1, P1);
    u32 borrow = sub16(T0, T0, P1);
    add_p_masked(T0 + 8, borrow);
}

struct Args {
    const u64* prog;
    u32 nins;
    u32 reps;
    u32* out;
};

// ops: 1 MUL d = a*c, 2 SQR d = a^2, 3 ADD d = a+b, 4 MUL (interleaved Karatsuba), 5 MUL with pre-additions (a+b)*(c+e)
template <int T, int MINB, int NS>
__global__ void __launch_bounds__(T, MINB) vmk(Args args) {
    extern __shared__ uint4 smem[];
    Slots<T> S;
    S.base = smem + threadIdx.x;
    {
        Fp2 v;
#pragma unroll
        for (int i = 0; i < 8; i++) { v.c0[i] = 0x01234567u * (i + 1) + threadIdx.x; v.c1[i] = 0x089abcdeu * (i + 3) + blockIdx.x; }
        v.c0[7] &= 0x1fffffffu; v.c1[7] &= 0x1fffffffu;
        for (int s = 0; s < NS; s++) { v.c0[0] += s; S.store(s, v); }
    }
    for (u32 rep = 0; rep < args.reps; rep++) {
        const u64* pc = args.prog;
        u64 ins = __ldg(pc++);
        for (u32 k = 0; k < args.nins; k++) {
            const u32 lo = (u32)ins, hi = (u32)(ins >> 32);
            const u32 op = lo & 0xffu, d = (lo >> 8) & 0xffu, a = (lo >> 16) & 0xffu, b = lo >> 24;
            const u32 c = hi & 0xffu, ee = (hi >> 8) & 0xffu;
            const u64 nxt = __ldg(pc++);
            Fp2 x, y, r;
            if (op == 1u || op == 4u || op == 5u) {
                S.load(x, a);
                S.load(y, c);
                if (op == 5u) {
                    Fp2 t;
                    S.load(t, b);
                    fp2_add_lazy(x, x, t);
                    S.load(t, ee);
                    fp2_add_lazy(y, y, t);
                }
                u32 T0[16], T1[16];
                if (op == 4u) fp2_mul_wide_il(T0, T1, x, y); else fp2_mul_wide(T0, T1, x, y);
                fp_redc_lazy(r.c0, T0);
                fp_redc_lazy(r.c1, T1);
                fp_canon(r.c0, op == 5u ? 1u : 0u);
                fp_canon(r.c1, op == 5u ? 1u : 0u);
                S.store(d, r);
            } else if (op == 6u) {   // two independent MULs in one instruction: d = a*c, d+1 = b*e
                Fp2 x2, y2, r2;
                S.load(x, a);
                S.load(y, c);
                S.load(x2, b);
                S.load(y2, ee);
                u32 T0[16], T1[16], U0[16], U1[16];
                fp2_mul_wide(T0, T1, x, y);
                fp2_mul_wide(U0, U1, x2, y2);
                fp_redc_lazy(r.c0, T0);
                fp_redc_lazy(r.c1, T1);
                fp_redc_lazy(r2.c0, U0);
                fp_redc_lazy(r2.c1, U1);
                fp_canon(r.c0, 0u);
                fp_canon(r.c1, 0u);
                fp_canon(r2.c0, 0u);
                fp_canon(r2.c1, 0u);
                S.store(d, r);
                S.store(d == NS - 1 ? 0 : d + 1, r2);
            } else if (op == 2u) {
                S.load(x, a);
                u32 T0[16], T1[16];
                fp2_sqr_wide(T0, T1, x);
                fp_redc_lazy(r.c0, T0);
                fp_redc_lazy(r.c1, T1);
                fp_canon(r.c0, 0u);
                fp_canon(r.c1, 0u);
                S.store(d, r);
            } else {
                S.load(x, a);
                S.load(y, b);
                fp2_add(r, x, y);
                S.store(d, r);
            }
            ins = nxt;
        }
    }
    Fp2 v;
    S.load(v, 0);
    u32 x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= v.c0[i] ^ v.c1[i];
    if (x == 0x12345678u) args.out[blockIdx.x * T + threadIdx.x] = x;
}

static std::vector<u64> make_prog(int kind, int ns, int n) {
    std::vector<u64> w;
    u32 st = 12345u;
    auto rnd = [&]() { st = st * 1664525u + 1013904223u; return st >> 8; };
    for (int i = 0; i < n; i++) {
        u32 op;
        if (kind == 0) op = 1; else if (kind == 1) op = 2; else if (kind == 2) op = 3; else if (kind == 3) op = 4; else if (kind == 4) op = 5; else if (kind == 6) op = 6;
        else { u32 r = rnd() % 100; op = r < 52 ? 1 : r < 79 ? 2 : 3; }   // program-like mix: 4038 MUL, 2091 SQR, ~1600 linear
        u32 d = i % ns, a = rnd() % ns, b = rnd() % ns, c = rnd() % ns, e = rnd() % ns;
        w.push_back((u64)op | ((u64)d << 8) | ((u64)a << 16) | ((u64)b << 24) | ((u64)c << 32) | ((u64)e << 40));
    }
    for (int i = 0; i < 8; i++) w.push_back(0);
    return w;
}

template <int T, int MINB, int NS>
static void run_cfg(const char* cfg) {
    const char* names[] = {"mul", "sqr", "add", "mul_il", "mul_pre", "mix", "mul2"};
    const double macs[] = {336, 272, 0, 336, 336, 0.52 * 336 + 0.27 * 272, 672};
    cudaFuncSetAttribute(vmk<T, MINB, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, NS * 64 * T);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, vmk<T, MINB, NS>, T, NS * 64 * T);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, vmk<T, MINB, NS>);
    printf("# %s: T=%d slots/thread=%d regs=%d local=%zu B blocks/SM=%d (warps/SM %d)\n", cfg, T, NS, fa.numRegs, (size_t)fa.localSizeBytes, nb, nb * T / 32);
    u32* out;
    cudaMalloc(&out, 1 << 24);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int kind = 0; kind < 7; kind++) {
        const int nins = 2048, reps = 4;
        std::vector<u64> w = make_prog(kind, NS, nins);
        u64* dp;
        cudaMalloc(&dp, w.size() * 8);
        cudaMemcpy(dp, w.data(), w.size() * 8, cudaMemcpyHostToDevice);
        Args a{dp, (u32)nins, (u32)reps, out};
        float best = 1e9;
        for (int it = 0; it < 4; it++) {
            cudaEventRecord(e0);
            vmk<T, MINB, NS><<<148 * nb, T, NS * 64 * T>>>(a);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it) best = ms < best ? ms : best;
        }
        double warp_ops = (double)148 * nb * (T / 32) * nins * reps;
        double cyc = best * 1e-3 * 1.965e9 / (warp_ops / (148 * 4));
        double pipe = macs[kind] * 4;
        printf("%-8s %8.3f ms  %8.1f SMSP-cycles per warp-op (32 Fq2 ops)", names[kind], best, cyc);
        if (pipe > 0) printf("   pipe min %.0f: %.1f %%", pipe, 100.0 * pipe / cyc);
        printf("\n");
        cudaFree(dp);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("status %s\n", cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    run_cfg<64, 4, 14>("8 warps/SM");
    run_cfg<64, 6, 9>("12 warps/SM");
    return 0;
}
