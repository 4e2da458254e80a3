// The Fq2 sequencer: one thread = one pairing, all threads run the same straight-line program.
//
// Why a sequencer instead of one giant inlined kernel (B200-first reasoning):
//   * a fused pairing is ~27 000 Fq2-level operations; inlined that is >10^7 SASS instructions, far
//     beyond the 32 KB L1.5 / ~40 KB L2 instruction caches.  Here every Fq2 operation exists once
//     (~25 KB of SASS in total) and stays resident in the instruction cache;
//   * the per-pairing state (f: 384 B, R: 192 B, lines, Karatsuba temporaries) exceeds the register
//     file at any useful occupancy, and registers cannot be indexed dynamically.  The state lives in
//     shared memory as 64-byte Fq2 slots laid out [slot][quad][thread] so that every access is a
//     conflict-free LDS.128/STS.128, and registers hold only the operands of the running operation;
//   * control flow is warp- and grid-uniform (the program counter is the same for every thread), so
//     there is no divergence and the instruction word is fetched with a broadcast load.
#pragma once
#include "fp2.cuh"
#include "microcode_ops.h"

#define BNP_NARR 5
#define BNP_MAX_CONST 128

struct VmArgs {
    const u64* prog;        // instruction words (device global memory)
    u64* arr[BNP_NARR];     // SoA arrays: [K][4][stride] u64 (ids in microcode/isa.py)
    uint4* scratch;         // [n_scratch][4][total_threads] uint4
    u32 n;                  // elements to process
    u32 stride;             // elements per limb row of the arrays (>= n)
    u32* counter;           // work counter (zeroed before launch): warps claim 32-element chunks
};

__device__ __constant__ u32 BNP_CONSTS[BNP_MAX_CONST][16];

template <int T>
struct Slots {
    uint4* base;  // shared memory, already offset by threadIdx.x
    __device__ __forceinline__ void load(Fp2& r, u32 s) const {
        const uint4* p = base + s * (4 * T);
        uint4 q0 = p[0], q1 = p[T], q2 = p[2 * T], q3 = p[3 * T];
        r.c0[0] = q0.x; r.c0[1] = q0.y; r.c0[2] = q0.z; r.c0[3] = q0.w;
        r.c0[4] = q1.x; r.c0[5] = q1.y; r.c0[6] = q1.z; r.c0[7] = q1.w;
        r.c1[0] = q2.x; r.c1[1] = q2.y; r.c1[2] = q2.z; r.c1[3] = q2.w;
        r.c1[4] = q3.x; r.c1[5] = q3.y; r.c1[6] = q3.z; r.c1[7] = q3.w;
    }
    __device__ __forceinline__ void load_half(u32* r, u32 s, u32 half) const {
        const uint4* p = base + s * (4 * T) + half * (2 * T);
        uint4 q0 = p[0], q1 = p[T];
        r[0] = q0.x; r[1] = q0.y; r[2] = q0.z; r[3] = q0.w;
        r[4] = q1.x; r[5] = q1.y; r[6] = q1.z; r[7] = q1.w;
    }
    __device__ __forceinline__ void store(u32 s, const Fp2& r) const {
        uint4* p = base + s * (4 * T);
        p[0] = make_uint4(r.c0[0], r.c0[1], r.c0[2], r.c0[3]);
        p[T] = make_uint4(r.c0[4], r.c0[5], r.c0[6], r.c0[7]);
        p[2 * T] = make_uint4(r.c1[0], r.c1[1], r.c1[2], r.c1[3]);
        p[3 * T] = make_uint4(r.c1[4], r.c1[5], r.c1[6], r.c1[7]);
    }
};

// one Fq (4 x u64 limbs, stride n) of element e
__device__ __forceinline__ void ldg_fp(u32* r, const u64* arr, u32 f, u32 n, u32 e) {
    const u64* p = arr + (size_t)f * 4 * n + e;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        u64 v = __ldg(p + (size_t)j * n);
        r[2 * j] = (u32)v;
        r[2 * j + 1] = (u32)(v >> 32);
    }
}

__device__ __forceinline__ void stg_fp(u64* arr, u32 f, u32 n, u32 e, const u32* r) {
    u64* p = arr + (size_t)f * 4 * n + e;
#pragma unroll
    for (int j = 0; j < 4; j++) p[(size_t)j * n] = (u64)r[2 * j] | ((u64)r[2 * j + 1] << 32);
}

// K * p for K = 0 .. BNP_LIN_MAX_K (9 limbs each): the offset that makes a LIN accumulator non-negative
__device__ __constant__ u32 BNP_KP[BNP_LIN_MAX_K + 1][9];

// LIN: d = sum_i diag(m0,m1) x_i + xi * sum_j diag(m0,m1) x_j, accumulated lazily, reduced once.
// `pc` points at the first term word; returns the advanced pointer.
template <int T>
__device__ __forceinline__ const u64* vm_lin(const Slots<T>& S, const u64* pc, u32 d, u32 nterms, u32 K) {
    u32 A0[9], A1[9], X0[9], X1[9];
#pragma unroll
    for (int i = 0; i < 9; i++) A0[i] = A1[i] = X0[i] = X1[i] = 0u;
    bool any_xi = false;
    u64 w = 0;
#pragma unroll 1
    for (u32 j = 0; j < nterms; j++) {
        if ((j & 1u) == 0u) w = __ldg(pc++);
        const u32 t = (j & 1u) ? (u32)(w >> 32) : (u32)w;
        const u32 slot = t & 0xffu;
        const bool xi = (t >> 8) & BNP_LIN_XI;
        const int m0 = (int)(signed char)((t >> 16) & 0xffu);
        const int m1 = (int)(signed char)(t >> 24);
        Fp2 x;
        S.load(x, slot);
        if (xi) {
            any_xi = true;
            if (m0) acc9_term(X0, m0, x.c0);
            if (m1) acc9_term(X1, m1, x.c1);
        } else {
            if (m0) acc9_term(A0, m0, x.c0);
            if (m1) acc9_term(A1, m1, x.c1);
        }
    }
    if (any_xi) {
        // (A0, A1) += xi * (X0, X1) = (9 X0 - X1, X0 + 9 X1)
        u32 n0[9], n1[9];
        acc9_times9(n0, X0);
        acc9_times9(n1, X1);
        acc9_addsub(A0, n0, 0u);
        acc9_addsub(A0, X1, 1u);
        acc9_addsub(A1, X0, 0u);
        acc9_addsub(A1, n1, 0u);
    }
    u32 kp[9];
#pragma unroll
    for (int i = 0; i < 9; i++) kp[i] = BNP_KP[K][i];
    acc9_addsub(A0, kp, 0u);
    acc9_addsub(A1, kp, 0u);
    Fp2 r;
    fp_reduce_lazy(r.c0, A0);
    fp_reduce_lazy(r.c1, A1);
    S.store(d, r);
    return pc;
}

template <int T>
__global__ void __launch_bounds__(T) bnp_vm_kernel(VmArgs args) {
    extern __shared__ uint4 bnp_smem[];
    Slots<T> S;
    S.base = bnp_smem + threadIdx.x;
    const u32 total = gridDim.x * T;
    const u32 gtid = blockIdx.x * T + threadIdx.x;
    uint4* scr = args.scratch + gtid;
    const u32 n = args.n, stride = args.stride;

    // Persistent warps: each warp claims the next chunk of 32 elements when it finishes one.  A pairing
    // is ~10 ms of warp time and a 2^16 batch is only ~2 chunks per resident warp, so a static
    // grid-stride split would leave the last round badly unbalanced.
    const u32 lane = threadIdx.x & 31u;
    for (;;) {
        u32 chunk = 0;
        if (lane == 0) chunk = atomicAdd(args.counter, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        const u32 base = chunk * 32u;
        if (base >= n) break;
        const u32 e_raw = base + lane;
        const bool active = e_raw < n;
        const u32 e = active ? e_raw : n - 1;  // idle lanes shadow the last element and never store
        const u64* pc = args.prog;
        u64 ins = __ldg(pc++);
        for (;;) {
            const u32 lo = (u32)ins, hi = (u32)(ins >> 32);
            const u32 op = lo & 0xffu, d = (lo >> 8) & 0xffu, a = (lo >> 16) & 0xffu, b = lo >> 24;
            const u32 c = hi & 0xffu, ee = (hi >> 8) & 0xffu, imm = hi >> 16;
            if (op == BNP_OP_END) break;
            if (op == BNP_OP_LIN) {
                pc = vm_lin<T>(S, pc, d, a, imm);
                ins = __ldg(pc++);
                continue;
            }
            const u64 nxt = __ldg(pc++);  // prefetch (every program ends with END followed by padding)
            Fp2 x, y, r;
            switch (op) {
                case BNP_OP_MUL: {
                    S.load(x, a);
                    S.load(y, c);
                    if (imm) {  // Karatsuba operands: (a +- b) * (c +- e), sums kept lazy (< 2p)
                        Fp2 t;
                        if (imm & BNP_MUL_B) {
                            S.load(t, b);
                            if (imm & BNP_MUL_BNEG) fp2_sub_lazy(x, x, t); else fp2_add_lazy(x, x, t);
                        }
                        if (imm & BNP_MUL_E) {
                            S.load(t, ee);
                            if (imm & BNP_MUL_ENEG) fp2_sub_lazy(y, y, t); else fp2_add_lazy(y, y, t);
                        }
                    }
                    fp2_mul(r, x, y, imm != 0);
                    S.store(d, r);
                    break;
                }
                case BNP_OP_SQR:
                    S.load(x, a);
                    if (imm & BNP_MUL_B) {
                        S.load(y, b);
                        if (imm & BNP_MUL_BNEG) fp2_sub(x, x, y); else fp2_add(x, x, y);
                    }
                    fp2_sqr(r, x);
                    S.store(d, r);
                    break;
                case BNP_OP_MULFP: {
                    u32 s[8];
                    S.load(x, a);
                    S.load_half(s, b, imm & 1u);
                    fp2_mul_fp(r, x, s);
                    S.store(d, r);
                    break;
                }
                case BNP_OP_LDC:
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        r.c0[i] = BNP_CONSTS[imm][i];
                        r.c1[i] = BNP_CONSTS[imm][8 + i];
                    }
                    S.store(d, r);
                    break;
                case BNP_OP_LDG:
                    ldg_fp(r.c0, args.arr[imm], a, stride, e);
                    ldg_fp(r.c1, args.arr[imm], b, stride, e);
                    S.store(d, r);
                    break;
                case BNP_OP_STG:
                    S.load(x, a);
                    if (active) {
                        stg_fp(args.arr[imm], d, stride, e, x.c0);
                        stg_fp(args.arr[imm], b, stride, e, x.c1);
                    }
                    break;
                case BNP_OP_SPILL: {
                    const uint4* p = S.base + a * (4 * T);
                    uint4* q = scr + (size_t)imm * 4 * total;
                    q[0] = p[0];
                    q[total] = p[T];
                    q[2 * (size_t)total] = p[2 * T];
                    q[3 * (size_t)total] = p[3 * T];
                    break;
                }
                case BNP_OP_FILL: {
                    uint4* p = S.base + d * (4 * T);
                    const uint4* q = scr + (size_t)imm * 4 * total;
                    p[0] = q[0];
                    p[T] = q[total];
                    p[2 * T] = q[2 * (size_t)total];
                    p[3 * T] = q[3 * (size_t)total];
                    break;
                }
                case BNP_OP_INV:
                    S.load(x, a);
                    fp2_inv(r, x);
                    S.store(d, r);
                    break;
                case BNP_OP_ADD:
                    S.load(x, a);
                    S.load(y, b);
                    fp2_add(r, x, y);
                    S.store(d, r);
                    break;
                case BNP_OP_SUB:
                    S.load(x, a);
                    S.load(y, b);
                    fp2_sub(r, x, y);
                    S.store(d, r);
                    break;
                case BNP_OP_DBL:
                    S.load(x, a);
                    fp2_add(r, x, x);
                    S.store(d, r);
                    break;
                case BNP_OP_NEG:
                    S.load(x, a);
                    fp2_neg(r, x);
                    S.store(d, r);
                    break;
                case BNP_OP_CONJ:
                    S.load(x, a);
                    fp2_conj(r, x);
                    S.store(d, r);
                    break;
                case BNP_OP_MULXI:
                    S.load(x, a);
                    fp2_mul_xi(r, x);
                    S.store(d, r);
                    break;
                default:
                    break;
            }
            ins = nxt;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Roofline denominators, measured on the box: dependency-light loops of the two multiply forms.
//   bnp_imad_wide_peak_kernel : IMAD.WIDE.U32[.X] in 4-deep carry chains - the exact idiom of fp.cuh
//                               (32 MACs per thread per iteration, registers only)
//   bnp_imad_lo_peak_kernel   : plain 32-bit IMAD (informational: the pipe's nominal integer rate)
// SASS-checked: the loop bodies are 32 IMAD.WIDE.U32[.X] / 32 IMAD and nothing else but the loop counter.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bnp_imad_wide_peak_kernel(u64* out, u32 iters, u32 seed) {
    u32 a = seed * 2654435761u + threadIdx.x, b = a ^ 0x9e3779b9u;
    u64 acc[8];
#pragma unroll
    for (int c = 0; c < 8; c++) acc[c] = a + c;
#pragma unroll 1
    for (u32 it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 8; c += 4)
                asm volatile(
                    "{ .reg .u32 l0,h0,l1,h1,l2,h2,l3,h3;\n\t"
                    "mov.b64 {l0,h0}, %0; mov.b64 {l1,h1}, %1; mov.b64 {l2,h2}, %2; mov.b64 {l3,h3}, %3;\n\t"
                    "mad.lo.cc.u32 l0, %4, %5, l0; madc.hi.cc.u32 h0, %4, %5, h0;\n\t"
                    "madc.lo.cc.u32 l1, %4, %5, l1; madc.hi.cc.u32 h1, %4, %5, h1;\n\t"
                    "madc.lo.cc.u32 l2, %4, %5, l2; madc.hi.cc.u32 h2, %4, %5, h2;\n\t"
                    "madc.lo.cc.u32 l3, %4, %5, l3; madc.hi.u32 h3, %4, %5, h3;\n\t"
                    "mov.b64 %0, {l0,h0}; mov.b64 %1, {l1,h1}; mov.b64 %2, {l2,h2}; mov.b64 %3, {l3,h3}; }"
                    : "+l"(acc[c]), "+l"(acc[c + 1]), "+l"(acc[c + 2]), "+l"(acc[c + 3])
                    : "r"(a), "r"(b));
    }
    u64 x = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) x ^= acc[c];
    if (x == 0x1234567812345678ull) out[blockIdx.x * blockDim.x + threadIdx.x] = x;  // keeps the loop alive
}

__global__ void __launch_bounds__(256) bnp_imad_lo_peak_kernel(u64* out, u32 iters, u32 seed) {
    u32 a = seed * 2654435761u + threadIdx.x, b = a ^ 0x9e3779b9u;
    u32 acc[8];
#pragma unroll
    for (int c = 0; c < 8; c++) acc[c] = a + c;
#pragma unroll 1
    for (u32 it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 8; c++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[c]) : "r"(a), "r"(b));
    }
    u32 x = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) x ^= acc[c];
    if (x == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
