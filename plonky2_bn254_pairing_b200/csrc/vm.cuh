// The Fq2 sequencer, component-split: TWO lanes = one pairing, all lanes run the same straight-line program.
//
// Why a sequencer instead of one giant inlined kernel (B200-first reasoning):
//   * a fused pairing is ~27 000 Fq2-level operations; inlined that is >10^7 SASS instructions, far
//     beyond the instruction caches.  Here every Fq2 operation exists once and stays resident;
//   * the per-pairing state (f: 384 B, R: 192 B, lines, temporaries) exceeds the register file at any
//     useful occupancy, and registers cannot be indexed dynamically.  The state lives in shared memory
//     as Fq2 slots, registers hold only the operands of the running operation;
//   * control flow is warp- and grid-uniform (the program counter is the same for every thread).
//
// Why two lanes per pairing (measured, profiles/ncu_r1_attach_phases*.txt): with one thread per pairing the
// 896 B of slots per thread cap the SM at 8 warps (2 per scheduler) and the kernel sits at 43 % of the
// IMAD.WIDE pipe with 46 % of its stall samples in fixed-latency waits - there are not enough warps to cover the
// serial carry chains around the products.  Lane l < 16 of a warp computes component c0 of pairing l of the
// warp's 16-pairing chunk, lane l + 16 computes c1.  Per lane the slots are 32 bytes, so the same shared
// memory holds twice the warps (16 per SM, 4 per scheduler), the per-lane register footprint halves, and every
// Fq2 product becomes a two-term dot product with ONE reduction per lane:
//       c0 = x0*y0 + x1*(kp - y1)         c1 = x0*y1 + x1*y0
// (no Karatsuba: 4 wide products instead of 3, but no 512-bit subtractions, no borrow fix-up and a single
// merge-add; squarings and Fq-scalar products split evenly with no extra work).  A lane reads its partner's
// component straight from shared memory (layout [slot][quad][thread]: own, partner and broadcast reads are
// all conflict-free LDS.128); lanes only ever differ in addresses and select masks, never in control flow.
#pragma once
#include "fp2.cuh"
#include "microcode_ops.h"

#define BNP_NARR 6
#define BNP_MAX_CONST 128
#define BNP_MAX_PHASES 8
#define BNP_CHUNK 16  // pairings per warp

struct VmArgs {
    const u64* prog[BNP_MAX_PHASES];  // instruction words of each phase (device global memory)
    u64* arr[BNP_NARR];     // SoA arrays: [K][4][stride] u64 (ids in microcode/isa.py); arr[5] = phase state
    uint4* scratch;         // [n_scratch][2][total_threads] uint4
    u32 n;                  // elements to process
    u32 stride;             // elements per limb row of the arrays (>= n)
    u32* counter;           // work counter (zeroed before launch): warps claim (phase, 16-element chunk) tasks
    u32* progress;          // per chunk: number of completed phases (zeroed before launch; unused when n_phases == 1)
    u32 n_phases;
};

__device__ __constant__ u32 BNP_CONSTS[BNP_MAX_CONST][16];

// Shared-memory slots.  A lane holds ONE Fq component (32 B = two uint4) of every slot; the two lanes of a
// pairing are 16 apart in the warp.
template <int T>
struct Slots {
    uint4* h0;  // where component 0 / 1 of this lane's pairing lives (one of them is the lane's own)
    uint4* h1;
    uint4* own;
    uint4* oth;
    __device__ __forceinline__ const uint4* half(u32 k) const { return k ? h1 : h0; }
    __device__ __forceinline__ void load(u32* r, const uint4* base, u32 s) const {
        const uint4* p = base + s * (2 * T);
        uint4 q0 = p[0], q1 = p[T];
        r[0] = q0.x; r[1] = q0.y; r[2] = q0.z; r[3] = q0.w;
        r[4] = q1.x; r[5] = q1.y; r[6] = q1.z; r[7] = q1.w;
    }
    __device__ __forceinline__ void store(u32 s, const u32* r) const {
        uint4* p = own + s * (2 * T);
        p[0] = make_uint4(r[0], r[1], r[2], r[3]);
        p[T] = make_uint4(r[4], r[5], r[6], r[7]);
    }
};

// one Fq (4 x u64 limbs, stride n) of element e
__device__ __forceinline__ void ldg_fp(u32* r, const u64* arr, u32 f, u32 n, u32 e) {
    const u64* p = arr + (size_t)f * 4 * n + e;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        u64 v = __ldcg(p + (size_t)j * n);  // L2-coherent: the phase-state array is written by other SMs during the launch
        r[2 * j] = (u32)v;
        r[2 * j + 1] = (u32)(v >> 32);
    }
}

__device__ __forceinline__ void stg_fp(u64* arr, u32 f, u32 n, u32 e, const u32* r) {
    u64* p = arr + (size_t)f * 4 * n + e;
#pragma unroll
    for (int j = 0; j < 4; j++) p[(size_t)j * n] = (u64)r[2 * j] | ((u64)r[2 * j + 1] << 32);
}

// r = sel ? a : b, limb-wise (sel is per lane: never a branch)
__device__ __forceinline__ void sel8(u32* r, bool sel, const u32* a, const u32* b) {
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = sel ? a[i] : b[i];
}

// 512-bit r = a + b (no carry out by the callers' bounds)
__device__ __forceinline__ void add16(u32* r, const u32* a, const u32* b) {
    asm("add.cc.u32  %0, %16, %32;\n\t"
        "addc.cc.u32 %1, %17, %33;\n\t"
        "addc.cc.u32 %2, %18, %34;\n\t"
        "addc.cc.u32 %3, %19, %35;\n\t"
        "addc.cc.u32 %4, %20, %36;\n\t"
        "addc.cc.u32 %5, %21, %37;\n\t"
        "addc.cc.u32 %6, %22, %38;\n\t"
        "addc.cc.u32 %7, %23, %39;\n\t"
        "addc.cc.u32 %8, %24, %40;\n\t"
        "addc.cc.u32 %9, %25, %41;\n\t"
        "addc.cc.u32 %10, %26, %42;\n\t"
        "addc.cc.u32 %11, %27, %43;\n\t"
        "addc.cc.u32 %12, %28, %44;\n\t"
        "addc.cc.u32 %13, %29, %45;\n\t"
        "addc.cc.u32 %14, %30, %46;\n\t"
        "addc.u32    %15, %31, %47;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]),
          "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]),
          "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]), "r"(b[0]), "r"(b[1]),
          "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]), "r"(b[9]), "r"(b[10]),
          "r"(b[11]), "r"(b[12]), "r"(b[13]), "r"(b[14]), "r"(b[15]));
}

// r = k p - a, k = 1 or 2 (a <= k p); `two` is warp-uniform (an instruction flag), the constants are immediates
__device__ __forceinline__ void fp_kp_minus(u32* r, const u32* a, bool two) {
    if (two) {
        const u32 p2[8] = {0xb0f9fa8eu, 0x7841182du, 0xd0e3951au, 0x2f02d522u, 0x0302b0bbu, 0x70a08b6du, 0xc2634053u, 0x60c89ce5u};
        sub8(r, p2, a);
    } else {
        fp_p_minus(r, a);
    }
}

// ---------------------------------------------------------------------------------------------
// LIN engine, one output component per lane:
//     out = sum_j mult_j * (neg_j ? p - z_j : z_j)  <  1024 p,  reduced once by a quotient estimate.
// An entry is one pair of 4-deep IMAD.WIDE chains into 64-bit-column accumulators (E: even limb positions,
// O: odd).  Entries are 16 bits, [slot:8][half:1][neg:1][mult:6], and come in (component 0, component 1)
// pairs - a lane takes the entry of its own component, so the two lanes of a pairing differ only in the slot
// address, the negate select and the multiplier.  One pair per 32-bit word, fetched through the read-only cache
// one iteration ahead (the loop is software-pipelined by hand); the negate-select is skipped for pairs in which
// neither lane negates (a warp-uniform test).
// ---------------------------------------------------------------------------------------------
#define BNP_LIN_FETCH(J, Z, TT, NG)                                           \
    {                                                                         \
        const u32 t_ = __ldg(ents + (J));      /* one (component 0, component 1) pair, same word for every lane */ \
        NG = (t_ & 0x02000200u) != 0u;         /* warp-uniform: does either lane negate? */                        \
        TT = comp ? (t_ >> 16) : (t_ & 0xffffu);                              \
        S.load(Z, S.half((TT >> 8) & 1u), TT & 0xffu);                        \
    }
#define BNP_LIN_ACC(Z, TT, NG)                                                \
    {                                                                         \
        if (NG) {                                                             \
            u32 nz_[8];                                                       \
            fp_p_minus(nz_, Z);                                               \
            sel8(Z, (TT & 0x200u) != 0u, nz_, Z);                             \
        }                                                                     \
        chain_acc64(E, TT >> 10, Z[0], Z[2], Z[4], Z[6]);                     \
        chain_acc64(O, TT >> 10, Z[1], Z[3], Z[5], Z[7]);                     \
    }

// first entry of a LIN: the accumulators are written, not accumulated into (no zero-initialisation)
#define BNP_LIN_ACC_FIRST(Z, TT, NG)                                          \
    {                                                                         \
        if (NG) {                                                             \
            u32 nz_[8];                                                       \
            fp_p_minus(nz_, Z);                                               \
            sel8(Z, (TT & 0x200u) != 0u, nz_, Z);                             \
        }                                                                     \
        const u32 m_ = TT >> 10;                                              \
        _Pragma("unroll") for (int c_ = 0; c_ < 4; c_++) {                    \
            E[c_] = (u64)m_ * Z[2 * c_];                                      \
            O[c_] = (u64)m_ * Z[2 * c_ + 1];                                  \
        }                                                                     \
        E[4] = O[4] = 0ull;                                                   \
    }

// v = E + (O << 32): nine limbs (the total is below 2^264, so limb 9 of either part is zero)
__device__ __forceinline__ void lin_merge(u32* v, const u64* Ec, const u64* Oc) {
    u32 E[10], O[10];   // E, O: 64-bit columns at even / odd limb positions
#pragma unroll
    for (int c = 0; c < 5; c++) {
        E[2 * c] = (u32)Ec[c];
        E[2 * c + 1] = (u32)(Ec[c] >> 32);
        O[2 * c] = (u32)Oc[c];
        O[2 * c + 1] = (u32)(Oc[c] >> 32);
    }
    v[0] = E[0];
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, %23;"
        : "=&r"(v[1]), "=&r"(v[2]), "=&r"(v[3]), "=&r"(v[4]), "=&r"(v[5]), "=&r"(v[6]), "=&r"(v[7]), "=&r"(v[8])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(O[0]),
          "r"(O[1]), "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
}

template <int T>
__device__ __forceinline__ void vm_lin(const Slots<T>& S, bool comp, u32* out, u32 n, const u64* more) {
    const u32* ents = (const u32*)more;  // pair j is the j-th 32-bit word of the entry list
    u64 E[5], O[5];
#ifndef BNP_LIN_SINGLE_COPY
    u32 za[8], ya[8], ta, ua;
    bool na, nu;
    BNP_LIN_FETCH(0u, za, ta, na);
    if (1u < n) BNP_LIN_FETCH(1u, ya, ua, nu);
    BNP_LIN_ACC_FIRST(za, ta, na);
#pragma unroll 1
    for (u32 j = 1; j < n; j += 2u) {   // ya holds pair j
        if (j + 1u < n) BNP_LIN_FETCH(j + 1u, za, ta, na);
        BNP_LIN_ACC(ya, ua, nu);
        if (j + 1u >= n) break;
        if (j + 2u < n) BNP_LIN_FETCH(j + 2u, ya, ua, nu);
        BNP_LIN_ACC(za, ta, na);
    }
#else
    // one copy of the entry body: 56 instructions smaller, measured 0.5 % slower than the hand-pipelined form above
#pragma unroll
    for (int i = 0; i < 5; i++) E[i] = O[i] = 0ull;
#pragma unroll 1
    for (u32 j = 0; j < n; j++) {
        u32 za[8], ta;
        bool na;
        BNP_LIN_FETCH(j, za, ta, na);
        BNP_LIN_ACC(za, ta, na);
    }
#endif
    u32 v[9];
    lin_merge(v, E, O);
    fp_reduce_lazy(out, v);
}

// ---------------------------------------------------------------------------------------------
// Product class (MUL / SQR / MULFP), one shared body:
//   T = this lane's wide value;  T += 2^256 * hi terms;  r' = canon(redc(T));  [S[d] = r'];  [S[d2] = LIN(r', ...)]
// On entry `pc` points at the word after the instruction; on exit `ins` holds the next instruction
// and `pc` points past it.
// ---------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ void vm_product(const Slots<T>& S, bool comp, u32 op, const u64*& pc, u64& ins, u32 d, u32 a,
                                           u32 b, u32 c, u32 ee, u32 imm) {
    const u64* p0 = pc;
    const u64 w0 = __ldg(p0), w1 = __ldg(p0 + 1);  // extension word (or the next instruction), the word after it
    // Every product-class opcode is  T = u1 * v1 [+ u2 * v2]  followed by the same tail; only the operand
    // preparation differs, so the wide products, the reduction and the epilogue exist ONCE in the instruction
    // stream (the kernel is instruction-fetch sensitive: L0 is ~6 KB per scheduler, L1.5 32 KB per SM).
    u32 u1[8], v1[8], u2[8], v2[8];
    if (op == BNP_OP_MUL) {
        // lane c:  x0 * y[c]  +  x1 * (c ? y0 : k p - y1)
        S.load(u1, S.h0, a);
        S.load(u2, S.h1, a);
        S.load(v1, S.own, c);
        S.load(v2, S.oth, c);
        if (imm & (BNP_MUL_B | BNP_MUL_E)) {  // Karatsuba-level operands: (a +- b) * (c +- e), sums kept lazy (< 2p)
            u32 t[8];
            if (imm & BNP_MUL_B) {
                S.load(t, S.h0, b);
                if (imm & BNP_MUL_BNEG) fp_sub_lazy(u1, u1, t); else add8(u1, u1, t);
                S.load(t, S.h1, b);
                if (imm & BNP_MUL_BNEG) fp_sub_lazy(u2, u2, t); else add8(u2, u2, t);
                if (imm & BNP_MUL_BCANON) {
                    fp_cond_sub_p(u1);
                    fp_cond_sub_p(u2);
                }
            }
            if (imm & BNP_MUL_E) {
                S.load(t, S.own, ee);
                if (imm & BNP_MUL_ENEG) fp_sub_lazy(v1, v1, t); else add8(v1, v1, t);
                S.load(t, S.oth, ee);
                if (imm & BNP_MUL_ENEG) fp_sub_lazy(v2, v2, t); else add8(v2, v2, t);
            }
        }
        u32 nb[8];
        fp_kp_minus(nb, v2, (imm & BNP_MUL_E) != 0u);
        sel8(v2, comp, v2, nb);
    } else if (op == BNP_OP_SQR) {
        // lane 0: (x0 + x1) * (x0 - x1),  lane 1: (x0 + x0) * x1      (x canonical)
        u32 x0[8], x1[8];
        S.load(x0, S.h0, a);
        S.load(x1, S.h1, a);
        if (imm & BNP_MUL_B) {
            u32 t[8];
            S.load(t, S.h0, b);
            if (imm & BNP_MUL_BNEG) fp_sub(x0, x0, t); else fp_add(x0, x0, t);
            S.load(t, S.h1, b);
            if (imm & BNP_MUL_BNEG) fp_sub(x1, x1, t); else fp_add(x1, x1, t);
        }
        u32 s[8], dd[8];
        sel8(s, comp, x0, x1);
        add8(u1, x0, s);
        fp_sub(dd, x0, x1);
        sel8(v1, comp, x1, dd);
    } else {
        S.load(u1, S.own, a);
        S.load(v1, S.half(imm & BNP_MULFP_HALF), b);
    }
    u32 TT[16];
    if (op == BNP_OP_MUL)
        fp_mul2_wide(TT, u1, v1, u2, v2);  // both terms in one pair of column accumulators (v1, v2 <= 2p)
    else
        fp_mul_wide(TT, u1, v1);
    u32 np = 0, d2 = 0;  // np: entry pairs of the post LIN
    bool store_r = true;
    if (imm & BNP_MUL_EXT) {
        const u32 xl = (u32)w0, xh = (u32)(w0 >> 32);
        d2 = xl & 0xffu;
        const u32 hflags = xh & 0xffu;
        np = (xh >> 8) & 0xffu;
        store_r = np == 0u || (hflags & BNP_EXT_STORE_R);
        const u32 n_hi = hflags & 3u;
#pragma unroll 1
        for (u32 i = 0; i < n_hi; i++) {
            u32 h[8];
            S.load(h, S.own, (xl >> (8u * (i + 1u))) & 0xffu);
            if (hflags & (4u << i)) fp_p_minus(h, h);
            wide_add_hi(TT, h);
        }
    }
    u32 r[8];
    fp_redc_lazy(r, TT);
    // the two components of a product have the same bound, but take the larger level so the ladder is uniform
    const u32 l0 = (imm >> BNP_MUL_CANON_SHIFT) & 3u, l1 = (imm >> (BNP_MUL_CANON_SHIFT + 2)) & 3u;
    fp_canon(r, l0 > l1 ? l0 : l1);
    __syncwarp();  // every lane has read its operands (a destination may alias a source slot)
    if (!(imm & BNP_MUL_EXT)) {
        S.store(d, r);
        ins = w0;
        pc = p0 + 1;
        return;
    }
    if (np) {
        // Post stage: park r' where the entries expect it, then run the LIN engine on the entry list.
#ifdef BNP_POST_LIN_REDISPATCH
        // (hand the entry list to the LIN handler of the main loop as a synthetic "LIN d2, np" instruction: one copy
        //  of the engine in the instruction stream, one more decode per post stage - measured 1.1 % slower)
        S.store(store_r ? d : d2, r);
        ins = (u64)BNP_OP_LIN | ((u64)d2 << 8) | ((u64)np << 16);
        pc = p0 + 1;
        return;
#else
        const u32 nw = (np + 1u) >> 1;
        const u64 nx = __ldg(p0 + 1 + nw);
        S.store(store_r ? d : d2, r);
        __syncwarp();
        u32 o[8];
        vm_lin<T>(S, comp, o, np, p0 + 1);
        __syncwarp();
        S.store(d2, o);
        ins = nx;
        pc = p0 + 2 + nw;
        return;
#endif
    }
    S.store(d, r);
    ins = w1;
    pc = p0 + 2;
}

#ifndef BNP_LS_MIN
#define BNP_LS_MIN 128  // blocks of at least this many threads run in lockstep (see below)
#endif
#ifndef BNP_MINB
#define BNP_MINB 8  // resident blocks of 64 threads per SM the register allocation must allow (16 warps: <= 128 registers)
#endif

// LOCKSTEP (blocks of >= 256 threads): the warps of a block run the program instruction by instruction behind a
// block barrier.  The handlers are long straight-line code (a product body is ~7 KB against a ~6 KB L0
// instruction cache per scheduler and 32 KB L1.5 per SM): with free-running warps every warp streams its own
// copy of the code through the instruction caches and the kernel becomes instruction-fetch bound (measured:
// 20 % of all stall samples are "no instruction", 72 % of them on the first instruction of a 128-byte line).
// In lockstep the four warps of a scheduler execute the same lines at nearly the same time, so a line is
// fetched once per scheduler instead of once per warp.
template <int T>
__global__ void __launch_bounds__(T, (T == 96) ? 6 : (BNP_MINB * 64) / T) bnp_vm_kernel(VmArgs args) {
    constexpr bool LS = T >= BNP_LS_MIN;
    constexpr u32 WPB = T / 32;
    extern __shared__ uint4 bnp_smem[];
    __shared__ u32 s_task;
    const u32 lane = threadIdx.x & 31u;
    const u32 warp = threadIdx.x >> 5;
    const bool comp = (lane & 16u) != 0u;
    Slots<T> S;
    S.h0 = bnp_smem + (threadIdx.x & ~16u);
    S.h1 = bnp_smem + (threadIdx.x | 16u);
    S.own = bnp_smem + threadIdx.x;
    S.oth = bnp_smem + (threadIdx.x ^ 16u);
    const u32 total = gridDim.x * T;
    const u32 gtid = blockIdx.x * T + threadIdx.x;
    uint4* scr = args.scratch + gtid;
    const u32 n = args.n, stride = args.stride;

    // Persistent warps (blocks in lockstep mode): each claims the next task when it finishes one.  A task is one
    // phase of the program over one chunk of 16 elements (lockstep: over WPB consecutive chunks, one per warp),
    // handed out breadth-first (every chunk's phase 0, then every chunk's phase 1, ...), so that only the last
    // phase of a batch runs on a partly filled machine.  Phase p of a chunk waits for phase p-1 of the same
    // chunk, which was claimed earlier by a warp that never waits on anything later - so the wait cannot
    // deadlock and is almost never taken.
    const u32 n_chunks = (n + (BNP_CHUNK - 1u)) / BNP_CHUNK;
    const u32 n_groups = LS ? (n_chunks + WPB - 1u) / WPB : n_chunks;
    const u32 n_tasks = n_groups * args.n_phases;
    for (;;) {
        u32 task = 0;
        if (LS) {
            __syncthreads();
            if (threadIdx.x == 0) s_task = atomicAdd(args.counter, 1u);
            __syncthreads();
            task = s_task;
        } else {
            if (lane == 0) task = atomicAdd(args.counter, 1u);
            task = __shfl_sync(0xffffffffu, task, 0);
        }
        if (task >= n_tasks) break;
        const u32 phase = task / n_groups;
        const u32 grp = task - phase * n_groups;
        const u32 chunk_raw = LS ? grp * WPB + warp : grp;
        const bool chunk_ok = chunk_raw < n_chunks;
        const u32 chunk = chunk_ok ? chunk_raw : n_chunks - 1u;  // a warp without a chunk shadows the last one and never stores
        if (phase) {
            if (lane == 0) {
                u32 done;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(done) : "l"(args.progress + chunk) : "memory");
                    if (done >= phase) break;
                    __nanosleep(256);
                }
            }
            __syncwarp();
        }
        const u32 e_raw = chunk * BNP_CHUNK + (lane & 15u);
        const bool active = chunk_ok && e_raw < n;
        const u32 e = active ? e_raw : n - 1;  // idle lanes shadow the last element and never store
        const u64* pc = args.prog[0];
#pragma unroll
        for (int k = 1; k < BNP_MAX_PHASES; k++)
            if (phase == (u32)k) pc = args.prog[k];
        u64 ins = __ldg(pc++);
        for (;;) {
            const u32 lo = (u32)ins, hi = (u32)(ins >> 32);
            const u32 op = lo & 0xffu, d = (lo >> 8) & 0xffu, a = (lo >> 16) & 0xffu, b = lo >> 24;
            const u32 c = hi & 0xffu, ee = (hi >> 8) & 0xffu, imm = hi >> 16;
            if (op == BNP_OP_END) break;
            if (LS) __syncthreads(); else __syncwarp();  // the previous instruction's stores are visible to the partner lane
            // (A straight-line fast path for flag-free MUL/SQR was measured and removed: +10 KB of hot code pushed the
            //  working set past the 32 KB L1.5 instruction cache, -8 % free-running, no gain in lockstep.)
            if (op == BNP_OP_MUL || op == BNP_OP_SQR || op == BNP_OP_MULFP) {
                vm_product<T>(S, comp, op, pc, ins, d, a, b, c, ee, imm);
                continue;
            }
            if (op == BNP_OP_LIN) {  // d = LIN(slots), a = number of entry pairs
                const u32 nw = (a + 1u) >> 1;
                const u64 nx = __ldg(pc + nw);
                u32 o[8];
                vm_lin<T>(S, comp, o, a, pc);
                __syncwarp();
                S.store(d, o);
                ins = nx;
                pc += nw + 1;
                continue;
            }
            const u64 nxt = __ldg(pc++);  // prefetch (every program ends with END followed by padding)
            u32 x[8], y[8], r[8];
            switch (op) {
                case BNP_OP_LDC:
#pragma unroll
                    for (int i = 0; i < 8; i++) r[i] = comp ? BNP_CONSTS[imm][8 + i] : BNP_CONSTS[imm][i];
                    S.store(d, r);
                    break;
                case BNP_OP_LDG:
                    ldg_fp(r, args.arr[imm], comp ? b : a, stride, e);
                    S.store(d, r);
                    break;
                case BNP_OP_STG:
                    S.load(x, S.own, a);
                    if (active) stg_fp(args.arr[imm], comp ? b : d, stride, e, x);
                    break;
                case BNP_OP_SPILL: {
                    const uint4* p = S.own + a * (2 * T);
                    uint4* q = scr + (size_t)imm * 2 * total;
                    q[0] = p[0];
                    q[total] = p[T];
                    break;
                }
                case BNP_OP_FILL: {
                    uint4* p = S.own + d * (2 * T);
                    const uint4* q = scr + (size_t)imm * 2 * total;
                    // read-once data: bypass L1 (what little L1 the slots leave holds the instruction words)
                    p[0] = __ldcg(q);
                    p[T] = __ldcg(q + total);
                    break;
                }
                case BNP_OP_INV: {  // once or twice per program: both lanes run the whole Fq2 inversion
                    Fp2 v, w;
                    S.load(v.c0, S.h0, a);
                    S.load(v.c1, S.h1, a);
                    fp2_inv(w, v);
                    sel8(r, comp, w.c1, w.c0);
                    __syncwarp();
                    S.store(d, r);
                    break;
                }
                case BNP_OP_ADD:
                    S.load(x, S.own, a);
                    S.load(y, S.own, b);
                    fp_add(r, x, y);
                    S.store(d, r);
                    break;
                case BNP_OP_SUB:
                    S.load(x, S.own, a);
                    S.load(y, S.own, b);
                    fp_sub(r, x, y);
                    S.store(d, r);
                    break;
                case BNP_OP_DBL:
                    S.load(x, S.own, a);
                    fp_add(r, x, x);
                    S.store(d, r);
                    break;
                case BNP_OP_NEG:
                    S.load(x, S.own, a);
                    fp_neg(r, x);
                    S.store(d, r);
                    break;
                case BNP_OP_CONJ:
                    S.load(x, S.own, a);
                    fp_neg(y, x);
                    sel8(r, comp, y, x);
                    S.store(d, r);
                    break;
                case BNP_OP_MULXI: {  // (9 + u)(a0 + a1 u): lane 0 = 9 a0 + (p - a1), lane 1 = 9 a1 + a0
                    u32 v[9];
                    S.load(x, S.own, a);
                    S.load(y, S.oth, a);
                    fp_p_minus(r, y);
                    sel8(y, comp, y, r);
                    mul9_add(v, x, y);
                    fp_reduce_small(r, v);
                    __syncwarp();
                    S.store(d, r);
                    break;
                }
                default:
                    break;
            }
            ins = nxt;
        }
        if (args.n_phases > 1u) {  // publish: this chunk's state is complete up to and including `phase`
            __threadfence();
            __syncwarp();
            if (lane == 0 && chunk_ok)
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(args.progress + chunk), "r"(phase + 1u) : "memory");
        }
        __syncwarp();  // slots are reused by the next task
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
