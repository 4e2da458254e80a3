// The Fq2 sequencer: ONE THREAD = ONE PAIRING, every thread of the grid runs the same straight-line program.
//
// Why a sequencer instead of one giant inlined kernel (B200-first reasoning):
//   * a fused pairing is ~27 000 Fq2-level operations; inlined that is >10^7 SASS instructions, far
//     beyond the instruction caches (L0 ~6 KB per scheduler, L1.5 32 KB per SM).  Here every Fq2 operation
//     exists once and the whole hot path stays inside the L1.5;
//   * the per-pairing state (f: 384 B, R: 192 B, lines, Karatsuba temporaries) exceeds the register file at any
//     useful occupancy, and registers cannot be indexed dynamically.  The state lives in shared memory as
//     64-byte Fq2 slots laid out [slot][quad][thread] (every access a conflict-free LDS.128 / STS.128);
//     registers hold only the operands of the running operation;
//   * control flow is grid-uniform: the program counter is the same for every thread, every branch of a handler
//     is decided by instruction fields, so there is no divergence and no per-lane select anywhere.
//
// Why one thread per pairing (round 2; round 1 shipped a two-lanes-per-pairing component split):
// measured on B200 with tools/mb/mb2.cu (profiles/mb_r2_solo_probe.txt), a lane that runs a WHOLE Fq2 product
// (Karatsuba: 3 wide products, 2 reductions, 336 MACs) from shared-memory slots sustains 79-82 % of the IMAD.WIDE
// pipe at 8 warps per SM and gains nothing from 12 or 14 warps, while the component split paid 4 wide products
// per Fq2 product plus a partner negate/select and topped out at 72 % of the pipe on executed MACs (62 % on
// algorithmic ones).  Per Fq2 product the thread-per-pairing form issues ~620 instructions for 32 pairings, the
// split ~870 for 16 + 16 lanes.  Occupancy is then set by shared memory alone (14 slots x 64 B x 256 threads).
#pragma once
#include "fp2.cuh"
#include "microcode_ops.h"

#define BNP_NARR 6
#define BNP_MAX_CONST 128
#ifndef BNP_MAX_PHASES
#define BNP_MAX_PHASES 16
#endif
#define BNP_CHUNK 32  // pairings per warp-task

struct VmArgs {
    const u64* prog[BNP_MAX_PHASES];  // instruction words of each phase (device global memory)
    u64* arr[BNP_NARR];     // SoA arrays: [K][4][stride] u64 (ids in microcode/isa.py); arr[5] = phase state
    uint4* scratch;         // [n_scratch][4][total_threads] uint4
    u32 n;                  // elements to process
    u32 stride;             // elements per limb row of the arrays (>= n)
    u32* counter;           // work counter (zeroed before launch): warps claim (phase, 32-element chunk) tasks
    u32* progress;          // per chunk: number of completed phases (zeroed before launch; unused when n_phases == 1)
    u32 n_phases;
    u32 aux_bcast;          // the AUX array holds one element that every element of the batch reads (stride 1, index 0)
    u32 tmem_cols;          // TMEM columns this block allocates (a power of two >= 32)
    u32 tmem_cols_per_warp; // 8 * number of slots: the column range of one warp
};

__device__ __constant__ u32 BNP_CONSTS[BNP_MAX_CONST][16];

// The slot file: every 64-byte Fq2 slot is SPLIT between shared memory and TENSOR MEMORY.
//
// This kernel issues no tcgen05.mma, so the SM's 256 KB of TMEM would sit idle while shared memory - the only other
// indexable on-chip store - caps the kernel at 9 slots per pairing and makes every second sequencer instruction a
// spill or a re-load.  tcgen05.ld / tcgen05.st in the 32x32b shape give every thread of a warp its own TMEM lane and
// N consecutive 32-bit columns per access, addressed by a register: a per-thread indexable scratchpad.  A warp reaches
// the lane quarter 32 * (warp % 4); the warps of a block that share a quarter take consecutive column ranges.
//
// Layout: limbs 0-3 of both components live in shared memory, [slot][component][thread] as uint4 (conflict-free
// LDS.128 / STS.128), limbs 4-7 in tensor memory, 8 columns per slot (c0 limbs 4-7, then c1 limbs 4-7).  Every access
// is the same straight-line pair of instructions - no second access path, no branch, no flag (measured first: slots
// living EITHER in shared memory OR in tensor memory, picked by a compare per access, cost 70 cycles of branch latency
// per access, 8 % of the kernel; profiles/experiments_r2.txt).  One 384-thread block per SM then holds 18 slots per
// pairing instead of 9, and a fused pairing shrinks from 18 400 to 9 300 sequencer instructions (spills and re-loads
// from 9 600 to 1 000); a 512-thread block holds 14 slots at 16 warps per SM.
#ifndef BNP_UNIFORM_DECODE
#define BNP_UNIFORM_DECODE 1
#endif
#ifndef BNP_ST_WAIT_LATE
#define BNP_ST_WAIT_LATE 0
#endif
#if BNP_ST_WAIT_LATE   // experiment: wait for outstanding TMEM stores in front of the next TMEM load instead of behind the store
#define BNP_ST_WAIT_PREFIX "tcgen05.wait::st.sync.aligned;\n\t"
#define BNP_ST_WAIT_SUFFIX ""
#else
#define BNP_ST_WAIT_PREFIX ""
#define BNP_ST_WAIT_SUFFIX "\n\ttcgen05.wait::st.sync.aligned;"
#endif
template <int T>
struct Slots {
    uint4* base;  // shared memory: this thread's entry of slot 0, component 0
    u32 tbase;    // tensor memory: this warp's lane quarter, first column of its range (slot s: tbase + 8 s)

    // A load is ISSUED (both halves in flight) and later WAITED for, so that the operands of one instruction share one
    // tcgen05.wait::ld.  `wait*` carries the real wait instruction for every TMEM load issued so far; `dep*` emits
    // nothing and only ties further registers to the statement order (volatile statements keep their order), so that
    // no use of them is scheduled before the wait.
    // one Fq by its index h = 2 * slot + half (the form LIN entries carry): one multiplication per address
    __device__ __forceinline__ void issue_fq(u32* r, u32 h) const {
        const uint4 q = base[h * T];
        r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w;
        asm volatile(BNP_ST_WAIT_PREFIX "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tbase + h * 4u));
    }
    __device__ __forceinline__ void issue_half(u32* r, u32 s, u32 half) const {
        const uint4 q = base[s * (2 * T) + half * T];
        r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w;
        asm volatile(BNP_ST_WAIT_PREFIX "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tbase + s * 8u + half * 4u));
    }
    __device__ __forceinline__ void issue(Fp2& r, u32 s) const {
        const uint4* p = base + s * (2 * T);
        const uint4 q0 = p[0], q1 = p[T];
        r.c0[0] = q0.x; r.c0[1] = q0.y; r.c0[2] = q0.z; r.c0[3] = q0.w;
        r.c1[0] = q1.x; r.c1[1] = q1.y; r.c1[2] = q1.z; r.c1[3] = q1.w;
        asm volatile(BNP_ST_WAIT_PREFIX "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r.c0[4]), "=r"(r.c0[5]), "=r"(r.c0[6]), "=r"(r.c0[7]), "=r"(r.c1[4]), "=r"(r.c1[5]),
                       "=r"(r.c1[6]), "=r"(r.c1[7])
                     : "r"(tbase + s * 8u));
    }
    __device__ __forceinline__ static void wait_half(u32* r) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
    }
    __device__ __forceinline__ static void dep_half(u32* r) {
        asm volatile("" : "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
    }
    __device__ __forceinline__ static void wait(Fp2& r) {
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+r"(r.c0[4]), "+r"(r.c0[5]), "+r"(r.c0[6]), "+r"(r.c0[7]), "+r"(r.c1[4]), "+r"(r.c1[5]),
                       "+r"(r.c1[6]), "+r"(r.c1[7]));
    }
    __device__ __forceinline__ static void dep(Fp2& r) {
        asm volatile(""
                     : "+r"(r.c0[4]), "+r"(r.c0[5]), "+r"(r.c0[6]), "+r"(r.c0[7]), "+r"(r.c1[4]), "+r"(r.c1[5]),
                       "+r"(r.c1[6]), "+r"(r.c1[7]));
    }
    __device__ __forceinline__ void load_half(u32* r, u32 s, u32 half) const {
        issue_half(r, s, half);
        wait_half(r);
    }
    __device__ __forceinline__ void load(Fp2& r, u32 s) const {
        issue(r, s);
        wait(r);
    }
    // (the asm statements are volatile: a load of a slot is never moved across a store)
    __device__ __forceinline__ void store(u32 s, const Fp2& r) const {
        uint4* p = base + s * (2 * T);
        p[0] = make_uint4(r.c0[0], r.c0[1], r.c0[2], r.c0[3]);
        p[T] = make_uint4(r.c1[0], r.c1[1], r.c1[2], r.c1[3]);
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};" BNP_ST_WAIT_SUFFIX
            :
            : "r"(r.c0[4]), "r"(r.c0[5]), "r"(r.c0[6]), "r"(r.c0[7]), "r"(r.c1[4]), "r"(r.c1[5]), "r"(r.c1[6]),
              "r"(r.c1[7]), "r"(tbase + s * 8u)
            : "memory");
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

// ---------------------------------------------------------------------------------------------
// LIN engine:  out_c = sum_j mult_j * (neg_j ? p - z_j : z_j)  <  1024 p  for c = 0, 1, each reduced once by a
// quotient estimate.  An entry is one pair of 4-deep IMAD.WIDE chains into 64-bit-column accumulators (E: even limb
// positions, O: odd).  Entries are 16 bits, [2 * slot + half : 9][neg:1][mult:6], and come in (component 0, component 1)
// pairs, one pair per 32-bit word; the two components are independent chains the scheduler interleaves.  A pair is
// fetched one iteration ahead of its use (the loop is software-pipelined by hand); multiplier 0 pads the shorter
// list.  Every test below is warp-uniform.
// ---------------------------------------------------------------------------------------------
#if BNP_UNIFORM_DECODE
#define BNP_UNI(x) __reduce_or_sync(0xffffffffu, (x))
#else
#define BNP_UNI(x) (x)
#endif
#define BNP_LIN_FETCH(PE, ZA, ZB, TT)                                         \
    {                                                                         \
        TT = BNP_UNI(__ldg(PE));                                              \
        S.issue_fq(ZA, TT & 0x1ffu);                                          \
        S.issue_fq(ZB, (TT >> 16) & 0x1ffu);                                  \
        S.wait_half(ZA);                                                      \
        S.dep_half(ZB);                                                       \
    }
#define BNP_LIN_ACC(ZA, ZB, TT)                                               \
    {                                                                         \
        if (TT & 0x00000200u) fp_p_minus(ZA, ZA);                             \
        if (TT & 0x02000000u) fp_p_minus(ZB, ZB);                             \
        const u32 ma_ = (TT >> 10) & 63u, mb_ = TT >> 26;                     \
        chain_acc64(E0, ma_, ZA[0], ZA[2], ZA[4], ZA[6]);                     \
        chain_acc64(O0, ma_, ZA[1], ZA[3], ZA[5], ZA[7]);                     \
        chain_acc64(E1, mb_, ZB[0], ZB[2], ZB[4], ZB[6]);                     \
        chain_acc64(O1, mb_, ZB[1], ZB[3], ZB[5], ZB[7]);                     \
    }
// first pair of a LIN: the accumulators are written, not accumulated into (no zero-initialisation)
#define BNP_LIN_ACC_FIRST(ZA, ZB, TT)                                         \
    {                                                                         \
        if (TT & 0x00000200u) fp_p_minus(ZA, ZA);                             \
        if (TT & 0x02000000u) fp_p_minus(ZB, ZB);                             \
        const u32 ma_ = (TT >> 10) & 63u, mb_ = TT >> 26;                     \
        _Pragma("unroll") for (int c_ = 0; c_ < 4; c_++) {                    \
            E0[c_] = (u64)ma_ * ZA[2 * c_];                                   \
            O0[c_] = (u64)ma_ * ZA[2 * c_ + 1];                               \
            E1[c_] = (u64)mb_ * ZB[2 * c_];                                   \
            O1[c_] = (u64)mb_ * ZB[2 * c_ + 1];                               \
        }                                                                     \
        E0[4] = O0[4] = E1[4] = O1[4] = 0ull;                                 \
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
__device__ __forceinline__ void vm_lin(const Slots<T>& S, Fp2& out, u32 n, const u64* more) {
    const u32* ents = (const u32*)more;  // pair j is the j-th 32-bit word of the entry list
    u64 E0[5], O0[5], E1[5], O1[5];
#pragma unroll
    for (int i = 0; i < 5; i++) E0[i] = O0[i] = E1[i] = O1[i] = 0ull;
#pragma unroll 1
    for (u32 j = 0; j < n; j++) {
        u32 za[8], zb[8], ta;
        BNP_LIN_FETCH(ents + j, za, zb, ta);
        BNP_LIN_ACC(za, zb, ta);
    }
    u32 v0[9], v1[9];
    lin_merge(v0, E0, O0);
    lin_merge(v1, E1, O1);
    fp_reduce_lazy(out.c0, v0);
    fp_reduce_lazy(out.c1, v1);
}

// ---------------------------------------------------------------------------------------------
// Product class (MUL / SQR / MULFP), one shared tail:
//   (T0, T1) = the two wide components;  T += 2^256 * hi terms;  r' = canon(redc(T));  [S[d] = r'];  [S[d2] = LIN(r', ...)]
// On entry `pc` points at the word after the instruction; on exit `ins` holds the next instruction
// and `pc` points past it.
// ---------------------------------------------------------------------------------------------
// raw operand fetch of a product-class instruction (no arithmetic): x = S[a]; y = S[c] (MUL) or the Fq half of
// S[b] (MULFP, into y.c0); tb = S[b], te = S[e] when the pre-addition flags ask for them
template <int T>
__device__ __forceinline__ void vm_product_fetch(const Slots<T>& S, u32 op, u32 a, u32 b, u32 c, u32 ee, u32 imm, Fp2& x,
                                                 Fp2& y, Fp2& tb, Fp2& te) {
    S.issue(x, a);
    if (op == BNP_OP_MUL) {
        S.issue(y, c);
        if (imm & BNP_MUL_B) S.issue(tb, b);
        if (imm & BNP_MUL_E) S.issue(te, ee);
        S.wait(x);
        S.dep(y);
        if (imm & BNP_MUL_B) S.dep(tb);
        if (imm & BNP_MUL_E) S.dep(te);
    } else if (op == BNP_OP_SQR) {
        if (imm & BNP_MUL_B) S.issue(tb, b);
        S.wait(x);
        if (imm & BNP_MUL_B) S.dep(tb);
    } else {
        S.issue_half(y.c0, b, (imm & BNP_MULFP_HALF) ? 1u : 0u);
        S.wait(x);
        S.dep_half(y.c0);
    }
}

template <int T>
__device__ __forceinline__ void vm_product(const Slots<T>& S, u32 op, const u64*& pc, u64& ins, const u64 w0, u32 d, u32 a,
                                           u32 b, u32 c, u32 ee, u32 imm) {
    const u64* p0 = pc;
    const u64 w1 = __ldg(p0 + 1);  // w0: extension word (or the next instruction), w1: the word after it
    // Every product-class opcode starts with the same TWO independent wide products A = u0 * v0, B = u1 * v1
    // (one copy of that code in the instruction stream: the hot path has to stay inside the 32 KB L1.5):
    //   MUL:    A = x0 y0, B = x1 y1, then C = (x0 + x1)(y0 + y1);  T0 = A - B (+ p 2^256 if negative), T1 = C - A - B
    //   SQR:    T0 = (x0 + x1)(x0 - x1), T1 = (2 x0) x1
    //   MULFP:  T0 = x0 s, T1 = x1 s
    u32 T0[16], T1[16];
    u32 u0[8], v0[8], u1[8], v1[8], sx[8], sy[8];
    Fp2 x, y, tb, te;
    vm_product_fetch<T>(S, op, a, b, c, ee, imm, x, y, tb, te);
    if (op == BNP_OP_MUL) {
        if (op == BNP_OP_MUL && (imm & (BNP_MUL_B | BNP_MUL_E))) {  // Karatsuba-level operands: (a +- b) * (c +- e), sums kept lazy (< 2p)
            if (imm & BNP_MUL_B) {
                if (imm & BNP_MUL_BNEG) fp2_sub_lazy(x, x, tb); else fp2_add_lazy(x, x, tb);
                if (imm & BNP_MUL_BCANON) {
                    fp_cond_sub_p(x.c0);
                    fp_cond_sub_p(x.c1);
                }
            }
            if (imm & BNP_MUL_E) {
                if (imm & BNP_MUL_ENEG) fp2_sub_lazy(y, y, te); else fp2_add_lazy(y, y, te);
            }
        }
        add8(sx, x.c0, x.c1);  // < 4p < 2^256
        add8(sy, y.c0, y.c1);
#pragma unroll
        for (int i = 0; i < 8; i++) { u0[i] = x.c0[i]; v0[i] = y.c0[i]; u1[i] = x.c1[i]; v1[i] = y.c1[i]; }
    } else if (op == BNP_OP_SQR) {
        if (imm & BNP_MUL_B) {
            if (imm & BNP_MUL_BNEG) fp2_sub(x, x, tb); else fp2_add(x, x, tb);
        }
        add8(u0, x.c0, x.c1);
        fp_sub(v0, x.c0, x.c1);
        add8(u1, x.c0, x.c0);
#pragma unroll
        for (int i = 0; i < 8; i++) { v1[i] = x.c1[i]; sx[i] = sy[i] = 0u; }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) { u0[i] = x.c0[i]; u1[i] = x.c1[i]; v0[i] = v1[i] = y.c0[i]; sx[i] = sy[i] = 0u; }
    }
    fp_mul_wide(T0, u0, v0);
    fp_mul_wide(T1, u1, v1);
    if (op == BNP_OP_MUL) {
        u32 P2[16];
        fp_mul_wide(P2, sx, sy);
        sub16(P2, P2, T0);
        sub16(P2, P2, T1);
        const u32 borrow = sub16(T0, T0, T1);
        add_p_masked(T0 + 8, borrow);
#pragma unroll
        for (int i = 0; i < 16; i++) T1[i] = P2[i];
    }
    u32 np = 0, d2 = 0;  // np: entry pairs of the post LIN
    bool store_r = true;
    if (imm & BNP_MUL_EXT) {
#if BNP_UNIFORM_DECODE
        const u32 xl = __reduce_or_sync(0xffffffffu, (u32)w0), xh = __reduce_or_sync(0xffffffffu, (u32)(w0 >> 32));
#else
        const u32 xl = (u32)w0, xh = (u32)(w0 >> 32);
#endif
        d2 = xl & 0xffu;
        const u32 hflags = xh & 0xffu;
        np = (xh >> 8) & 0xffu;
        store_r = np == 0u || (hflags & BNP_EXT_STORE_R);
        const u32 n_hi = hflags & 3u;
#pragma unroll 1
        for (u32 i = 0; i < n_hi; i++) {
            Fp2 h;
            S.load(h, (xl >> (8u * (i + 1u))) & 0xffu);
            if (hflags & (4u << i)) {
                fp_p_minus(h.c0, h.c0);
                fp_p_minus(h.c1, h.c1);
            }
            wide_add_hi(T0, h.c0);
            wide_add_hi(T1, h.c1);
        }
    }
    Fp2 r;
    fp_redc_lazy(r.c0, T0);
    fp_redc_lazy(r.c1, T1);
    // canonicalisation: the common case (both components below 2p) is ONE uniform test away from the final conditional
    // subtraction - four separately guarded ladder steps were four reconvergence regions per product
    const u32 cl = (imm >> BNP_MUL_CANON_SHIFT) & 15u;
    if (cl) {
        fp_canon_upper(r.c0, cl & 3u);
        fp_canon_upper(r.c1, cl >> 2);
    }
    fp_cond_sub_p(r.c0);
    fp_cond_sub_p(r.c1);
    if (!(imm & BNP_MUL_EXT)) {
        S.store(d, r);
        ins = w0;
        pc = p0 + 1;
        return;
    }
    if (np) {
        // Post stage: park r' where the entries expect it, then run the LIN engine on the entry list.
        const u32 nw = (np + 1u) >> 1;
        const u64 nx = __ldg(p0 + 1 + nw);
        S.store(store_r ? d : d2, r);
        Fp2 o;
        vm_lin<T>(S, o, np, p0 + 1);
        S.store(d2, o);
        ins = nx;
        pc = p0 + 2 + nw;
        return;
    }
    S.store(d, r);
    ins = w1;
    pc = p0 + 2;
}

#ifndef BNP_MINB
// Resident blocks of 64 threads per SM the register allocation must allow.  Measured (profiles/slots_sweep_r2_solo.txt):
// 6 blocks = 12 warps per SM with 9 slots per pairing beats 4 blocks = 8 warps with 14 slots by 4.6 % although the
// smaller slot file costs three times the spill traffic and 1500 stand-alone linear instructions - a third warp per
// scheduler covers more of the carry-chain latency than the spills cost.  12 warps: up to 170 registers.
#define BNP_MINB 6
#endif

template <int T>
#ifdef BNP_MAXNREG   // experiments: an explicit register cap instead of the launch bounds
__global__ void __maxnreg__(BNP_MAXNREG) bnp_vm_kernel(VmArgs args) {
#else
__global__ void __launch_bounds__(T, (BNP_MINB * 64) / T > 0 ? (BNP_MINB * 64) / T : 1) bnp_vm_kernel(VmArgs args) {
#endif
    extern __shared__ uint4 bnp_smem[];
    __shared__ u32 tmem_base;
    const u32 lane = threadIdx.x & 31u;
    Slots<T> S;
    S.base = bnp_smem + threadIdx.x;
    // Tensor memory: warp 0 allocates the block's columns, every warp derives the address of its own lane quarter and
    // column range.
    if (threadIdx.x < 32u) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (u32)__cvta_generic_to_shared(&tmem_base)),
                     "r"(args.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const u32 warp = threadIdx.x >> 5;
        S.tbase = tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * args.tmem_cols_per_warp;
#if BNP_UNIFORM_DECODE
        S.tbase = __reduce_or_sync(0xffffffffu, S.tbase);  // the same in every lane: keep it in a uniform register
#endif
    }
    const u32 total = gridDim.x * T;
    const u32 gtid = blockIdx.x * T + threadIdx.x;
    uint4* scr = args.scratch + gtid;
    const u32 n = args.n, stride = args.stride;

    // Persistent warps: each claims the next task when it finishes one.  A task is one phase of the program over one
    // chunk of 32 elements, handed out breadth-first (every chunk's phase 0, then every chunk's phase 1, ...), so
    // that only the last phase of a batch runs on a partly filled machine.  Phase p of a chunk waits for phase p-1
    // of the same chunk, which was claimed earlier by a warp that never waits on anything later - so the wait cannot
    // deadlock and is almost never taken.
    const u32 n_chunks = (n + (BNP_CHUNK - 1u)) / BNP_CHUNK;
    const u32 n_tasks = n_chunks * args.n_phases;
    for (;;) {
        u32 task = 0;
        if (lane == 0) task = atomicAdd(args.counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= n_tasks) break;
        const u32 phase = task / n_chunks;
        const u32 chunk = task - phase * n_chunks;
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
        const u32 e_raw = chunk * BNP_CHUNK + lane;
        const bool active = e_raw < n;
        const u32 e = active ? e_raw : n - 1;  // idle lanes shadow the last element and never store
        const u64* pc = args.prog[0];
#pragma unroll
        for (int k = 1; k < BNP_MAX_PHASES; k++)
            if (phase == (u32)k) pc = args.prog[k];
        u64 ins = __ldg(pc++);
        bool running = true;
        while (running) {
#if BNP_UNIFORM_DECODE
            // Every lane holds the same instruction word, but the compiler cannot know: a warp-wide OR (REDUX) returns it in
            // a UNIFORM register, so that field extraction runs on the uniform datapath, branches on instruction fields
            // need no reconvergence barriers (BSSY / BSYNC) and tensor-memory addresses need no R2UR.
            const u32 lo = __reduce_or_sync(0xffffffffu, (u32)ins), hi = __reduce_or_sync(0xffffffffu, (u32)(ins >> 32));
#else
            const u32 lo = (u32)ins, hi = (u32)(ins >> 32);
#endif
            const u32 op = lo & 0xffu, d = (lo >> 8) & 0xffu, a = (lo >> 16) & 0xffu, b = lo >> 24;
            const u32 c = hi & 0xffu, ee = (hi >> 8) & 0xffu, imm = hi >> 16;
            // the word after the instruction: its extension word, its first entry word, or the next instruction
            // (every program ends with END followed by padding)
            const u64 w0 = __ldg(pc);
            Fp2 x, y, r;
            // one dense switch: a jump table instead of a ladder of compares (the sequencer runs ~20 000
            // instructions per pairing; the ladder was 10 % of all stall samples)
            // product-class opcodes (1, 2, 3) are two thirds of all instructions: one test, ahead of the switch
            static_assert(BNP_OP_MUL == 1 && BNP_OP_SQR == 2 && BNP_OP_MULFP == 3, "product-class opcodes must be 1, 2, 3");
            if (op - 1u < 3u) {
                vm_product<T>(S, op, pc, ins, w0, d, a, b, c, ee, imm);
                continue;
            }
            switch (op) {
                case BNP_OP_END:
                    running = false;
                    break;
                case BNP_OP_LIN: {  // d = LIN(slots), a = number of entry pairs
                    const u32 nw = (a + 1u) >> 1;
                    const u64 nx = __ldg(pc + nw);
                    vm_lin<T>(S, r, a, pc);
                    S.store(d, r);
                    ins = nx;
                    pc += nw + 1;
                    break;
                }
                case BNP_OP_LDC:
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        r.c0[i] = BNP_CONSTS[imm][i];
                        r.c1[i] = BNP_CONSTS[imm][8 + i];
                    }
                    S.store(d, r);
                    ins = w0;
                    pc++;
                    break;
                case BNP_OP_LDG: {
                    // imm = array id | page << 8 (Fq index = page * 256 + field byte).  The AUX array may be ONE element
                    // shared by the whole batch (prepared G2 points of a verifying key): every lane reads element 0.
                    const u32 arr_id = imm & 0xffu, pg = (imm >> 8) << 8;
                    const bool bc = arr_id == BNP_ARR_AUX && args.aux_bcast;
                    ldg_fp(r.c0, args.arr[arr_id], pg + a, bc ? 1u : stride, bc ? 0u : e);
                    ldg_fp(r.c1, args.arr[arr_id], pg + b, bc ? 1u : stride, bc ? 0u : e);
                    S.store(d, r);
                    ins = w0;
                    pc++;
                    break;
                }
                case BNP_OP_STG: {
                    S.load(x, a);
                    const u32 pg = (imm >> 8) << 8;
                    if (active) {
                        stg_fp(args.arr[imm & 0xffu], pg + d, stride, e, x.c0);
                        stg_fp(args.arr[imm & 0xffu], pg + b, stride, e, x.c1);
                    }
                    ins = w0;
                    pc++;
                    break;
                }
                case BNP_OP_SPILL: {
                    S.load(x, a);
                    uint4* q = scr + (size_t)imm * 4 * total;
                    q[0] = make_uint4(x.c0[0], x.c0[1], x.c0[2], x.c0[3]);
                    q[total] = make_uint4(x.c0[4], x.c0[5], x.c0[6], x.c0[7]);
                    q[2 * (size_t)total] = make_uint4(x.c1[0], x.c1[1], x.c1[2], x.c1[3]);
                    q[3 * (size_t)total] = make_uint4(x.c1[4], x.c1[5], x.c1[6], x.c1[7]);
                    ins = w0;
                    pc++;
                    break;
                }
                case BNP_OP_FILL: {
                    const uint4* q = scr + (size_t)imm * 4 * total;
                    // read-once data: bypass L1 (what little L1 the slots leave holds the instruction words)
                    const uint4 q0 = __ldcg(q), q1 = __ldcg(q + total), q2 = __ldcg(q + 2 * (size_t)total),
                                q3 = __ldcg(q + 3 * (size_t)total);
                    r.c0[0] = q0.x; r.c0[1] = q0.y; r.c0[2] = q0.z; r.c0[3] = q0.w;
                    r.c0[4] = q1.x; r.c0[5] = q1.y; r.c0[6] = q1.z; r.c0[7] = q1.w;
                    r.c1[0] = q2.x; r.c1[1] = q2.y; r.c1[2] = q2.z; r.c1[3] = q2.w;
                    r.c1[4] = q3.x; r.c1[5] = q3.y; r.c1[6] = q3.z; r.c1[7] = q3.w;
                    S.store(d, r);
                    ins = w0;
                    pc++;
                    break;
                }
                case BNP_OP_INV:  // once or twice per program
                    S.load(x, a);
                    fp2_inv(r, x);
                    S.store(d, r);
                    ins = w0;
                    pc++;
                    break;
                case BNP_OP_ADD:
                    S.load(x, a);
                    S.load(y, b);
                    fp2_add(r, x, y);
                    S.store(d, r);
                    ins = w0;
                    pc++;
                    break;
                case BNP_OP_SUB:
                    S.load(x, a);
                    S.load(y, b);
                    fp2_sub(r, x, y);
                    S.store(d, r);
                    ins = w0;
                    pc++;
                    break;
                case BNP_OP_DBL:
                    S.load(x, a);
                    fp2_add(r, x, x);
                    S.store(d, r);
                    ins = w0;
                    pc++;
                    break;
                case BNP_OP_NEG:
                    S.load(x, a);
                    fp2_neg(r, x);
                    S.store(d, r);
                    ins = w0;
                    pc++;
                    break;
                case BNP_OP_CONJ:
                    S.load(x, a);
                    fp2_conj(r, x);
                    S.store(d, r);
                    ins = w0;
                    pc++;
                    break;
                case BNP_OP_MULXI:
                    S.load(x, a);
                    fp2_mul_xi(r, x);
                    S.store(d, r);
                    ins = w0;
                    pc++;
                    break;
                default:
                    ins = w0;
                    pc++;
                    break;
            }
        }
        if (args.n_phases > 1u) {  // publish: this chunk's state is complete up to and including `phase`
            __threadfence();
            __syncwarp();
            if (lane == 0)
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(args.progress + chunk), "r"(phase + 1u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32u)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(args.tmem_cols) : "memory");
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
