// The Fq2 sequencer: every pairing owns one slot file in shared memory and TWO lanes of a warp ("units"); the
// whole grid runs the same straight-line program of BUNDLES - one Fq2 instruction per unit, same opcode.
//
// Why a sequencer instead of one giant inlined kernel (B200-first reasoning):
//   * a fused pairing is ~27 000 Fq2-level operations; inlined that is >10^7 SASS instructions, far
//     beyond the instruction caches (L0 ~6 KB per scheduler, L1.5 32 KB per SM).  Here every Fq2 operation
//     exists once and the hot path stays inside the L1.5;
//   * the per-pairing state (f: 384 B, R: 192 B, lines, Karatsuba temporaries) exceeds the register file at any
//     useful occupancy, and registers cannot be indexed dynamically.  The state lives in shared memory as
//     64-byte Fq2 slots laid out [slot][quad][pairing] (every access a conflict-free LDS.128 / STS.128);
//     registers hold only the operands of the running operation;
//   * control flow is grid-uniform: the program counter is the same for every thread and every branch of a
//     handler is decided by bundle fields that all lanes see.
//
// Why two units per pairing, each running a WHOLE Fq2 operation (round 2; measured, profiles/*_r2_*):
//   * a lane that runs a whole Karatsuba Fq2 product (3 wide products, 2 reductions, 336 MACs) issues ~620
//     instructions per product; round 1's component split (lane = one Fq component: 4 wide products per Fq2
//     product, partner negate/select) issued 2 x 435.  Thread-per-pairing cut a 2^16 launch from 16.4 G to 12.8 G
//     instructions - and ran no faster: the slot file (14 x 64 B) caps an SM at 256 pairings = 8 warps, each warp
//     is one serial dependency chain (5.3 cycles per issued instruction), and 2 warps per scheduler leave the
//     IMAD pipe idle during every carry-chain phase (issue slots 37 % used);
//   * so the parallelism has to come from INSIDE a pairing: the tower arithmetic is full of independent Fq2
//     operations (18 / 12 / 13 / 9 products in an Fq12 multiplication / squaring / sparse multiplication /
//     cyclotomic squaring, line evaluations next to the accumulator update).  microcode/sched.py pairs them at
//     build time; lane l < 16 of a warp runs unit A's instruction of pairing l, lane l + 16 unit B's, on the same
//     slots.  Same shared memory per pairing, twice the warps (16 per SM at <= 128 registers), the lean
//     per-lane instruction stream of the thread-per-pairing form.
// The units of a bundle differ only in operand fields, masks and selects - never in control flow.  Every store
// sits behind a warp barrier: both units have read their operands before either writes (a destination may
// reuse the slot of a source that dies in the bundle).
#pragma once
#include "fp2.cuh"
#include "microcode_ops.h"

#define BNP_NARR 6
#define BNP_MAX_CONST 128
#define BNP_MAX_PHASES 16
#define BNP_CHUNK 16  // pairings per warp-task (two lanes each)

struct VmArgs {
    const u64* prog[BNP_MAX_PHASES];  // bundles of each phase (device global memory, 16-byte aligned)
    u64* arr[BNP_NARR];     // SoA arrays: [K][4][stride] u64 (ids in microcode/isa.py); arr[5] = phase state
    uint4* scratch;         // [n_scratch][4][resident pairings] uint4
    u32 n;                  // elements to process
    u32 stride;             // elements per limb row of the arrays (>= n)
    u32* counter;           // work counter (zeroed before launch): warps claim (phase, 16-element chunk) tasks
    u32* progress;          // per chunk: number of completed phases (zeroed before launch; unused when n_phases == 1)
    u32 n_phases;
};

__device__ __constant__ u32 BNP_CONSTS[BNP_MAX_CONST][16];

// Shared-memory slots: [slot][quad][pairing] uint4; quads 0,1 = c0, quads 2,3 = c1.  P = pairings per block.
template <int T>
struct Slots {
    static constexpr int P = T / 2;
    uint4* base;  // already offset by the pairing's index in the block
    __device__ __forceinline__ void load_half(u32* r, u32 s, u32 half) const {
        const uint4* p = base + s * (4 * P) + half * (2 * P);
        uint4 q0 = p[0], q1 = p[P];
        r[0] = q0.x; r[1] = q0.y; r[2] = q0.z; r[3] = q0.w;
        r[4] = q1.x; r[5] = q1.y; r[6] = q1.z; r[7] = q1.w;
    }
    __device__ __forceinline__ void load(Fp2& r, u32 s) const {
        const uint4* p = base + s * (4 * P);
        uint4 q0 = p[0], q1 = p[P], q2 = p[2 * P], q3 = p[3 * P];
        r.c0[0] = q0.x; r.c0[1] = q0.y; r.c0[2] = q0.z; r.c0[3] = q0.w;
        r.c0[4] = q1.x; r.c0[5] = q1.y; r.c0[6] = q1.z; r.c0[7] = q1.w;
        r.c1[0] = q2.x; r.c1[1] = q2.y; r.c1[2] = q2.z; r.c1[3] = q2.w;
        r.c1[4] = q3.x; r.c1[5] = q3.y; r.c1[6] = q3.z; r.c1[7] = q3.w;
    }
    __device__ __forceinline__ void store(u32 s, const Fp2& r) const {
        uint4* p = base + s * (4 * P);
        p[0] = make_uint4(r.c0[0], r.c0[1], r.c0[2], r.c0[3]);
        p[P] = make_uint4(r.c0[4], r.c0[5], r.c0[6], r.c0[7]);
        p[2 * P] = make_uint4(r.c1[0], r.c1[1], r.c1[2], r.c1[3]);
        p[3 * P] = make_uint4(r.c1[4], r.c1[5], r.c1[6], r.c1[7]);
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

// What a unit adds for an optional, optionally negated canonical operand t:
//     t (present, positive)   p - t (present, negated)   0 (absent)
// any_neg / all_neg / any_absent describe the BUNDLE and are warp-uniform; my_* are this unit's own bits.  When the
// two units agree (the common case: the tower formulas are symmetric) there is no select and no mask.
__device__ __forceinline__ void prep_addend(u32* t, bool my_pres, bool my_neg, bool any_neg, bool all_neg, bool any_absent) {
    if (any_neg) {
        u32 nt[8];
        fp_p_minus(nt, t);
        if (all_neg) {
#pragma unroll
            for (int i = 0; i < 8; i++) t[i] = nt[i];
        } else {
            sel8(t, my_neg, nt, t);
        }
    }
    if (any_absent) {
        const u32 m = my_pres ? 0xffffffffu : 0u;
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] &= m;
    }
}

// ---------------------------------------------------------------------------------------------
// LIN engine:  out_c = sum_j mult_j * (neg_j ? p - z_j : z_j)  <  1024 p  for c = 0, 1, each reduced once by a
// quotient estimate.  An entry is one pair of 4-deep IMAD.WIDE chains into 64-bit-column accumulators (E: even limb
// positions, O: odd).  Entries are 16 bits, [slot:8][half:1][neg:1][mult:6]; one 64-bit word per entry index holds
// unit A's (c0, c1) pair in its low half and unit B's in its high half.  Multiplier 0 pads the shorter lists.
// ---------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void vm_lin(const Slots<T>& S, u32 unit, Fp2& out, u32 n, const u64* ents) {
    u64 E0[5], O0[5], E1[5], O1[5];
#pragma unroll
    for (int i = 0; i < 5; i++) E0[i] = O0[i] = E1[i] = O1[i] = 0ull;
#pragma unroll 1
    for (u32 j = 0; j < n; j++) {
        const u64 w = __ldg(ents + j);
        const u32 lo = (u32)w, hi = (u32)(w >> 32);
        const u32 t = unit ? hi : lo;
        const u32 any = lo | hi, all = lo & hi;
        u32 za[8], zb[8];
        S.load_half(za, t & 0xffu, (t >> 8) & 1u);
        S.load_half(zb, (t >> 16) & 0xffu, (t >> 24) & 1u);
        prep_addend(za, true, (t & 0x00000200u) != 0u, (any & 0x00000200u) != 0u, (all & 0x00000200u) != 0u, false);
        prep_addend(zb, true, (t & 0x02000000u) != 0u, (any & 0x02000000u) != 0u, (all & 0x02000000u) != 0u, false);
        const u32 ma = (t >> 10) & 63u, mb = t >> 26;
        chain_acc64(E0, ma, za[0], za[2], za[4], za[6]);
        chain_acc64(O0, ma, za[1], za[3], za[5], za[7]);
        chain_acc64(E1, mb, zb[0], zb[2], zb[4], zb[6]);
        chain_acc64(O1, mb, zb[1], zb[3], zb[5], zb[7]);
    }
    u32 v0[9], v1[9];
    lin_merge(v0, E0, O0);
    lin_merge(v1, E1, O1);
    fp_reduce_lazy(out.c0, v0);
    fp_reduce_lazy(out.c1, v1);
}

// ---------------------------------------------------------------------------------------------
// Product class (MUL / SQR / MULFP), one shared body:
//   (T0, T1) = the two wide components;  T += 2^256 * hi terms;  r' = canon(redc(T));  [S[d] = r'];  [S[d2] = LIN(r', ...)]
// `pc` points at the bundle; returns the address of the next bundle.
// ---------------------------------------------------------------------------------------------
template <int T>
__device__ __forceinline__ const u64* vm_product(const Slots<T>& S, u32 unit, bool act, u32 op, const u64* pc, u64 wA, u64 wB,
                                                 u64 mine) {
    const u32 lo = (u32)mine, hi = (u32)(mine >> 32);
    const u32 d = (lo >> 8) & 0xffu, a = (lo >> 16) & 0xffu, b = lo >> 24;
    const u32 c = hi & 0xffu, ee = (hi >> 8) & 0xffu, imm = hi >> 16;
    const u32 immA = (u32)(wA >> 48), immB = (u32)(wB >> 48);
    const u32 any = immA | immB, all = immA & immB;
    u64 xA = 0, xB = 0;
    if (any & BNP_MUL_EXT) {  // extension words of the two units
        xA = __ldg(pc + 2);
        xB = __ldg(pc + 3);
    }
    // Every product-class opcode starts with the same TWO independent wide products A = u0 * v0, B = u1 * v1
    // (one copy of that code in the instruction stream: the hot path has to stay inside the 32 KB L1.5):
    //   MUL:    A = x0 y0, B = x1 y1, then C = (x0 + x1)(y0 + y1);  T0 = A - B (+ p 2^256 if negative), T1 = C - A - B
    //   SQR:    T0 = (x0 + x1)(x0 - x1), T1 = (2 x0) x1
    //   MULFP:  T0 = x0 s, T1 = x1 s
    u32 T0[16], T1[16];
    u32 u0[8], v0[8], u1[8], v1[8], sx[8], sy[8];
    if (op == BNP_OP_MUL) {
        Fp2 x, y;
        S.load(x, a);
        S.load(y, c);
        // Karatsuba-level operands: (a +- b) * (c +- e), sums kept lazy (< 2p); a - b is computed as a + (p - b)
        if (any & BNP_MUL_B) {
            Fp2 t;
            S.load(t, b);
            const bool pres = (imm & BNP_MUL_B) != 0u, neg = (imm & BNP_MUL_BNEG) != 0u;
            const bool any_neg = (any & BNP_MUL_BNEG) != 0u, all_neg = (all & BNP_MUL_BNEG) != 0u, any_abs = !(all & BNP_MUL_B);
            prep_addend(t.c0, pres, neg, any_neg, all_neg, any_abs);
            prep_addend(t.c1, pres, neg, any_neg, all_neg, any_abs);
            add8(x.c0, x.c0, t.c0);
            add8(x.c1, x.c1, t.c1);
            if (any & BNP_MUL_BCANON) {
                fp_cond_sub_p(x.c0);
                fp_cond_sub_p(x.c1);
            }
        }
        if (any & BNP_MUL_E) {
            Fp2 t;
            S.load(t, ee);
            const bool pres = (imm & BNP_MUL_E) != 0u, neg = (imm & BNP_MUL_ENEG) != 0u;
            const bool any_neg = (any & BNP_MUL_ENEG) != 0u, all_neg = (all & BNP_MUL_ENEG) != 0u, any_abs = !(all & BNP_MUL_E);
            prep_addend(t.c0, pres, neg, any_neg, all_neg, any_abs);
            prep_addend(t.c1, pres, neg, any_neg, all_neg, any_abs);
            add8(y.c0, y.c0, t.c0);
            add8(y.c1, y.c1, t.c1);
        }
        add8(sx, x.c0, x.c1);  // < 4p < 2^256
        add8(sy, y.c0, y.c1);
#pragma unroll
        for (int i = 0; i < 8; i++) { u0[i] = x.c0[i]; v0[i] = y.c0[i]; u1[i] = x.c1[i]; v1[i] = y.c1[i]; }
    } else if (op == BNP_OP_SQR) {
        Fp2 x;
        S.load(x, a);
        if (any & BNP_MUL_B) {  // (a +- b)^2, the sum made canonical
            Fp2 t;
            S.load(t, b);
            const bool pres = (imm & BNP_MUL_B) != 0u, neg = (imm & BNP_MUL_BNEG) != 0u;
            const bool any_neg = (any & BNP_MUL_BNEG) != 0u, all_neg = (all & BNP_MUL_BNEG) != 0u, any_abs = !(all & BNP_MUL_B);
            prep_addend(t.c0, pres, neg, any_neg, all_neg, any_abs);
            prep_addend(t.c1, pres, neg, any_neg, all_neg, any_abs);
            add8(x.c0, x.c0, t.c0);
            add8(x.c1, x.c1, t.c1);
            fp_cond_sub_p(x.c0);
            fp_cond_sub_p(x.c1);
        }
        add8(u0, x.c0, x.c1);
        fp_sub(v0, x.c0, x.c1);
        add8(u1, x.c0, x.c0);
#pragma unroll
        for (int i = 0; i < 8; i++) { v1[i] = x.c1[i]; sx[i] = sy[i] = 0u; }
    } else {
        Fp2 x;
        S.load(x, a);
        S.load_half(v0, b, (imm & BNP_MULFP_HALF) ? 1u : 0u);
#pragma unroll
        for (int i = 0; i < 8; i++) { u0[i] = x.c0[i]; u1[i] = x.c1[i]; v1[i] = v0[i]; sx[i] = sy[i] = 0u; }
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
    const u64 xm = unit ? xB : xA;
    const u32 xl = (u32)xm, xh = (u32)(xm >> 32);
    u32 np = 0;
    if (any & BNP_MUL_EXT) {
        const u32 hfA = (u32)(xA >> 32), hfB = (u32)(xB >> 32);   // hflags in bits 0-7, present mask in bits 16-23
        const u32 hany = hfA | hfB, hall = hfA & hfB;
        np = (xh >> 8) & 0xffu;
        const u32 n_hi = xh & 3u;
#pragma unroll 1
        for (u32 i = 0; i < n_hi; i++) {
            Fp2 h;
            S.load(h, (xl >> (8u * (i + 1u))) & 0xffu);
            const bool pres = ((xh >> (16u + i)) & 1u) != 0u, neg = ((xh >> (2u + i)) & 1u) != 0u;
            const bool any_neg = ((hany >> (2u + i)) & 1u) != 0u, all_neg = ((hall >> (2u + i)) & 1u) != 0u;
            const bool any_abs = ((hall >> (16u + i)) & 1u) == 0u;
            prep_addend(h.c0, pres, neg, any_neg, all_neg, any_abs);
            prep_addend(h.c1, pres, neg, any_neg, all_neg, any_abs);
            wide_add_hi(T0, h.c0);
            wide_add_hi(T1, h.c1);
        }
    }
    Fp2 r;
    fp_redc_lazy(r.c0, T0);
    fp_redc_lazy(r.c1, T1);
    fp_canon(r.c0, (immA >> BNP_MUL_CANON_SHIFT) & 3u);
    fp_canon(r.c1, (immA >> (BNP_MUL_CANON_SHIFT + 2)) & 3u);
    __syncwarp();  // both units have read their operands (a destination may alias a source slot of either unit)
    if (!(any & BNP_MUL_EXT)) {
        if (act) S.store(d, r);
        return pc + 2;
    }
    const u32 d2 = xl & 0xffu;
    const u32 np_own = xh >> 24;
    const bool store_r = np_own == 0u || (xh & BNP_EXT_STORE_R);
    const u64* next = pc + 4 + np + (np & 1u);
    if (np) {
        // Post stage: park r' where the entries expect it, then run the LIN engine on the entry list.
        if (act) S.store(store_r ? d : d2, r);
        __syncwarp();
        Fp2 o;
        vm_lin<T>(S, unit, o, np, pc + 4);
        __syncwarp();
        if (act && np_own) S.store(d2, o);
        return next;
    }
    if (act) S.store(d, r);
    return next;
}

#ifndef BNP_MINB
#define BNP_MINB 8  // resident blocks of 64 threads per SM the register allocation must allow (16 warps: <= 128 registers)
#endif

template <int T>
__global__ void __launch_bounds__(T, (BNP_MINB * 64) / T) bnp_vm_kernel(VmArgs args) {
    extern __shared__ uint4 bnp_smem[];
    const u32 lane = threadIdx.x & 31u;
    const u32 unit = lane >> 4;
    const u32 pib = (threadIdx.x >> 5) * BNP_CHUNK + (lane & 15u);  // this pairing's index in the block
    Slots<T> S;
    S.base = bnp_smem + pib;
    const u32 totalp = gridDim.x * (T / 2);
    uint4* scr = args.scratch + (blockIdx.x * (T / 2) + pib);
    const u32 n = args.n, stride = args.stride;

    // Persistent warps: each claims the next task when it finishes one.  A task is one phase of the program over one
    // chunk of 16 elements, handed out breadth-first (every chunk's phase 0, then every chunk's phase 1, ...), so
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
        const u32 e_raw = chunk * BNP_CHUNK + (lane & 15u);
        const bool active = e_raw < n;
        const u32 e = active ? e_raw : n - 1;  // idle pairings shadow the last element and never store
        const u64* pc = args.prog[0];
#pragma unroll
        for (int k = 1; k < BNP_MAX_PHASES; k++)
            if (phase == (u32)k) pc = args.prog[k];
        for (;;) {
            const u64 wA = __ldg(pc), wB = __ldg(pc + 1);
            const u32 op = (u32)wA & 0xffu;
            if (op == BNP_OP_END) break;
            const bool act = !(unit && ((u32)wB & 0x80u));  // an idle unit B shadows unit A and stores nothing
            const u64 mine = unit ? wB : wA;
            __syncwarp();  // the previous bundle's stores are visible to the partner unit
            if (op == BNP_OP_MUL || op == BNP_OP_SQR || op == BNP_OP_MULFP) {
                pc = vm_product<T>(S, unit, act, op, pc, wA, wB, mine);
                continue;
            }
            const u32 lo = (u32)mine, hi = (u32)(mine >> 32);
            const u32 d = (lo >> 8) & 0xffu, a = (lo >> 16) & 0xffu, b = lo >> 24;
            const u32 imm = hi >> 16;
            if (op == BNP_OP_LIN) {  // d = LIN(slots), a = number of entry words
                Fp2 o;
                vm_lin<T>(S, unit, o, a, pc + 2);
                __syncwarp();
                if (act) S.store(d, o);
                pc += 2 + a + (a & 1u);
                continue;
            }
            pc += 2;
            Fp2 x, y, r;
            switch (op) {
                case BNP_OP_LDC:
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        r.c0[i] = BNP_CONSTS[imm][i];
                        r.c1[i] = BNP_CONSTS[imm][8 + i];
                    }
                    if (act) S.store(d, r);
                    break;
                case BNP_OP_LDG:
                    ldg_fp(r.c0, args.arr[imm], a, stride, e);
                    ldg_fp(r.c1, args.arr[imm], b, stride, e);
                    if (act) S.store(d, r);
                    break;
                case BNP_OP_STG:
                    S.load(x, a);
                    if (act && active) {
                        stg_fp(args.arr[imm], d, stride, e, x.c0);
                        stg_fp(args.arr[imm], b, stride, e, x.c1);
                    }
                    break;
                case BNP_OP_SPILL: {
                    const uint4* p = S.base + a * (4 * Slots<T>::P);
                    uint4* q = scr + (size_t)imm * 4 * totalp;
                    if (act) {
                        q[0] = p[0];
                        q[totalp] = p[Slots<T>::P];
                        q[2 * (size_t)totalp] = p[2 * Slots<T>::P];
                        q[3 * (size_t)totalp] = p[3 * Slots<T>::P];
                    }
                    break;
                }
                case BNP_OP_FILL: {
                    uint4* p = S.base + d * (4 * Slots<T>::P);
                    const uint4* q = scr + (size_t)imm * 4 * totalp;
                    // read-once data: bypass L1 (what little L1 the slots leave holds the instruction words)
                    const uint4 q0 = __ldcg(q), q1 = __ldcg(q + totalp), q2 = __ldcg(q + 2 * (size_t)totalp),
                                q3 = __ldcg(q + 3 * (size_t)totalp);
                    if (act) {
                        p[0] = q0;
                        p[Slots<T>::P] = q1;
                        p[2 * Slots<T>::P] = q2;
                        p[3 * Slots<T>::P] = q3;
                    }
                    break;
                }
                case BNP_OP_INV:  // once or twice per program, always alone in its bundle: unit B shadows unit A
                    S.load(x, a);
                    fp2_inv(r, x);
                    __syncwarp();
                    if (act) S.store(d, r);
                    break;
                case BNP_OP_ADD:
                    S.load(x, a);
                    S.load(y, b);
                    fp2_add(r, x, y);
                    __syncwarp();
                    if (act) S.store(d, r);
                    break;
                case BNP_OP_SUB:
                    S.load(x, a);
                    S.load(y, b);
                    fp2_sub(r, x, y);
                    __syncwarp();
                    if (act) S.store(d, r);
                    break;
                case BNP_OP_DBL:
                    S.load(x, a);
                    fp2_add(r, x, x);
                    __syncwarp();
                    if (act) S.store(d, r);
                    break;
                case BNP_OP_NEG:
                    S.load(x, a);
                    fp2_neg(r, x);
                    __syncwarp();
                    if (act) S.store(d, r);
                    break;
                case BNP_OP_CONJ:
                    S.load(x, a);
                    fp2_conj(r, x);
                    __syncwarp();
                    if (act) S.store(d, r);
                    break;
                case BNP_OP_MULXI:
                    S.load(x, a);
                    fp2_mul_xi(r, x);
                    __syncwarp();
                    if (act) S.store(d, r);
                    break;
                default:
                    break;
            }
        }
        if (args.n_phases > 1u) {  // publish: this chunk's state is complete up to and including `phase`
            __threadfence();
            __syncwarp();
            if (lane == 0)
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
