// BN254 base-field arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form with R = 2^256
// (bit-identical to ark-ff's `Fp<MontBackend<FqConfig,4>>` 4 x u64 limbs, little endian).
//
// Every 32x32->64 multiply-accumulate is written as a `mad.lo.cc` / `madc.hi.cc` PAIR into two
// adjacent registers so that ptxas fuses the pair into ONE `IMAD.WIDE.U32.X Rd, Pc, Ra, Rb, Rd, Pc`
// (carry in/out in a predicate): one FMA-pipe issue per MAC.  To make every pair register-aligned the
// running sum is kept in two accumulators: E holds 64-bit columns at even limb positions, O at odd
// positions (O[k] is limb k+1).  The value is E + (O << 32); they are merged once per product.
//
// Carry discipline: the PTX condition code never crosses an asm statement.
#pragma once
#include <stdint.h>

typedef uint32_t u32;
typedef uint64_t u64;

// p = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
#define BNP_P0 0xd87cfd47
#define BNP_P1 0x3c208c16
#define BNP_P2 0x6871ca8d
#define BNP_P3 0x97816a91
#define BNP_P4 0x8181585d
#define BNP_P5 0xb85045b6
#define BNP_P6 0xe131a029
#define BNP_P7 0x30644e72
#define BNP_N0INV 0xe4866389  // -p^-1 mod 2^32

#define BNP_STR2(x) #x
#define BNP_STR(x) BNP_STR2(x)

__device__ __constant__ u32 BNP_P[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                        0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};

// ---------------------------------------------------------------------------------------------
// column chains: x[B .. B+7] (+)= s * (q0, q1, q2, q3), four 64-bit columns with carry between them
// ---------------------------------------------------------------------------------------------

// all eight limbs fresh: four independent wide products (mul.wide, so that each is ONE IMAD.WIDE - written as
// mul.lo / mul.hi pairs ptxas fuses only some of them)
template <int B>
__device__ __forceinline__ void chain_fresh(u32* x, u32 s, u32 q0, u32 q1, u32 q2, u32 q3) {
    asm("{ .reg .u64 t0, t1, t2, t3;\n\t"
        "mul.wide.u32 t0, %8, %9;  mul.wide.u32 t1, %8, %10;\n\t"
        "mul.wide.u32 t2, %8, %11; mul.wide.u32 t3, %8, %12;\n\t"
        "mov.b64 {%0, %1}, t0; mov.b64 {%2, %3}, t1; mov.b64 {%4, %5}, t2; mov.b64 {%6, %7}, t3; }"
        : "=&r"(x[B]), "=&r"(x[B + 1]), "=&r"(x[B + 2]), "=&r"(x[B + 3]), "=&r"(x[B + 4]), "=&r"(x[B + 5]),
          "=&r"(x[B + 6]), "=&r"(x[B + 7])
        : "r"(s), "r"(q0), "r"(q1), "r"(q2), "r"(q3));
}

// limbs B..B+5 hold data, B+6 and B+7 are fresh (no carry out possible)
template <int B>
__device__ __forceinline__ void chain_top2(u32* x, u32 s, u32 q0, u32 q1, u32 q2, u32 q3) {
    asm("mad.lo.cc.u32  %0, %8, %9,  %0; madc.hi.cc.u32 %1, %8, %9,  %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %8, %11, %4; madc.hi.cc.u32 %5, %8, %11, %5;\n\t"
        "madc.lo.cc.u32 %6, %8, %12, 0;  madc.hi.u32    %7, %8, %12, 0;"
        : "+r"(x[B]), "+r"(x[B + 1]), "+r"(x[B + 2]), "+r"(x[B + 3]), "+r"(x[B + 4]), "+r"(x[B + 5]),
          "=&r"(x[B + 6]), "=&r"(x[B + 7])
        : "r"(s), "r"(q0), "r"(q1), "r"(q2), "r"(q3));
}

// limbs B..B+6 hold data (B+6 is a captured carry, 0 or 1), B+7 is fresh (no carry out possible)
template <int B>
__device__ __forceinline__ void chain_top1(u32* x, u32 s, u32 q0, u32 q1, u32 q2, u32 q3) {
    asm("mad.lo.cc.u32  %0, %8, %9,  %0; madc.hi.cc.u32 %1, %8, %9,  %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %8, %11, %4; madc.hi.cc.u32 %5, %8, %11, %5;\n\t"
        "madc.lo.cc.u32 %6, %8, %12, %6; madc.hi.u32    %7, %8, %12, 0;"
        : "+r"(x[B]), "+r"(x[B + 1]), "+r"(x[B + 2]), "+r"(x[B + 3]), "+r"(x[B + 4]), "+r"(x[B + 5]),
          "+r"(x[B + 6]), "=&r"(x[B + 7])
        : "r"(s), "r"(q0), "r"(q1), "r"(q2), "r"(q3));
}

// all eight limbs hold data; the carry out is captured into the fresh limb B+8
template <int B>
__device__ __forceinline__ void chain_full(u32* x, u32 s, u32 q0, u32 q1, u32 q2, u32 q3) {
    asm("mad.lo.cc.u32  %0, %9, %10, %0; madc.hi.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, %2; madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %12, %4; madc.hi.cc.u32 %5, %9, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %9, %13, %6; madc.hi.cc.u32 %7, %9, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(x[B]), "+r"(x[B + 1]), "+r"(x[B + 2]), "+r"(x[B + 3]), "+r"(x[B + 4]), "+r"(x[B + 5]),
          "+r"(x[B + 6]), "+r"(x[B + 7]), "=&r"(x[B + 8])
        : "r"(s), "r"(q0), "r"(q1), "r"(q2), "r"(q3));
}

// accumulating form for the LIN engine: x[B .. B+7] += s * (q0..q3), carry ADDED to x[B+8]
template <int B>
__device__ __forceinline__ void chain_acc(u32* x, u32 s, u32 q0, u32 q1, u32 q2, u32 q3) {
    asm("mad.lo.cc.u32  %0, %9, %10, %0; madc.hi.cc.u32 %1, %9, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %11, %2; madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %12, %4; madc.hi.cc.u32 %5, %9, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %9, %13, %6; madc.hi.cc.u32 %7, %9, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(x[B]), "+r"(x[B + 1]), "+r"(x[B + 2]), "+r"(x[B + 3]), "+r"(x[B + 4]), "+r"(x[B + 5]),
          "+r"(x[B + 6]), "+r"(x[B + 7]), "+r"(x[B + 8])
        : "r"(s), "r"(q0), "r"(q1), "r"(q2), "r"(q3));
}

// The same accumulation on 64-bit COLUMN variables: x[0..3] += s * (q0..q3), carry added to the low word of x[4].
// Typing the columns as u64 pins each one to an aligned register pair - inside a loop ptxas otherwise lets the
// two halves of a loop-carried column drift apart and moves them into a pair and back around every IMAD.WIDE
// (measured in the LIN loop: 30 moves per 8 multiply-accumulates).
__device__ __forceinline__ void chain_acc64(u64* x, u32 s, u32 q0, u32 q1, u32 q2, u32 q3) {
    asm("{ .reg .u32 l0, h0, l1, h1, l2, h2, l3, h3, l4, h4;\n\t"
        "mov.b64 {l0, h0}, %0; mov.b64 {l1, h1}, %1; mov.b64 {l2, h2}, %2; mov.b64 {l3, h3}, %3; mov.b64 {l4, h4}, %4;\n\t"
        "mad.lo.cc.u32  l0, %5, %6, l0; madc.hi.cc.u32 h0, %5, %6, h0;\n\t"
        "madc.lo.cc.u32 l1, %5, %7, l1; madc.hi.cc.u32 h1, %5, %7, h1;\n\t"
        "madc.lo.cc.u32 l2, %5, %8, l2; madc.hi.cc.u32 h2, %5, %8, h2;\n\t"
        "madc.lo.cc.u32 l3, %5, %9, l3; madc.hi.cc.u32 h3, %5, %9, h3;\n\t"
        "addc.u32 l4, l4, 0;\n\t"
        "mov.b64 %0, {l0, h0}; mov.b64 %1, {l1, h1}; mov.b64 %2, {l2, h2}; mov.b64 %3, {l3, h3}; mov.b64 %4, {l4, h4}; }"
        : "+l"(x[0]), "+l"(x[1]), "+l"(x[2]), "+l"(x[3]), "+l"(x[4])
        : "r"(s), "r"(q0), "r"(q1), "r"(q2), "r"(q3));
}

// ---------------------------------------------------------------------------------------------
// One row of a product or of a Montgomery reduction: two chains with the same multiplier s.
//   C chain (base D): C[D..D+5] hold data, C[D+6] and C[D+7] are fresh:   C[D..D+7]  = C[D..D+5] + s * (c0, c1, c2, c3)
//   A chain (base B): A[B..B+7] hold data:                                A[B..B+7] += s * (a0, a1, a2, a3)
// The A chain starts one limb below the C chain, so its carry out lands exactly on the limb position of
// C[D+7] and is added there.  C[D+7] = hi(s * c3) + (<= 1) has room for it whenever c3 is a TOP limb below
// 2^32 - 2: c3 is limb 7 of a value below 4p, or of p.  No carry is ever parked in a limb of its own, so no
// {carry, 0} register pair has to be built (measured before this form: one SEL + one IMAD.MOV per row,
// 10 % of all issued instructions were such IMAD.MOVs on the contended multiply pipe).
// ---------------------------------------------------------------------------------------------
template <int B, int D>
__device__ __forceinline__ void mul_row(u32* A, u32* C, u32 s, u32 a0, u32 a1, u32 a2, u32 a3, u32 c0, u32 c1, u32 c2,
                                        u32 c3) {
    asm("mad.lo.cc.u32  %0, %16, %21, %0;  madc.hi.cc.u32 %1, %16, %21, %1;\n\t"
        "madc.lo.cc.u32 %2, %16, %22, %2;  madc.hi.cc.u32 %3, %16, %22, %3;\n\t"
        "madc.lo.cc.u32 %4, %16, %23, %4;  madc.hi.cc.u32 %5, %16, %23, %5;\n\t"
        "madc.lo.cc.u32 %6, %16, %24, 0;   madc.hi.u32    %7, %16, %24, 0;\n\t"
        "mad.lo.cc.u32  %8,  %16, %17, %8;  madc.hi.cc.u32 %9,  %16, %17, %9;\n\t"
        "madc.lo.cc.u32 %10, %16, %18, %10; madc.hi.cc.u32 %11, %16, %18, %11;\n\t"
        "madc.lo.cc.u32 %12, %16, %19, %12; madc.hi.cc.u32 %13, %16, %19, %13;\n\t"
        "madc.lo.cc.u32 %14, %16, %20, %14; madc.hi.cc.u32 %15, %16, %20, %15;\n\t"
        "addc.u32 %7, %7, 0;"
        : "+r"(C[D]), "+r"(C[D + 1]), "+r"(C[D + 2]), "+r"(C[D + 3]), "+r"(C[D + 4]), "+r"(C[D + 5]), "=&r"(C[D + 6]),
          "=&r"(C[D + 7]), "+r"(A[B]), "+r"(A[B + 1]), "+r"(A[B + 2]), "+r"(A[B + 3]), "+r"(A[B + 4]), "+r"(A[B + 5]),
          "+r"(A[B + 6]), "+r"(A[B + 7])
        : "r"(s), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(c0), "r"(c1), "r"(c2), "r"(c3));
}

// ---------------------------------------------------------------------------------------------
// 256 x 256 -> 512-bit product: 64 IMAD.WIDE + 7 carry adds + 15 merge adds.   Requires b < 2^256 - 2^225.
// Row i multiplies a[i] into two accumulators of 64-bit columns: E (columns at even limb positions) and
// O (odd positions; O[k] is limb k + 1).  Even limbs of b land on the accumulator whose columns start at limb i,
// odd limbs of b on the other one, one limb higher.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fp_mul_wide(u32* r /*16*/, const u32* a /*8*/, const u32* b /*8*/) {
    u32 E[16], O[14];
    chain_fresh<0>(E, a[0], b[0], b[2], b[4], b[6]);                             // limbs 0..7
    chain_fresh<0>(O, a[0], b[1], b[3], b[5], b[7]);                             // limbs 1..8
    mul_row<0, 2>(O, E, a[1], b[0], b[2], b[4], b[6], b[1], b[3], b[5], b[7]);   // O: limbs 1..8,  E: limbs 2..9
    mul_row<2, 2>(E, O, a[2], b[0], b[2], b[4], b[6], b[1], b[3], b[5], b[7]);   // E: limbs 2..9,  O: limbs 3..10
    mul_row<2, 4>(O, E, a[3], b[0], b[2], b[4], b[6], b[1], b[3], b[5], b[7]);   // O: 3..10,  E: 4..11
    mul_row<4, 4>(E, O, a[4], b[0], b[2], b[4], b[6], b[1], b[3], b[5], b[7]);   // E: 4..11,  O: 5..12
    mul_row<4, 6>(O, E, a[5], b[0], b[2], b[4], b[6], b[1], b[3], b[5], b[7]);   // O: 5..12,  E: 6..13
    mul_row<6, 6>(E, O, a[6], b[0], b[2], b[4], b[6], b[1], b[3], b[5], b[7]);   // E: 6..13,  O: 7..14
    mul_row<6, 8>(O, E, a[7], b[0], b[2], b[4], b[6], b[1], b[3], b[5], b[7]);   // O: 7..14,  E: 8..15
    // merge: r = E + (O << 32);  O[0..13] are limbs 1..14
    r[0] = E[0];
    asm("add.cc.u32  %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32    %14, %29, 0;"
        : "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]),
          "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]),
          "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]), "r"(O[0]), "r"(O[1]),
          "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]),
          "r"(O[11]), "r"(O[12]), "r"(O[13]));
}

// ---------------------------------------------------------------------------------------------
// Two-term dot product  r = a * b + a2 * b2  (512 bits) in ONE pair of column accumulators: row i multiplies a[i]
// into the columns and then a2[i] into the SAME columns, so there is one merge instead of two merges and a 512-bit
// add.  Requires b, b2 <= 2p: their top limbs are below 2^31, so the top column of the odd chain,
// a[i] b[7] + a2[i] b2[7] + carries, still fits 64 bits and keeps room for the carries of the even chain.
// ---------------------------------------------------------------------------------------------
template <int B, int D>
__device__ __forceinline__ void mul2_row(u32* A, u32* C, u32 s, u32 s2, const u32* b, const u32* b2) {
    asm(// C chain, first term: top two limbs fresh
        "mad.lo.cc.u32  %0, %16, %19, %0;  madc.hi.cc.u32 %1, %16, %19, %1;\n\t"
        "madc.lo.cc.u32 %2, %16, %21, %2;  madc.hi.cc.u32 %3, %16, %21, %3;\n\t"
        "madc.lo.cc.u32 %4, %16, %23, %4;  madc.hi.cc.u32 %5, %16, %23, %5;\n\t"
        "madc.lo.cc.u32 %6, %16, %25, 0;   madc.hi.u32    %7, %16, %25, 0;\n\t"
        // C chain, second term: all eight limbs hold data, no carry out (see above)
        "mad.lo.cc.u32  %0, %17, %27, %0;  madc.hi.cc.u32 %1, %17, %27, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %29, %2;  madc.hi.cc.u32 %3, %17, %29, %3;\n\t"
        "madc.lo.cc.u32 %4, %17, %31, %4;  madc.hi.cc.u32 %5, %17, %31, %5;\n\t"
        "madc.lo.cc.u32 %6, %17, %33, %6;  madc.hi.u32    %7, %17, %33, %7;\n\t"
        // A chain, first term; carry -> C top limb
        "mad.lo.cc.u32  %8,  %16, %18, %8;  madc.hi.cc.u32 %9,  %16, %18, %9;\n\t"
        "madc.lo.cc.u32 %10, %16, %20, %10; madc.hi.cc.u32 %11, %16, %20, %11;\n\t"
        "madc.lo.cc.u32 %12, %16, %22, %12; madc.hi.cc.u32 %13, %16, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %16, %24, %14; madc.hi.cc.u32 %15, %16, %24, %15;\n\t"
        "addc.u32 %7, %7, 0;\n\t"
        // A chain, second term; carry -> C top limb
        "mad.lo.cc.u32  %8,  %17, %26, %8;  madc.hi.cc.u32 %9,  %17, %26, %9;\n\t"
        "madc.lo.cc.u32 %10, %17, %28, %10; madc.hi.cc.u32 %11, %17, %28, %11;\n\t"
        "madc.lo.cc.u32 %12, %17, %30, %12; madc.hi.cc.u32 %13, %17, %30, %13;\n\t"
        "madc.lo.cc.u32 %14, %17, %32, %14; madc.hi.cc.u32 %15, %17, %32, %15;\n\t"
        "addc.u32 %7, %7, 0;"
        : "+r"(C[D]), "+r"(C[D + 1]), "+r"(C[D + 2]), "+r"(C[D + 3]), "+r"(C[D + 4]), "+r"(C[D + 5]), "=&r"(C[D + 6]),
          "=&r"(C[D + 7]), "+r"(A[B]), "+r"(A[B + 1]), "+r"(A[B + 2]), "+r"(A[B + 3]), "+r"(A[B + 4]), "+r"(A[B + 5]),
          "+r"(A[B + 6]), "+r"(A[B + 7])
        : "r"(s), "r"(s2), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]),
          "r"(b2[0]), "r"(b2[1]), "r"(b2[2]), "r"(b2[3]), "r"(b2[4]), "r"(b2[5]), "r"(b2[6]), "r"(b2[7]));
}

// row 0: the first term writes the columns, the second accumulates (E: limbs 0..7, carry -> O[7]; O: limbs 1..8)
__device__ __forceinline__ void mul2_row0(u32* E, u32* O, u32 s, u32 s2, const u32* b, const u32* b2) {
    chain_fresh<0>(E, s, b[0], b[2], b[4], b[6]);
    chain_fresh<0>(O, s, b[1], b[3], b[5], b[7]);
    asm("mad.lo.cc.u32  %0, %16, %18, %0;  madc.hi.cc.u32 %1, %16, %18, %1;\n\t"
        "madc.lo.cc.u32 %2, %16, %20, %2;  madc.hi.cc.u32 %3, %16, %20, %3;\n\t"
        "madc.lo.cc.u32 %4, %16, %22, %4;  madc.hi.cc.u32 %5, %16, %22, %5;\n\t"
        "madc.lo.cc.u32 %6, %16, %24, %6;  madc.hi.u32    %7, %16, %24, %7;\n\t"
        "mad.lo.cc.u32  %8,  %16, %17, %8;  madc.hi.cc.u32 %9,  %16, %17, %9;\n\t"
        "madc.lo.cc.u32 %10, %16, %19, %10; madc.hi.cc.u32 %11, %16, %19, %11;\n\t"
        "madc.lo.cc.u32 %12, %16, %21, %12; madc.hi.cc.u32 %13, %16, %21, %13;\n\t"
        "madc.lo.cc.u32 %14, %16, %23, %14; madc.hi.cc.u32 %15, %16, %23, %15;\n\t"
        "addc.u32 %7, %7, 0;"
        : "+r"(O[0]), "+r"(O[1]), "+r"(O[2]), "+r"(O[3]), "+r"(O[4]), "+r"(O[5]), "+r"(O[6]), "+r"(O[7]), "+r"(E[0]),
          "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7])
        : "r"(s2), "r"(b2[0]), "r"(b2[1]), "r"(b2[2]), "r"(b2[3]), "r"(b2[4]), "r"(b2[5]), "r"(b2[6]), "r"(b2[7]));
}

__device__ __forceinline__ void wide_merge(u32* r /*16*/, const u32* E /*16*/, const u32* O /*14*/) {
    // r = E + (O << 32);  O[0..13] are limbs 1..14
    r[0] = E[0];
    asm("add.cc.u32  %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32    %14, %29, 0;"
        : "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]),
          "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]),
          "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]), "r"(O[0]), "r"(O[1]),
          "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]),
          "r"(O[11]), "r"(O[12]), "r"(O[13]));
}

__device__ __forceinline__ void fp_mul2_wide(u32* r /*16*/, const u32* a, const u32* b, const u32* a2, const u32* b2) {
    u32 E[16], O[14];
    mul2_row0(E, O, a[0], a2[0], b, b2);
    mul2_row<0, 2>(O, E, a[1], a2[1], b, b2);
    mul2_row<2, 2>(E, O, a[2], a2[2], b, b2);
    mul2_row<2, 4>(O, E, a[3], a2[3], b, b2);
    mul2_row<4, 4>(E, O, a[4], a2[4], b, b2);
    mul2_row<4, 6>(O, E, a[5], a2[5], b, b2);
    mul2_row<6, 6>(E, O, a[6], a2[6], b, b2);
    mul2_row<6, 8>(O, E, a[7], a2[7], b, b2);
    wide_merge(r, E, O);
}

// ---------------------------------------------------------------------------------------------
// Montgomery reduction rows.  Row i clears limb i of (E + O<<32) by adding m_i * p * 2^(32 i).
// The limb being cleared lives half in E and half in O; their sum s gives m_i, and the carry of
// that sum enters the chain that starts one limb higher.  Same carry discipline as mul_row: the chain at the
// cleared limb (even limbs of p) hands its carry to the top limb of the chain one limb higher (odd limbs
// of p, top product m * P7 < 2^62).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void redc_row0(u32* E, u32* O) {
    u32 m, junk;
    asm("{ .reg .u64 t0, t1, t2, t3;\n\t"
        "mul.lo.u32 %16, %0, " BNP_STR(BNP_N0INV) ";\n\t"
        // O base 0 (limbs 1..8), odd limbs of p, all fresh
        "mul.wide.u32 t0, %16, " BNP_STR(BNP_P1) "; mul.wide.u32 t1, %16, " BNP_STR(BNP_P3) ";\n\t"
        "mul.wide.u32 t2, %16, " BNP_STR(BNP_P5) "; mul.wide.u32 t3, %16, " BNP_STR(BNP_P7) ";\n\t"
        "mov.b64 {%8, %9}, t0; mov.b64 {%10, %11}, t1; mov.b64 {%12, %13}, t2; mov.b64 {%14, %15}, t3;\n\t"
        // E base 0 (limbs 0..7), even limbs of p, all data; the low word becomes zero and is dropped; carry -> limb 8 = O[7]
        "mad.lo.cc.u32  %17, %16, " BNP_STR(BNP_P0) ", %0; madc.hi.cc.u32 %1, %16, " BNP_STR(BNP_P0) ", %1;\n\t"
        "madc.lo.cc.u32 %2,  %16, " BNP_STR(BNP_P2) ", %2; madc.hi.cc.u32 %3, %16, " BNP_STR(BNP_P2) ", %3;\n\t"
        "madc.lo.cc.u32 %4,  %16, " BNP_STR(BNP_P4) ", %4; madc.hi.cc.u32 %5, %16, " BNP_STR(BNP_P4) ", %5;\n\t"
        "madc.lo.cc.u32 %6,  %16, " BNP_STR(BNP_P6) ", %6; madc.hi.cc.u32 %7, %16, " BNP_STR(BNP_P6) ", %7;\n\t"
        "addc.u32 %15, %15, 0; }"
        : "+r"(E[0]), "+r"(E[1]), "+r"(E[2]), "+r"(E[3]), "+r"(E[4]), "+r"(E[5]), "+r"(E[6]), "+r"(E[7]),
          "=&r"(O[0]), "=&r"(O[1]), "=&r"(O[2]), "=&r"(O[3]), "=&r"(O[4]), "=&r"(O[5]), "=&r"(O[6]), "=&r"(O[7]),
          "=&r"(m), "=&r"(junk));
}

// A: accumulator whose column starts AT the limb being cleared (A[B..B+7] all data).
// C: accumulator whose chain starts one limb higher (C[D..D+5] data, C[D+6], C[D+7] fresh); its other
//    contribution to the cleared limb is the single entry `hi` (already final).
template <int B, int D>
__device__ __forceinline__ void redc_row(u32* A, u32* C, u32 hi) {
    u32 m, s, junk;
    asm("add.cc.u32 %17, %0, %19;\n\t"
        "mul.lo.u32 %16, %17, " BNP_STR(BNP_N0INV) ";\n\t"
        // chain one limb higher: odd limbs of p, carry-in from the add above
        "madc.lo.cc.u32 %8,  %16, " BNP_STR(BNP_P1) ", %8;  madc.hi.cc.u32 %9,  %16, " BNP_STR(BNP_P1) ", %9;\n\t"
        "madc.lo.cc.u32 %10, %16, " BNP_STR(BNP_P3) ", %10; madc.hi.cc.u32 %11, %16, " BNP_STR(BNP_P3) ", %11;\n\t"
        "madc.lo.cc.u32 %12, %16, " BNP_STR(BNP_P5) ", %12; madc.hi.cc.u32 %13, %16, " BNP_STR(BNP_P5) ", %13;\n\t"
        "madc.lo.cc.u32 %14, %16, " BNP_STR(BNP_P7) ", 0;   madc.hi.u32    %15, %16, " BNP_STR(BNP_P7) ", 0;\n\t"
        // chain at the cleared limb: even limbs of p; low word becomes zero and is dropped; carry -> C[D+7]
        "mad.lo.cc.u32  %18, %16, " BNP_STR(BNP_P0) ", %17; madc.hi.cc.u32 %1, %16, " BNP_STR(BNP_P0) ", %1;\n\t"
        "madc.lo.cc.u32 %2,  %16, " BNP_STR(BNP_P2) ", %2;  madc.hi.cc.u32 %3, %16, " BNP_STR(BNP_P2) ", %3;\n\t"
        "madc.lo.cc.u32 %4,  %16, " BNP_STR(BNP_P4) ", %4;  madc.hi.cc.u32 %5, %16, " BNP_STR(BNP_P4) ", %5;\n\t"
        "madc.lo.cc.u32 %6,  %16, " BNP_STR(BNP_P6) ", %6;  madc.hi.cc.u32 %7, %16, " BNP_STR(BNP_P6) ", %7;\n\t"
        "addc.u32 %15, %15, 0;"
        : "+r"(A[B]), "+r"(A[B + 1]), "+r"(A[B + 2]), "+r"(A[B + 3]), "+r"(A[B + 4]), "+r"(A[B + 5]),
          "+r"(A[B + 6]), "+r"(A[B + 7]), "+r"(C[D]), "+r"(C[D + 1]), "+r"(C[D + 2]), "+r"(C[D + 3]),
          "+r"(C[D + 4]), "+r"(C[D + 5]), "=&r"(C[D + 6]), "=&r"(C[D + 7]), "=&r"(m), "=&r"(s), "=&r"(junk)
        : "r"(hi));
}

// r = (r >= p) ? r - p : r      (r < 2p on entry)
__device__ __forceinline__ void fp_cond_sub_p(u32* r) {
    u32 t[8], borrow;
    asm("sub.cc.u32  %0, %9,  " BNP_STR(BNP_P0) ";\n\t"
        "subc.cc.u32 %1, %10, " BNP_STR(BNP_P1) ";\n\t"
        "subc.cc.u32 %2, %11, " BNP_STR(BNP_P2) ";\n\t"
        "subc.cc.u32 %3, %12, " BNP_STR(BNP_P3) ";\n\t"
        "subc.cc.u32 %4, %13, " BNP_STR(BNP_P4) ";\n\t"
        "subc.cc.u32 %5, %14, " BNP_STR(BNP_P5) ";\n\t"
        "subc.cc.u32 %6, %15, " BNP_STR(BNP_P6) ";\n\t"
        "subc.cc.u32 %7, %16, " BNP_STR(BNP_P7) ";\n\t"
        "subc.u32    %8, 0, 0;"
        : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]),
          "=&r"(borrow)
        : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = borrow ? r[i] : t[i];
}

// r = (r >= c) ? r - c : r for a constant c given as eight limbs (c = 2p, 4p: the canonicalisation ladder
// of lazily bounded results)
__device__ __forceinline__ void fp_cond_sub_const(u32* r, u32 c0, u32 c1, u32 c2, u32 c3, u32 c4, u32 c5, u32 c6, u32 c7) {
    u32 t[8], borrow;
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]),
          "=&r"(borrow)
        : "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(c0), "r"(c1), "r"(c2),
          "r"(c3), "r"(c4), "r"(c5), "r"(c6), "r"(c7));
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = borrow ? r[i] : t[i];
}

// r in [0, 2p * (lvl + 1)) -> canonical residue; lvl is warp-uniform (an instruction field)
__device__ __forceinline__ void fp_canon(u32* r, u32 lvl) {
    if (lvl >= 2u)
        fp_cond_sub_const(r, 0x61f3f51cu, 0xf082305bu, 0xa1c72a34u, 0x5e05aa45u, 0x06056176u, 0xe14116dau, 0x84c680a6u,
                          0xc19139cbu);  // 4p
    if (lvl >= 1u)
        fp_cond_sub_const(r, 0xb0f9fa8eu, 0x7841182du, 0xd0e3951au, 0x2f02d522u, 0x0302b0bbu, 0x70a08b6du, 0xc2634053u,
                          0x60c89ce5u);  // 2p
    fp_cond_sub_const(r, (u32)BNP_P0, (u32)BNP_P1, (u32)BNP_P2, (u32)BNP_P3, (u32)BNP_P4, (u32)BNP_P5, (u32)BNP_P6,
                      (u32)BNP_P7);
}

// the upper steps of the ladder only: r in [0, 2p * (lvl + 1)) -> [0, 2p)
__device__ __forceinline__ void fp_canon_upper(u32* r, u32 lvl) {
    if (lvl >= 2u)
        fp_cond_sub_const(r, 0x61f3f51cu, 0xf082305bu, 0xa1c72a34u, 0x5e05aa45u, 0x06056176u, 0xe14116dau, 0x84c680a6u,
                          0xc19139cbu);  // 4p
    if (lvl >= 1u)
        fp_cond_sub_const(r, 0xb0f9fa8eu, 0x7841182du, 0xd0e3951au, 0x2f02d522u, 0x0302b0bbu, 0x70a08b6du, 0xc2634053u,
                          0x60c89ce5u);  // 2p
}

// Montgomery reduction of a 512-bit T < p * 2^256: r = T / 2^256 mod p, canonical.
// 8 IMAD + 64 IMAD.WIDE.
// fp_redc_lazy: any T < 2^512 - p * 2^256; the result is T / 2^256 + (< p), NOT canonicalised.
__device__ __forceinline__ void fp_redc_lazy(u32* r /*8*/, const u32* T /*16*/) {
    u32 E[16], O[14];
#pragma unroll
    for (int i = 0; i < 8; i++) E[i] = T[i];
    redc_row0(E, O);                 // clears limb 0;  E: limbs 0..7, O: limbs 1..8
    redc_row<0, 2>(O, E, E[1]);      // row 1: limb 1 = O[0] + E[1];  O chain limbs 1..8, E chain limbs 2..9
    redc_row<2, 2>(E, O, O[1]);      // row 2: limb 2 = E[2] + O[1];  E chain limbs 2..9, O chain limbs 3..10
    redc_row<2, 4>(O, E, E[3]);      // row 3
    redc_row<4, 4>(E, O, O[3]);      // row 4
    redc_row<4, 6>(O, E, E[5]);      // row 5
    redc_row<6, 6>(E, O, O[5]);      // row 6:  E chain limbs 6..13, O chain limbs 7..14
    redc_row<6, 8>(O, E, E[7]);      // row 7:  O chain limbs 7..14, E chain limbs 8..15
    // result limbs k = 0..7 (limb 8 + k): E[8+k] + O[7+k] + T[8+k];  O[14] (limb 15) does not exist
    u32 u[8];
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, 0;"
        : "=&r"(u[0]), "=&r"(u[1]), "=&r"(u[2]), "=&r"(u[3]), "=&r"(u[4]), "=&r"(u[5]), "=&r"(u[6]), "=&r"(u[7])
        : "r"(E[8]), "r"(E[9]), "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]),
          "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]), "r"(O[12]), "r"(O[13]));
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, %23;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
        : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(T[8]),
          "r"(T[9]), "r"(T[10]), "r"(T[11]), "r"(T[12]), "r"(T[13]), "r"(T[14]), "r"(T[15]));
}

__device__ __forceinline__ void fp_redc(u32* r /*8*/, const u32* T /*16*/) {
    fp_redc_lazy(r, T);
    fp_cond_sub_p(r);
}

// ---------------------------------------------------------------------------------------------
// 256-bit and 512-bit add / sub helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 add8(u32* r, const u32* a, const u32* b) {  // returns carry
    u32 c;
    asm("add.cc.u32  %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
}

__device__ __forceinline__ u32 sub8(u32* r, const u32* a, const u32* b) {  // returns 0 or 0xffffffff (borrow)
    u32 c;
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
}

// r += p & mask   (mask is 0 or 0xffffffff)
__device__ __forceinline__ void add_p_masked(u32* r, u32 mask) {
    asm("add.cc.u32  %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32    %7, %7, %15;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
        : "r"(mask & (u32)BNP_P0), "r"(mask & (u32)BNP_P1), "r"(mask & (u32)BNP_P2), "r"(mask & (u32)BNP_P3),
          "r"(mask & (u32)BNP_P4), "r"(mask & (u32)BNP_P5), "r"(mask & (u32)BNP_P6), "r"(mask & (u32)BNP_P7));
}

// 512-bit r = a - b, returns borrow mask
__device__ __forceinline__ u32 sub16(u32* r, const u32* a, const u32* b) {
    u32 c;
    asm("sub.cc.u32  %0, %17, %33;\n\t"
        "subc.cc.u32 %1, %18, %34;\n\t"
        "subc.cc.u32 %2, %19, %35;\n\t"
        "subc.cc.u32 %3, %20, %36;\n\t"
        "subc.cc.u32 %4, %21, %37;\n\t"
        "subc.cc.u32 %5, %22, %38;\n\t"
        "subc.cc.u32 %6, %23, %39;\n\t"
        "subc.cc.u32 %7, %24, %40;\n\t"
        "subc.cc.u32 %8, %25, %41;\n\t"
        "subc.cc.u32 %9, %26, %42;\n\t"
        "subc.cc.u32 %10, %27, %43;\n\t"
        "subc.cc.u32 %11, %28, %44;\n\t"
        "subc.cc.u32 %12, %29, %45;\n\t"
        "subc.cc.u32 %13, %30, %46;\n\t"
        "subc.cc.u32 %14, %31, %47;\n\t"
        "subc.cc.u32 %15, %32, %48;\n\t"
        "subc.u32    %16, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]),
          "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15]),
          "=&r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]),
          "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]), "r"(b[0]), "r"(b[1]),
          "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]), "r"(b[8]), "r"(b[9]), "r"(b[10]),
          "r"(b[11]), "r"(b[12]), "r"(b[13]), "r"(b[14]), "r"(b[15]));
    return c;
}

// ---------------------------------------------------------------------------------------------
// canonical Fp ops (inputs and outputs in [0, p))
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fp_add(u32* r, const u32* a, const u32* b) {
    add8(r, a, b);  // < 2p < 2^255: no carry
    fp_cond_sub_p(r);
}

__device__ __forceinline__ void fp_sub(u32* r, const u32* a, const u32* b) {
    u32 borrow = sub8(r, a, b);
    add_p_masked(r, borrow);
}

__device__ __forceinline__ void fp_neg(u32* r, const u32* a) {
    u32 nz = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) nz |= a[i];
    u32 pp[8] = {(u32)BNP_P0, (u32)BNP_P1, (u32)BNP_P2, (u32)BNP_P3, (u32)BNP_P4, (u32)BNP_P5, (u32)BNP_P6, (u32)BNP_P7};
    u32 t[8];
    sub8(t, pp, a);
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = nz ? t[i] : 0u;
}

__device__ __forceinline__ void fp_mul(u32* r, const u32* a, const u32* b) {
    u32 T[16];
    fp_mul_wide(T, a, b);
    fp_redc(r, T);
}
