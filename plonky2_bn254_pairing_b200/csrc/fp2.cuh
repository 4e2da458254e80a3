// Fq2 = Fq[u]/(u^2+1) on top of fp.cuh; all values canonical Montgomery residues.
// Karatsuba with lazy reduction: an Fq2 product is 3 wide products + 2 Montgomery reductions.
#pragma once
#include "fp.cuh"

struct Fp2 {
    u32 c0[8];
    u32 c1[8];
};

// Operands may be LAZY sums (each component < 2p, as produced by fp2_add_lazy / fp2_sub_lazy) when
// `lazy` is set: then a0+a1 < 4p < 2^256, every product < 16p^2 < 2^512, a0b0 - a1b1 (+ p*2^256) stays in
// [0, p*2^256), and a0b1 + a1b0 < 8p^2, whose reduction is < 2.51p and needs a second conditional subtraction.
__device__ __forceinline__ void fp2_mul(Fp2& r, const Fp2& a, const Fp2& b, bool lazy = false) {
    u32 sa[8], sb[8];
    add8(sa, a.c0, a.c1);  // < 2p < 2^255 (lazy: < 4p < 2^256)
    add8(sb, b.c0, b.c1);
    u32 P0[16], P1[16], P2[16];
    fp_mul_wide(P0, a.c0, b.c0);
    fp_mul_wide(P1, a.c1, b.c1);
    fp_mul_wide(P2, sa, sb);  // < 4p^2 < p*2^256
    sub16(P2, P2, P0);
    sub16(P2, P2, P1);  // a0b1 + a1b0 in [0, 2p^2)
    u32 borrow = sub16(P0, P0, P1);
    add_p_masked(P0 + 8, borrow);  // a0b0 - a1b1 (+ p*2^256 when negative) in [0, p*2^256)
    fp_redc(r.c0, P0);
    fp_redc(r.c1, P2);
    if (lazy) fp_cond_sub_p(r.c1);
}

// Wide (unreduced) forms used by the sequencer's product-class epilogue: T0, T1 are the 512-bit values whose
// Montgomery reductions are the two components of the result.
//   mul:   T0 = a0 b0 - a1 b1 (+ p 2^256 when negative) in [0, p 2^256),  T1 = a0 b1 + a1 b0
//   sqr:   T0 = (a0 + a1)(a0 - a1),  T1 = 2 a0 a1         (a canonical)
//   mulfp: T0 = a0 s, T1 = a1 s
__device__ __forceinline__ void fp2_mul_wide(u32* T0, u32* T1, const Fp2& a, const Fp2& b) {
    u32 sa[8], sb[8];
    add8(sa, a.c0, a.c1);
    add8(sb, b.c0, b.c1);
    u32 P1[16];
    fp_mul_wide(T0, a.c0, b.c0);
    fp_mul_wide(P1, a.c1, b.c1);
    fp_mul_wide(T1, sa, sb);
    sub16(T1, T1, T0);
    sub16(T1, T1, P1);
    u32 borrow = sub16(T0, T0, P1);
    add_p_masked(T0 + 8, borrow);
}

__device__ __forceinline__ void fp2_sqr_wide(u32* T0, u32* T1, const Fp2& a) {
    u32 s[8], d[8], t[8];
    add8(s, a.c0, a.c1);
    fp_sub(d, a.c0, a.c1);
    add8(t, a.c0, a.c0);
    fp_mul_wide(T0, s, d);
    fp_mul_wide(T1, t, a.c1);
}

__device__ __forceinline__ void fp2_mul_fp_wide(u32* T0, u32* T1, const Fp2& a, const u32* s) {
    fp_mul_wide(T0, a.c0, s);
    fp_mul_wide(T1, a.c1, s);
}

// T[8..15] += h (no carry out: the sequencer's bound analysis guarantees T + h 2^256 < 2^512)
__device__ __forceinline__ void wide_add_hi(u32* T, const u32* h) {
    asm("add.cc.u32  %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32    %7, %7, %15;"
        : "+r"(T[8]), "+r"(T[9]), "+r"(T[10]), "+r"(T[11]), "+r"(T[12]), "+r"(T[13]), "+r"(T[14]), "+r"(T[15])
        : "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]));
}

// r = p - a for canonical a (a = 0 gives p, which is fine for a lazy term)
__device__ __forceinline__ void fp_p_minus(u32* r, const u32* a) {
    const u32 pp[8] = {(u32)BNP_P0, (u32)BNP_P1, (u32)BNP_P2, (u32)BNP_P3,
                       (u32)BNP_P4, (u32)BNP_P5, (u32)BNP_P6, (u32)BNP_P7};
    sub8(r, pp, a);
}

// lazy pre-additions for MUL / SQR operands: results in [0, 2p), no reduction
__device__ __forceinline__ void fp2_add_lazy(Fp2& r, const Fp2& a, const Fp2& b) {
    add8(r.c0, a.c0, b.c0);
    add8(r.c1, a.c1, b.c1);
}

__device__ __forceinline__ void fp_sub_lazy(u32* r, const u32* a, const u32* b) {
    // (a - b) mod 2^256 + p: equals a - b + p in (0, 2p) whether or not the subtraction wrapped
    u32 t[8];
    sub8(t, a, b);
    asm("add.cc.u32  %0, %8,  " BNP_STR(BNP_P0) ";\n\t"
        "addc.cc.u32 %1, %9,  " BNP_STR(BNP_P1) ";\n\t"
        "addc.cc.u32 %2, %10, " BNP_STR(BNP_P2) ";\n\t"
        "addc.cc.u32 %3, %11, " BNP_STR(BNP_P3) ";\n\t"
        "addc.cc.u32 %4, %12, " BNP_STR(BNP_P4) ";\n\t"
        "addc.cc.u32 %5, %13, " BNP_STR(BNP_P5) ";\n\t"
        "addc.cc.u32 %6, %14, " BNP_STR(BNP_P6) ";\n\t"
        "addc.u32    %7, %15, " BNP_STR(BNP_P7) ";"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]));
}

__device__ __forceinline__ void fp2_sub_lazy(Fp2& r, const Fp2& a, const Fp2& b) {
    fp_sub_lazy(r.c0, a.c0, b.c0);
    fp_sub_lazy(r.c1, a.c1, b.c1);
}

__device__ __forceinline__ void fp2_sqr(Fp2& r, const Fp2& a) {
    u32 s[8], d[8], t[8];
    add8(s, a.c0, a.c1);     // a0 + a1 < 2p
    fp_sub(d, a.c0, a.c1);   // a0 - a1 mod p
    add8(t, a.c0, a.c0);     // 2 a0 < 2p
    u32 W0[16], W1[16];
    fp_mul_wide(W0, s, d);      // < 2p^2
    fp_mul_wide(W1, t, a.c1);   // < 2p^2
    fp_redc(r.c0, W0);
    fp_redc(r.c1, W1);
}

// r = a * s, s in Fq
__device__ __forceinline__ void fp2_mul_fp(Fp2& r, const Fp2& a, const u32* s) {
    u32 W0[16], W1[16];
    fp_mul_wide(W0, a.c0, s);
    fp_mul_wide(W1, a.c1, s);
    fp_redc(r.c0, W0);
    fp_redc(r.c1, W1);
}

__device__ __forceinline__ void fp2_add(Fp2& r, const Fp2& a, const Fp2& b) {
    fp_add(r.c0, a.c0, b.c0);
    fp_add(r.c1, a.c1, b.c1);
}

__device__ __forceinline__ void fp2_sub(Fp2& r, const Fp2& a, const Fp2& b) {
    fp_sub(r.c0, a.c0, b.c0);
    fp_sub(r.c1, a.c1, b.c1);
}

__device__ __forceinline__ void fp2_neg(Fp2& r, const Fp2& a) {
    fp_neg(r.c0, a.c0);
    fp_neg(r.c1, a.c1);
}

__device__ __forceinline__ void fp2_conj(Fp2& r, const Fp2& a) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.c0[i] = a.c0[i];
    fp_neg(r.c1, a.c1);
}

// v (9 limbs, v < 2^258) -> r = v mod p, canonical.  One quotient estimate from the top 32 bits
// (q_est in {q-1, q}), one multiply-subtract, one conditional subtraction.
__device__ __forceinline__ void fp_reduce_small(u32* r, const u32* v /*9*/) {
    u32 h = (v[8] << 30) | (v[7] >> 2);      // floor(v / 2^226)
    u32 q = h / 202970013u;                   // ceil(p / 2^226); q <= floor(v/p) <= q+1
    // qp = q * p (q <= 20): 9 limbs
    u32 E[8], O[8];
    chain_fresh<0>(E, q, (u32)BNP_P0, (u32)BNP_P2, (u32)BNP_P4, (u32)BNP_P6);
    chain_fresh<0>(O, q, (u32)BNP_P1, (u32)BNP_P3, (u32)BNP_P5, (u32)BNP_P7);
    u32 qp[8];
    qp[0] = E[0];
    asm("add.cc.u32  %0, %7,  %14;\n\t"
        "addc.cc.u32 %1, %8,  %15;\n\t"
        "addc.cc.u32 %2, %9,  %16;\n\t"
        "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t"
        "addc.cc.u32 %5, %12, %19;\n\t"
        "addc.u32    %6, %13, %20;"
        : "=&r"(qp[1]), "=&r"(qp[2]), "=&r"(qp[3]), "=&r"(qp[4]), "=&r"(qp[5]), "=&r"(qp[6]), "=&r"(qp[7])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(O[0]), "r"(O[1]),
          "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]));
    // v - q*p < 2p < 2^255, so limb 8 of the difference is zero and limbs 0..7 suffice
    sub8(r, v, qp);
    fp_cond_sub_p(r);
}

// v = 9*a + c   (9 limbs), a, c < 2^256
__device__ __forceinline__ void mul9_add(u32* v /*9*/, const u32* a, const u32* c) {
    u32 E[9], O[8];
#pragma unroll
    for (int i = 0; i < 8; i++) E[i] = c[i];
    // E columns at limbs 0,2,4,6 accumulate onto c; carry out -> E[8]
    chain_full<0>(E, 9u, a[0], a[2], a[4], a[6]);
    chain_fresh<0>(O, 9u, a[1], a[3], a[5], a[7]);
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

// r = a * (9 + u) = (9 a0 - a1) + (a0 + 9 a1) u
__device__ __forceinline__ void fp2_mul_xi(Fp2& r, const Fp2& a) {
    const u32 pp[8] = {(u32)BNP_P0, (u32)BNP_P1, (u32)BNP_P2, (u32)BNP_P3,
                       (u32)BNP_P4, (u32)BNP_P5, (u32)BNP_P6, (u32)BNP_P7};
    u32 n1[8], v[9], t0[8], t1[8];
    sub8(n1, pp, a.c1);        // p - a1 in (0, p]
    mul9_add(v, a.c0, n1);     // 9 a0 + p - a1 < 10p
    fp_reduce_small(t0, v);
    mul9_add(v, a.c1, a.c0);   // 9 a1 + a0 < 10p
    fp_reduce_small(t1, v);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.c0[i] = t0[i];
        r.c1[i] = t1[i];
    }
}

// Fq inversion by Fermat: a^(p-2).
// Runs once or twice per pairing (easy part / Miller scale), so it is written for size, not speed.
__device__ __noinline__ void fp_inv(u32* r, const u32* a) {
    // p - 2, 32-bit limbs little endian
    const u32 e[8] = {(u32)BNP_P0 - 2u, (u32)BNP_P1, (u32)BNP_P2, (u32)BNP_P3,
                      (u32)BNP_P4, (u32)BNP_P5, (u32)BNP_P6, (u32)BNP_P7};
    u32 acc[8], base[8];
#pragma unroll
    for (int i = 0; i < 8; i++) base[i] = a[i];
    // acc = 1 in Montgomery form = R mod p
    acc[0] = 0xc58f0d9du; acc[1] = 0xd35d438du; acc[2] = 0xf5c70b3du; acc[3] = 0x0a78eb28u;
    acc[4] = 0x7879462cu; acc[5] = 0x666ea36fu; acc[6] = 0x9a07df2fu; acc[7] = 0x0e0a77c1u;
    // left-to-right binary ladder, 254 squarings + popcount(p-2) multiplications.  The exponent is a constant, so the
    // schedule is the same for the whole grid; squaring and multiplication share ONE fp_mul body in the instruction
    // stream (the second factor is selected), because this loop runs 381 times per pairing and the kernel's warm code
    // has to fit the 32 KB instruction cache.
    int bit = 253;
    bool mul_step = false;
#pragma unroll 1
    while (bit >= 0) {
        u32 f[8], t[8];
#pragma unroll
        for (int i = 0; i < 8; i++) f[i] = mul_step ? base[i] : acc[i];
        fp_mul(t, acc, f);
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = t[i];
        if (mul_step) {
            mul_step = false;
            bit--;
        } else {
            u32 w = e[0];
#pragma unroll
            for (int k = 1; k < 8; k++) w = ((bit >> 5) == k) ? e[k] : w;  // no dynamic register indexing
            if ((w >> (bit & 31)) & 1u) mul_step = true; else bit--;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = acc[i];
}

// r = 1 / a in Fq2: conj(a) / (a0^2 + a1^2).  Zero maps to zero (the reference panics instead).
__device__ __forceinline__ void fp2_inv(Fp2& r, const Fp2& a) {
    u32 W0[16], W1[16], n[8], ni[8];
    fp_mul_wide(W0, a.c0, a.c0);
    fp_mul_wide(W1, a.c1, a.c1);
    u32 t0[8], t1[8];
    fp_redc(t0, W0);
    fp_redc(t1, W1);
    fp_add(n, t0, t1);
    fp_inv(ni, n);
    u32 m1[8];
    fp_mul(r.c0, a.c0, ni);
    fp_mul(m1, a.c1, ni);
    fp_neg(r.c1, m1);
}

// ---------------------------------------------------------------------------------------------
// LIN support: reduction of a lazily accumulated nine-limb value
// ---------------------------------------------------------------------------------------------

// t (9 limbs) = m * x, m < 2^10, x 8 limbs
__device__ __forceinline__ void mul_small9(u32* t, u32 m, const u32* x) {
    u32 E[8], O[8];
    chain_fresh<0>(E, m, x[0], x[2], x[4], x[6]);
    chain_fresh<0>(O, m, x[1], x[3], x[5], x[7]);
    t[0] = E[0];
    asm("add.cc.u32  %0, %8,  %15;\n\t"
        "addc.cc.u32 %1, %9,  %16;\n\t"
        "addc.cc.u32 %2, %10, %17;\n\t"
        "addc.cc.u32 %3, %11, %18;\n\t"
        "addc.cc.u32 %4, %12, %19;\n\t"
        "addc.cc.u32 %5, %13, %20;\n\t"
        "addc.cc.u32 %6, %14, %21;\n\t"
        "addc.u32    %7, %22, 0;"
        : "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(t[8])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(O[0]), "r"(O[1]),
          "r"(O[2]), "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]));
}

// v (9 limbs, 0 <= v < 1024 p) -> canonical residue.  Quotient estimate from the top 32 bits:
// h = floor(v / 2^232), D = ceil(p / 2^232) = 3171407, q = floor(h / D) is floor(v/p) or one less.
__device__ __forceinline__ void fp_reduce_lazy(u32* r, const u32* v) {
    const u32 h = (v[8] << 24) | (v[7] >> 8);
    const u32 q = h / 3171407u;
    u32 qp[9];
    const u32 pp[8] = {(u32)BNP_P0, (u32)BNP_P1, (u32)BNP_P2, (u32)BNP_P3,
                       (u32)BNP_P4, (u32)BNP_P5, (u32)BNP_P6, (u32)BNP_P7};
    mul_small9(qp, q, pp);
    sub8(r, v, qp);  // v - q p < 2p < 2^255: the low 8 limbs are the whole value
    fp_cond_sub_p(r);
}
