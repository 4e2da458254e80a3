// Wire formats on the device: points arrive as bytes - ark-serialize 0.4 `CanonicalSerialize`
// (what a Groth16 verifier written against arkworks holds) or EIP-196/197 big-endian words (what an Ethereum client
// holds) - and leave as the Montgomery structure-of-arrays the pairing kernels read; Fq12 results leave as ark bytes.
// One thread per element, plain fp.cuh arithmetic (no sequencer): decoding is <3 % of a pairing even with the square
// roots of the compressed forms, and it is branchy per element (flags, infinity, malformed input), which the
// grid-uniform sequencer cannot be.
//
// ark-ec 0.4.2 short-Weierstrass affine serialisation (models/short_weierstrass/mod.rs::serialize_with_mode,
// serialization_flags.rs::SWFlags), restated:
//   uncompressed: x || y, compressed: x; every Fq as 32 little-endian bytes of the CANONICAL (non-Montgomery) integer,
//   an Fq2 as c0 || c1; the two top bits of the LAST byte carry the flags: bit 7 = "y is negative" (y > -y in the
//   field's ordering: for Fq2, c1 is compared first, then c0), bit 6 = point at infinity; both set is invalid.
// EIP-196 / EIP-197: 32-byte big-endian words, G1 = x || y, G2 = x.c1 || x.c0 || y.c1 || y.c0 (imaginary part first),
//   all-zero = point at infinity, no flags.
#pragma once
#include "fp2.cuh"

#define BNP_WIRE_ARK_UNCOMPRESSED 0
#define BNP_WIRE_ARK_COMPRESSED 1
#define BNP_WIRE_EIP197 2

// per-element status of a decode (include/bnp.h: BNP_POINT_*)
#define BNP_PT_OK 0
#define BNP_PT_INFINITY 1
#define BNP_PT_NOT_CANONICAL 2   // a coordinate >= p, or an invalid flag combination
#define BNP_PT_NOT_ON_CURVE 3    // y^2 != x^3 + b, or x^3 + b has no square root (compressed)
#define BNP_PT_NOT_IN_SUBGROUP 4 // G2 only: on the twist but outside the r-torsion (set by the host from validate_g2)

__device__ __constant__ u32 WIRE_R2[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                           0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};  // R^2 mod p
__device__ __constant__ u32 WIRE_THREE[8] = {0x50ad28d7u, 0x7a17caa9u, 0xe15521b9u, 0x1f6ac17au,
                                              0x696bd284u, 0x334bea4eu, 0xce179d8eu, 0x2a1f6744u};  // 3 (Montgomery)
// b' = 3 / (9 + u), the twist's constant (Montgomery)
__device__ __constant__ u32 WIRE_B2C0[8] = {0x77b802a8u, 0x3bf938e3u, 0x3633535du, 0x020b1b27u,
                                             0x49755260u, 0x26b7edf0u, 0x4384a86du, 0x2514c632u};
__device__ __constant__ u32 WIRE_B2C1[8] = {0xd1dcff67u, 0x38e7ecccu, 0x93ce0d3eu, 0x65f0b37du,
                                             0x22ac00aau, 0xd749d0ddu, 0x4a688d4du, 0x0141b9ceu};
__device__ __constant__ u32 WIRE_EXP_SQRT[8] = {0xb61f3f52u, 0x4f082305u, 0x5a1c72a3u, 0x65e05aa4u,
                                                 0xa0605617u, 0x6e14116du, 0xb84c680au, 0x0c19139cu};  // (p + 1) / 4
__device__ __constant__ u32 WIRE_HALF[8] = {0x6c3e7ea3u, 0x9e10460bu, 0xb438e546u, 0xcbc0b548u,
                                             0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};  // (p - 1) / 2
__device__ __constant__ u32 WIRE_INV2[8] = {0x4f060572u, 0x87bee7d2u, 0x2f1c6ae5u, 0xd0fd2addu,
                                             0xfcfd4f44u, 0x8f5f7492u, 0x3d9cbfacu, 0x1f37631au};  // 1/2 (Montgomery)

// ---- small helpers on eight-limb values ----
__device__ __forceinline__ bool w_is_zero(const u32* a) {
    u32 x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= a[i];
    return x == 0u;
}
__device__ __forceinline__ bool w_eq(const u32* a, const u32* b) {
    u32 x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x |= a[i] ^ b[i];
    return x == 0u;
}
// a > b as 256-bit integers
__device__ __forceinline__ bool w_gt(const u32* a, const u32* b) {
    u32 t[8];
    return sub8(t, b, a) != 0u;  // borrow of b - a
}
__device__ __forceinline__ bool w_lt_p(const u32* a) {
    u32 t[8];
    return sub8(t, a, BNP_P) != 0u;  // borrow of a - p
}
__device__ __forceinline__ void w_set(u32* r, const u32* a) {
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = a[i];
}
// Montgomery -> canonical integer: redc(a, 0)
__device__ __forceinline__ void w_from_mont(u32* r, const u32* a) {
    u32 T[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { T[i] = a[i]; T[8 + i] = 0u; }
    fp_redc(r, T);
}
__device__ __forceinline__ void w_to_mont(u32* r, const u32* a) { fp_mul(r, a, WIRE_R2); }

// 32 bytes at `p` (4-byte aligned) -> eight limbs, little-endian limb order; `be`: the bytes are one big-endian word
__device__ __forceinline__ void w_read32(u32* r, const unsigned char* p, bool be) {
    const u32* q = reinterpret_cast<const u32*>(p);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const u32 w = q[be ? 7 - i : i];
        r[i] = be ? __byte_perm(w, 0u, 0x0123) : w;
    }
}
__device__ __forceinline__ void w_write32_le(unsigned char* p, const u32* a) {
    u32* q = reinterpret_cast<u32*>(p);
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = a[i];
}
__device__ __forceinline__ void w_store_soa(u64* arr, u32 f, size_t n, size_t e, const u32* a) {
    u64* p = arr + (size_t)f * 4 * n + e;
#pragma unroll
    for (int j = 0; j < 4; j++) p[(size_t)j * n] = (u64)a[2 * j] | ((u64)a[2 * j + 1] << 32);
}
__device__ __forceinline__ void w_load_soa(u32* a, const u64* arr, u32 f, size_t n, size_t e) {
    const u64* p = arr + (size_t)f * 4 * n + e;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const u64 v = p[(size_t)j * n];
        a[2 * j] = (u32)v;
        a[2 * j + 1] = (u32)(v >> 32);
    }
}

// a^((p+1)/4): the square root of a quadratic residue (p = 3 mod 4); the caller checks r^2 == a
__device__ __noinline__ void w_sqrt_candidate(u32* r, const u32* a) {
    u32 acc[8], base[8];
    w_set(base, a);
    acc[0] = 0xc58f0d9du; acc[1] = 0xd35d438du; acc[2] = 0xf5c70b3du; acc[3] = 0x0a78eb28u;
    acc[4] = 0x7879462cu; acc[5] = 0x666ea36fu; acc[6] = 0x9a07df2fu; acc[7] = 0x0e0a77c1u;  // 1
#pragma unroll 1
    for (int bit = 251; bit >= 0; bit--) {  // (p + 1) / 4 has 252 bits
        u32 t[8];
        fp_mul(t, acc, acc);
        w_set(acc, t);
        if ((WIRE_EXP_SQRT[bit >> 5] >> (bit & 31)) & 1u) {
            fp_mul(t, acc, base);
            w_set(acc, t);
        }
    }
    w_set(r, acc);
}
// r = sqrt(a) in Fq if a is a square; returns whether it is
__device__ __forceinline__ bool w_fp_sqrt(u32* r, const u32* a) {
    u32 c[8], c2[8];
    w_sqrt_candidate(c, a);
    fp_mul(c2, c, c);
    w_set(r, c);
    return w_eq(c2, a);
}
// r = sqrt(a) in Fq2 = Fq[u]/(u^2 + 1) (complex method: the norm's root s, then x0^2 = (a0 +- s) / 2, x1 = a1 / (2 x0));
// returns whether a is a square.  Either root may come out; the caller picks the sign.
__device__ __forceinline__ bool w_fp2_sqrt(Fp2& r, const Fp2& a) {
    u32 zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (w_is_zero(a.c1)) {
        // a in Fq: sqrt(a0) if a0 is a residue, else u * sqrt(-a0) (exactly one of a0, -a0 is, p = 3 mod 4)
        u32 s[8], na[8];
        if (w_fp_sqrt(s, a.c0)) {
            w_set(r.c0, s);
            w_set(r.c1, zero);
            return true;
        }
        fp_neg(na, a.c0);
        const bool ok = w_fp_sqrt(s, na);
        w_set(r.c0, zero);
        w_set(r.c1, s);
        return ok;
    }
    u32 n[8], t[8], s[8], d[8], x0[8], x1[8];
    fp_mul(n, a.c0, a.c0);
    fp_mul(t, a.c1, a.c1);
    fp_add(n, n, t);            // norm
    if (!w_fp_sqrt(s, n)) return false;
    fp_add(d, a.c0, s);
    fp_mul(d, d, WIRE_INV2);    // (a0 + s) / 2
    if (!w_fp_sqrt(x0, d)) {
        fp_sub(d, a.c0, s);
        fp_mul(d, d, WIRE_INV2);  // (a0 - s) / 2
        if (!w_fp_sqrt(x0, d)) return false;
    }
    fp_add(t, x0, x0);
    fp_inv(t, t);
    fp_mul(x1, a.c1, t);
    w_set(r.c0, x0);
    w_set(r.c1, x1);
    Fp2 chk;
    fp2_sqr(chk, r);
    return w_eq(chk.c0, a.c0) && w_eq(chk.c1, a.c1);
}

// "y > -y" on canonical integers (ark-ff's Ord for Fp: the integer order)
__device__ __forceinline__ bool w_fp_is_larger(const u32* y_mont) {
    u32 y[8];
    w_from_mont(y, y_mont);
    return w_gt(y, WIRE_HALF);
}

// ---------------------------------------------------------------------------------------------
// G1: y^2 = x^3 + 3.  in: n elements of 64 (uncompressed, EIP) or 32 (compressed) bytes; out: [2][4][n] u64 SoA
// ---------------------------------------------------------------------------------------------
__global__ void bnp_decode_g1_kernel(int fmt, const unsigned char* in, size_t stride, size_t n, u64* out,
                                     unsigned char* status) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const bool be = fmt == BNP_WIRE_EIP197, comp = fmt == BNP_WIRE_ARK_COMPRESSED;
    const unsigned char* p = in + e * stride;  // stride: 64 / 32 bytes for a packed array, 192 inside EIP-197 input
    u32 x[8], y[8], zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    u32 st = BNP_PT_OK, flags = 0u;
    w_read32(x, p, be);
    if (comp) {
        flags = x[7] >> 30;
        x[7] &= 0x3fffffffu;
        w_set(y, zero);
    } else {
        w_read32(y, p + 32, be);
        if (!be) {
            flags = y[7] >> 30;
            y[7] &= 0x3fffffffu;
        }
    }
    const bool inf = be ? (w_is_zero(x) && w_is_zero(y)) : (flags & 1u) != 0u;
    if (!be && flags == 3u) st = BNP_PT_NOT_CANONICAL;
    else if (inf) st = BNP_PT_INFINITY;
    else if (!w_lt_p(x) || !w_lt_p(y)) st = BNP_PT_NOT_CANONICAL;
    if (st == BNP_PT_OK) {
        u32 xm[8], ym[8], t[8], y2[8];
        w_to_mont(xm, x);
        fp_mul(t, xm, xm);
        fp_mul(t, t, xm);
        fp_add(t, t, WIRE_THREE);  // x^3 + 3
        if (comp) {
            if (!w_fp_sqrt(ym, t)) st = BNP_PT_NOT_ON_CURVE;
            else if (w_fp_is_larger(ym) != ((flags & 2u) != 0u)) fp_neg(ym, ym);
        } else {
            w_to_mont(ym, y);
            fp_mul(y2, ym, ym);
            if (!w_eq(y2, t)) st = BNP_PT_NOT_ON_CURVE;
        }
        if (st == BNP_PT_OK) {
            w_store_soa(out, 0, n, e, xm);
            w_store_soa(out, 1, n, e, ym);
        }
    }
    if (st != BNP_PT_OK) {
        w_store_soa(out, 0, n, e, zero);
        w_store_soa(out, 1, n, e, zero);
    }
    status[e] = (unsigned char)st;
}

// ---------------------------------------------------------------------------------------------
// G2: y^2 = x^3 + 3/(9+u) over Fq2.  in: 128 (uncompressed, EIP) or 64 (compressed) bytes; out: [4][4][n] u64 SoA
// (x.c0, x.c1, y.c0, y.c1).  The r-torsion test is the sequencer program validate_g2, run by the host afterwards.
// ---------------------------------------------------------------------------------------------
__global__ void bnp_decode_g2_kernel(int fmt, const unsigned char* in, size_t stride, size_t n, u64* out,
                                     unsigned char* status) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const bool be = fmt == BNP_WIRE_EIP197, comp = fmt == BNP_WIRE_ARK_COMPRESSED;
    const unsigned char* p = in + e * stride;
    u32 x0[8], x1[8], y0[8], y1[8], zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    u32 st = BNP_PT_OK, flags = 0u;
    // EIP-197 puts the imaginary part first
    w_read32(be ? x1 : x0, p, be);
    w_read32(be ? x0 : x1, p + 32, be);
    if (comp) {
        flags = x1[7] >> 30;
        x1[7] &= 0x3fffffffu;
        w_set(y0, zero);
        w_set(y1, zero);
    } else {
        w_read32(be ? y1 : y0, p + 64, be);
        w_read32(be ? y0 : y1, p + 96, be);
        if (!be) {
            flags = y1[7] >> 30;
            y1[7] &= 0x3fffffffu;
        }
    }
    const bool inf = be ? (w_is_zero(x0) && w_is_zero(x1) && w_is_zero(y0) && w_is_zero(y1)) : (flags & 1u) != 0u;
    if (!be && flags == 3u) st = BNP_PT_NOT_CANONICAL;
    else if (inf) st = BNP_PT_INFINITY;
    else if (!w_lt_p(x0) || !w_lt_p(x1) || !w_lt_p(y0) || !w_lt_p(y1)) st = BNP_PT_NOT_CANONICAL;
    if (st == BNP_PT_OK) {
        Fp2 X, Y, T, Y2;
        w_to_mont(X.c0, x0);
        w_to_mont(X.c1, x1);
        fp2_sqr(T, X);
        fp2_mul(T, T, X);
        fp_add(T.c0, T.c0, WIRE_B2C0);
        fp_add(T.c1, T.c1, WIRE_B2C1);  // x^3 + b'
        if (comp) {
            if (!w_fp2_sqrt(Y, T)) st = BNP_PT_NOT_ON_CURVE;
            else {
                // ark's order on Fq2: c1 first, then c0
                const bool larger = w_is_zero(Y.c1) ? w_fp_is_larger(Y.c0) : w_fp_is_larger(Y.c1);
                if (larger != ((flags & 2u) != 0u)) fp2_neg(Y, Y);
            }
        } else {
            w_to_mont(Y.c0, y0);
            w_to_mont(Y.c1, y1);
            fp2_sqr(Y2, Y);
            if (!w_eq(Y2.c0, T.c0) || !w_eq(Y2.c1, T.c1)) st = BNP_PT_NOT_ON_CURVE;
        }
        if (st == BNP_PT_OK) {
            w_store_soa(out, 0, n, e, X.c0);
            w_store_soa(out, 1, n, e, X.c1);
            w_store_soa(out, 2, n, e, Y.c0);
            w_store_soa(out, 3, n, e, Y.c1);
        }
    }
    if (st != BNP_PT_OK)
        for (u32 f = 0; f < 4; f++) w_store_soa(out, f, n, e, zero);
    status[e] = (unsigned char)st;
}

// status[e] = NOT_IN_SUBGROUP (and zeroed coordinates, like every other failure) where the point decoded fine but
// validate_g2 said no
__global__ void bnp_merge_subgroup_kernel(unsigned char* status, const unsigned char* ok, size_t n, u64* g2) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (status[e] == BNP_PT_OK && !ok[e]) {
        status[e] = BNP_PT_NOT_IN_SUBGROUP;
        const u32 zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        for (u32 f = 0; f < 4; f++) w_store_soa(g2, f, n, e, zero);
    }
}

// ---------------------------------------------------------------------------------------------
// Fq12 <-> ark bytes.  MyFq12 coefficient i is (coeffs[i] + coeffs[i + 6] u) w^i; ark's Fq12 = c0 + c1 w over
// Fq6 = Fq2[v], v = w^2, serialises c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2 = w^0, w^2, w^4, w^1, w^3, w^5
// (SURVEY A.1), each Fq2 as c0 || c1, each Fq as 32 little-endian bytes of the canonical integer: 384 bytes.
// ---------------------------------------------------------------------------------------------
__device__ __constant__ unsigned char WIRE_ARK_ORDER[6] = {0, 2, 4, 1, 3, 5};

__global__ void bnp_encode_fq12_kernel(const u64* in, size_t n, unsigned char* out) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    unsigned char* p = out + e * 384;
    for (u32 k = 0; k < 6; k++) {
        const u32 i = WIRE_ARK_ORDER[k];
        u32 a[8], c[8];
        w_load_soa(a, in, i, n, e);
        w_from_mont(c, a);
        w_write32_le(p + 64 * k, c);
        w_load_soa(a, in, i + 6, n, e);
        w_from_mont(c, a);
        w_write32_le(p + 64 * k + 32, c);
    }
}

// returns per element 0 = fine, BNP_PT_NOT_CANONICAL if some coefficient is >= p (the output is zeroed then)
__global__ void bnp_decode_fq12_kernel(const unsigned char* in, size_t n, u64* out, unsigned char* status) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned char* p = in + e * 384;
    bool ok = true;
    for (u32 k = 0; k < 12; k++) {
        u32 a[8];
        w_read32(a, p + 32 * k, false);
        ok = ok && w_lt_p(a);
    }
    u32 zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    for (u32 k = 0; k < 6; k++) {
        const u32 i = WIRE_ARK_ORDER[k];
        u32 a[8], m[8];
        w_read32(a, p + 64 * k, false);
        w_to_mont(m, a);
        w_store_soa(out, i, n, e, ok ? m : zero);
        w_read32(a, p + 64 * k + 32, false);
        w_to_mont(m, a);
        w_store_soa(out, i + 6, n, e, ok ? m : zero);
    }
    status[e] = ok ? BNP_PT_OK : BNP_PT_NOT_CANONICAL;
}
