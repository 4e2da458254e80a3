// Scalar multiplication on G1 and G2: out_i = k_i * P_i, the kernel a verifier needs to fold many
// pairing checks into one by a random linear combination (sum_i r_i * A_i on G1, sum_i r_i * C_i ... - the batched
// Groth16 check), so that the points a pairing batch consumes never have to be produced on the host.
//
// One thread per point, plain fp.cuh / fp2.cuh arithmetic (no sequencer: the doubling-and-addition schedule follows
// the bits of a per-element scalar, which the grid-uniform sequencer cannot do).  Jacobian coordinates on
// y^2 = x^3 + b (a = 0), Z = 0 for the point at infinity; signed 4-bit windows, left to right: a table of P .. 8 P per
// thread (local memory), four doublings and one addition of +-T[|d|] per digit - a warp executes the addition of a
// bit-serial method for nearly every bit anyway (some lane always has the bit set), so the windows cut the additions
// from 256 to 65 (measured: 1.5 x).  The reference gets the same values from ark-ec
// (`G1.mul(s).into()`, `G2.mul(t).into()`: final_exp_native.rs:245-250 builds its test points that way); the group law is
// exact, so any correct schedule gives the same affine result.  Exceptional cases of the addition (accumulator at
// infinity, equal to +-P) are handled per thread: they cannot occur for scalars below r on points of order r, but the
// entry accepts every 256-bit scalar and every on-curve point.
#pragma once
#include "fp2.cuh"

// ---- one interface over Fq (G1) and Fq2 (G2) ----
struct Fp1 {
    u32 v[8];
};
__device__ __forceinline__ void f_mul(Fp1& r, const Fp1& a, const Fp1& b) { fp_mul(r.v, a.v, b.v); }
__device__ __forceinline__ void f_sqr(Fp1& r, const Fp1& a) { fp_mul(r.v, a.v, a.v); }
__device__ __forceinline__ void f_add(Fp1& r, const Fp1& a, const Fp1& b) { fp_add(r.v, a.v, b.v); }
__device__ __forceinline__ void f_sub(Fp1& r, const Fp1& a, const Fp1& b) { fp_sub(r.v, a.v, b.v); }
__device__ __forceinline__ void f_inv(Fp1& r, const Fp1& a) { fp_inv(r.v, a.v); }
__device__ __forceinline__ bool f_is_zero(const Fp1& a) {
    u32 t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= a.v[i];
    return t == 0;
}
__device__ __forceinline__ void f_zero(Fp1& r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
}
__device__ __forceinline__ void f_one(Fp1& r) {  // R mod p (Montgomery 1)
    const u32 one[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = one[i];
}
__device__ __forceinline__ void f_load(Fp1& r, const u64* arr, u32 f, size_t n, size_t e) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const u64 w = arr[((size_t)f * 4 + j) * n + e];
        r.v[2 * j] = (u32)w;
        r.v[2 * j + 1] = (u32)(w >> 32);
    }
}
__device__ __forceinline__ void f_store(u64* arr, u32 f, size_t n, size_t e, const Fp1& a) {
#pragma unroll
    for (int j = 0; j < 4; j++) arr[((size_t)f * 4 + j) * n + e] = (u64)a.v[2 * j] | ((u64)a.v[2 * j + 1] << 32);
}

__device__ __forceinline__ void f_mul(Fp2& r, const Fp2& a, const Fp2& b) { fp2_mul(r, a, b); }
__device__ __forceinline__ void f_sqr(Fp2& r, const Fp2& a) { fp2_sqr(r, a); }
__device__ __forceinline__ void f_add(Fp2& r, const Fp2& a, const Fp2& b) { fp2_add(r, a, b); }
__device__ __forceinline__ void f_sub(Fp2& r, const Fp2& a, const Fp2& b) { fp2_sub(r, a, b); }
__device__ __forceinline__ void f_inv(Fp2& r, const Fp2& a) { fp2_inv(r, a); }
__device__ __forceinline__ bool f_is_zero(const Fp2& a) {
    u32 t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= a.c0[i] | a.c1[i];
    return t == 0;
}
__device__ __forceinline__ void f_zero(Fp2& r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.c0[i] = r.c1[i] = 0;
}
__device__ __forceinline__ void f_one(Fp2& r) {
    Fp1 o;
    f_one(o);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.c0[i] = o.v[i];
        r.c1[i] = 0;
    }
}
// an Fq2 is two consecutive Fq of the SoA array (c0 at field 2 f, c1 at 2 f + 1)
__device__ __forceinline__ void f_load(Fp2& r, const u64* arr, u32 f, size_t n, size_t e) {
    Fp1 a, b;
    f_load(a, arr, 2 * f, n, e);
    f_load(b, arr, 2 * f + 1, n, e);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.c0[i] = a.v[i];
        r.c1[i] = b.v[i];
    }
}
__device__ __forceinline__ void f_store(u64* arr, u32 f, size_t n, size_t e, const Fp2& a) {
    Fp1 x, y;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x.v[i] = a.c0[i];
        y.v[i] = a.c1[i];
    }
    f_store(arr, 2 * f, n, e, x);
    f_store(arr, 2 * f + 1, n, e, y);
}

template <class F>
struct Jac {
    F X, Y, Z;
};

// (X, Y, Z) <- 2 (X, Y, Z), a = 0: 2 products + 5 squares.  Z = 0 stays Z = 0; a point with Y = 0 (order 2, only on
// the twist outside the subgroup) doubles to Z = 0.
template <class F>
__device__ __noinline__ void jac_double(Jac<F>& p) {
    F A, B, C, D, E, G, t;
    f_sqr(A, p.X);
    f_sqr(B, p.Y);
    f_sqr(C, B);
    f_add(t, p.X, B);
    f_sqr(t, t);
    f_sub(t, t, A);
    f_sub(t, t, C);
    f_add(D, t, t);   // D = 2 ((X + B)^2 - A - C) = 4 X Y^2
    f_add(E, A, A);
    f_add(E, E, A);   // E = 3 X^2
    f_sqr(G, E);
    f_mul(t, p.Y, p.Z);
    f_add(p.Z, t, t); // Z3 = 2 Y Z
    f_sub(G, G, D);
    f_sub(p.X, G, D); // X3 = E^2 - 2 D
    f_sub(t, D, p.X);
    f_mul(t, E, t);
    f_add(C, C, C);
    f_add(C, C, C);
    f_add(C, C, C);   // 8 Y^4
    f_sub(p.Y, t, C); // Y3 = E (D - X3) - 8 C
}

// (X, Y, Z) <- (X, Y, Z) + (x, y), the second point affine and not at infinity: 7 products + 4 squares
template <class F>
__device__ __noinline__ void jac_add_affine(Jac<F>& p, const F& x, const F& y) {
    if (f_is_zero(p.Z)) {
        p.X = x;
        p.Y = y;
        f_one(p.Z);
        return;
    }
    F Z2, U2, S2, H, Rr, t, H2, H3, V;
    f_sqr(Z2, p.Z);
    f_mul(U2, x, Z2);
    f_mul(t, p.Z, Z2);
    f_mul(S2, y, t);
    f_sub(H, U2, p.X);
    f_sub(Rr, S2, p.Y);
    if (f_is_zero(H)) {
        if (f_is_zero(Rr)) {
            jac_double(p);  // the same point
        } else {
            f_zero(p.X);    // opposite points
            f_zero(p.Y);
            f_zero(p.Z);
        }
        return;
    }
    f_sqr(H2, H);
    f_mul(H3, H2, H);
    f_mul(V, p.X, H2);
    f_sqr(t, Rr);
    f_sub(t, t, H3);
    f_sub(t, t, V);
    f_sub(p.X, t, V);       // X3 = R^2 - H^3 - 2 V
    f_sub(t, V, p.X);
    f_mul(t, Rr, t);
    f_mul(H3, p.Y, H3);
    f_sub(p.Y, t, H3);      // Y3 = R (V - X3) - Y1 H^3
    f_mul(p.Z, p.Z, H);     // Z3 = Z1 H
}

// (X, Y, Z) <- (X, Y, Z) + (X2, +-Y2, Z2), both Jacobian: 11 products + 5 squares (add-2007-bl)
template <class F>
__device__ __noinline__ void jac_add(Jac<F>& p, const Jac<F>& q, bool negate) {
    if (f_is_zero(q.Z)) return;
    F y2 = q.Y;
    if (negate) {
        F z;
        f_zero(z);
        f_sub(y2, z, q.Y);
    }
    if (f_is_zero(p.Z)) {
        p.X = q.X;
        p.Y = y2;
        p.Z = q.Z;
        return;
    }
    F Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, Rr, V, t;
    f_sqr(Z1Z1, p.Z);
    f_sqr(Z2Z2, q.Z);
    f_mul(U1, p.X, Z2Z2);
    f_mul(U2, q.X, Z1Z1);
    f_mul(t, q.Z, Z2Z2);
    f_mul(S1, p.Y, t);
    f_mul(t, p.Z, Z1Z1);
    f_mul(S2, y2, t);
    f_sub(H, U2, U1);
    f_sub(Rr, S2, S1);
    if (f_is_zero(H)) {
        if (f_is_zero(Rr)) {
            jac_double(p);  // the same point
        } else {
            f_zero(p.X);    // opposite points
            f_zero(p.Y);
            f_zero(p.Z);
        }
        return;
    }
    f_add(Rr, Rr, Rr);      // r = 2 (S2 - S1)
    f_add(I, H, H);
    f_sqr(I, I);            // I = (2 H)^2
    f_mul(J, H, I);
    f_mul(V, U1, I);
    f_add(t, p.Z, q.Z);
    f_sqr(t, t);
    f_sub(t, t, Z1Z1);
    f_sub(t, t, Z2Z2);
    f_mul(p.Z, t, H);       // Z3 = ((Z1 + Z2)^2 - Z1Z1 - Z2Z2) H
    f_sqr(t, Rr);
    f_sub(t, t, J);
    f_sub(t, t, V);
    f_sub(p.X, t, V);       // X3 = r^2 - J - 2 V
    f_sub(t, V, p.X);
    f_mul(t, Rr, t);
    f_mul(S1, S1, J);
    f_add(S1, S1, S1);
    f_sub(p.Y, t, S1);      // Y3 = r (V - X3) - 2 S1 J
}

// pts: [2 NF][4][n] (x, y; NF = 1 Fq per coordinate on G1, 2 on G2), scalars: [4][n] plain 256-bit integers,
// out: [2 NF][4][n], inf[e] = 1 where the result is the point at infinity (coordinates zeroed).  An input of (0, 0) -
// the coordinates ark gives the identity - is the point at infinity.
template <class F>
__global__ void __launch_bounds__(128) bnp_scalar_mul_kernel(const u64* pts, const u64* scalars, u64* out, unsigned char* inf,
                                                             size_t n) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    F x, y;
    f_load(x, pts, 0, n, e);
    f_load(y, pts, 1, n, e);
    u64 k[4];
#pragma unroll
    for (int j = 0; j < 4; j++) k[j] = scalars[(size_t)j * n + e];
    Jac<F> acc;
    f_zero(acc.X);
    f_zero(acc.Y);
    f_zero(acc.Z);
    const bool base_inf = f_is_zero(x) && f_is_zero(y);
    if (!base_inf) {
        // T[j] = (j + 1) P
        Jac<F> T[8];
        T[0].X = x;
        T[0].Y = y;
        f_one(T[0].Z);
        T[1] = T[0];
        jac_double(T[1]);
#pragma unroll 1
        for (int j = 2; j < 8; j++) {
            T[j] = T[j - 1];
            jac_add_affine(T[j], x, y);
        }
        // k = sum d_i 16^i, d_i in [-8, 8): digit 64 is the carry out of the top nibble
        signed char d[65];
        u32 carry = 0;
#pragma unroll 1
        for (int i = 0; i < 64; i++) {
            u32 v = (u32)((k[i >> 4] >> ((i & 15) * 4)) & 15ull) + carry;
            carry = v >= 8u ? 1u : 0u;
            d[i] = (signed char)((int)v - (int)(carry << 4));
        }
        d[64] = (signed char)carry;
#pragma unroll 1
        for (int i = 64; i >= 0; i--) {
            if (i != 64) {
                jac_double(acc);
                jac_double(acc);
                jac_double(acc);
                jac_double(acc);
            }
            const int di = d[i];
            if (di != 0) jac_add(acc, T[(di < 0 ? -di : di) - 1], di < 0);
        }
    }
    const bool is_inf = f_is_zero(acc.Z);
    F ox, oy;
    if (is_inf) {
        f_zero(ox);
        f_zero(oy);
    } else {
        F zi, zi2;
        f_inv(zi, acc.Z);
        f_sqr(zi2, zi);
        f_mul(ox, acc.X, zi2);
        f_mul(zi2, zi2, zi);
        f_mul(oy, acc.Y, zi2);
    }
    f_store(out, 0, n, e, ox);
    f_store(out, 1, n, e, oy);
    inf[e] = is_inf ? 1 : 0;
}
