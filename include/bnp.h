/* bnp.h - C ABI of the B200-native batched BN254 pairing engine (libbnp.so).
 *
 * Drop-in boundary for the native path of qope/plonky2-bn254-pairing.  The reference has no FFI
 * layer of its own: the boundary is its public Rust function set, and each entry point below is
 * what a thin `extern "C"` crate binds for the function it names.
 *
 * Data layout (all entry points): structure-of-arrays by 64-bit limb, Montgomery form with
 * R = 2^256, canonical (< p) - i.e. exactly ark-ff's `Fp.0.0: [u64; 4]`.  An array of n elements
 * with K base-field values per element is   u64 buf[K][4][n]   (limb j of value k of element e at
 * buf[(k*4 + j)*n + e]).
 *     G1Affine : K = 2   (x, y)
 *     G2Affine : K = 4   (x.c0, x.c1, y.c0, y.c1)
 *     MyFq12   : K = 12  coeffs[0..11]; coeffs[i] + coeffs[i+6]*u is the Fq2 coefficient of w^i
 * For k-way products the k points of one product sit side by side: G1 K = 2k, G2 K = 4k.
 *
 * Preconditions are the reference's: points are non-identity, on-curve, in the r-torsion subgroup;
 * Fq12 inputs to the final exponentiation are non-zero (the reference panics on zero; this
 * library returns zeros for that element).
 *
 * All functions return 0 on success or a negative BNP_E* code; nothing unwinds across the ABI.
 * Host-pointer calls are synchronous.  `*_dev` calls take device pointers on the given device and
 * only enqueue work on `stream` (a cudaStream_t passed as void*; NULL = the library's own NON-BLOCKING stream, which is not
 * ordered with the legacy default stream: callers that work on the default stream pass cudaStreamLegacy, (void*)0x1).
 * There is no CPU fallback: without a CUDA device every compute call fails with BNP_ENODEV.
 */
#ifndef BNP_H
#define BNP_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNP_OK 0
#define BNP_EINVAL (-1)   /* bad argument */
#define BNP_ENODEV (-2)   /* no usable CUDA device / not initialised */
#define BNP_ECUDA (-3)    /* CUDA runtime error (see bnp_last_error) */
#define BNP_ENOMEM (-4)
#define BNP_EUNSUPPORTED (-5)
#define BNP_EMALFORMED (-6) /* bnp_eip197_pairing_check: a point of the input is not a valid encoding */

#define BNP_VARIANT_REFERENCE 0 /* exponent (p^12-1)/r, bit-equal to final_exp_native.rs:209 */
#define BNP_VARIANT_ARK 1       /* ark-ec 0.4.2 Bn254 final exponentiation (= variant 0 ^ 2x(6x^2+3x+1)) */

/* Create per-device state (programs, constants, scratch, stream) on the listed CUDA devices;
 * devices == NULL means device 0.  Idempotent. */
int bnp_init(const int* devices, int n_devices);
void bnp_shutdown(void);
int bnp_device_count(void); /* devices initialised by bnp_init */
/* How bnp_pairing_product moves the per-device partial products (384 bytes each) in a multi-device context: "nccl" (one
 * ncclAllGather over the communicators bnp_init created; libnccl is loaded at run time) or "peer-copy" (NCCL absent,
 * failed, or a single device). */
const char* bnp_gather_transport(void);
const char* bnp_strerror(int code);
const char* bnp_last_error(void); /* detail of the last BNP_ECUDA on this thread */

/* miller_loop_native(Q, P) for n independent pairs  (miller_loop_native.rs:320). */
int bnp_miller_loop_batch(const uint64_t* g1, const uint64_t* g2, uint64_t* out, size_t n);
/* multi_miller_loop_native(pairs) for n independent products of k pairs each, k in 1..4
 * (miller_loop_native.rs:324). */
int bnp_multi_miller_loop_batch(const uint64_t* g1, const uint64_t* g2, uint64_t* out, size_t n, int k);
/* final_exp_native(a)  (final_exp_native.rs:209). */
int bnp_final_exp_batch(const uint64_t* in, uint64_t* out, size_t n, int variant);
/* Witness sourcing for the reference's final-exponentiation CIRCUIT (final_exp_target.rs:65-185): one pass that
 * returns, per element, the five native MyFq12 values the circuit and its three Fq12ExpU64 starks consume,
 * out[60][4][n] = { m = easy part (offset BNP_WITNESS_M), m^x, m^(x^2), m^(x^3) (final_exp_target.rs:89-117, the
 * stark outputs with offset 1), final_exp_native(in) (final_exp_target.rs:240) }, each in MyFq12 coefficient order. */
#define BNP_WITNESS_M 0
#define BNP_WITNESS_MX 12
#define BNP_WITNESS_MX2 24
#define BNP_WITNESS_MX3 36
#define BNP_WITNESS_OUT 48
#define BNP_WITNESS_FQ 60
int bnp_final_exp_witness_batch(const uint64_t* in, uint64_t* out, size_t n);
/* pairing(p, q) = final_exp_native(miller_loop_native(&q, &p))  (pairing.rs:20), one fused launch.
 * Output in MyFq12 coefficient order; the Rust wrapper applies MyFq12 -> Fq12. */
int bnp_pairing_batch(const uint64_t* g1, const uint64_t* g2, uint64_t* out, size_t n, int variant);
/* n independent k-way products  prod_j pairing(p_j, q_j), k in 1..4 (shared squarings, one final
 * exponentiation per product): the Groth16-verify shape. */
int bnp_multi_pairing_batch(const uint64_t* g1, const uint64_t* g2, uint64_t* out, size_t n, int k, int variant);
/* ONE product over all n pairs: out[12][4] = final_exp(prod_i miller(q_i, p_i)).  The pairs are split
 * over the initialised devices; per-device partial products are combined on device 0. */
int bnp_pairing_product(const uint64_t* g1, const uint64_t* g2, uint64_t* out, size_t n, int variant);
/* frobenius_map_native(a, power), any power (reduced mod 12)  (final_exp_native.rs:17). */
int bnp_frobenius_batch(const uint64_t* in, uint64_t* out, size_t n, size_t power);
/* pow_native(a, exp)  (final_exp_native.rs:56): a^exp for ANY MyFq12 a (not only cyclotomic ones - the reference's
 * test_pow feeds a random element, :266-273); exp is `n_limbs` little-endian 64-bit limbs, the same for the whole
 * batch (the reference's Vec<u64>).  The reference's NAF walk (get_naf, :86-128) with its `res / a` for -1 digits;
 * a = 0 with a negative digit makes the reference panic (division by zero) - here that element comes back as 0.
 * exp = 0 returns a itself, like the reference (its accumulator starts at `a` and the loop never starts, :57,83). */
int bnp_pow_u64_batch(const uint64_t* in, uint64_t* out, size_t n, const uint64_t* exp, size_t n_limbs);
/* Input validation on the device: ok[e] = 1 iff the points of element e would be accepted by
 * `G1Affine::new` / `G2Affine::new` - on the curve (G1: y^2 = x^3 + 3, cofactor 1) resp. on the twist
 * y^2 = x^3 + 3/(9+u) AND in the r-torsion subgroup (ark-bn254 0.4's own test [6x^2]Q == psi(Q)).  This is the
 * assertion hidden behind miller_loop_native.rs:303,311 that the pairing entry points do NOT re-check: they compute
 * on whatever coordinates they are given.  g1 or g2 may be NULL (only the other group is checked); coordinates
 * (0, 0) - ark's encoding of the identity, which carries a separate flag there - are reported as invalid. */
int bnp_validate_batch(const uint64_t* g1 /* [2][4][n] */, const uint64_t* g2 /* [4][4][n] */, unsigned char* ok /* [n] */, size_t n);
/* Scalar multiplication on the device: out[e] = scalars[e] * pts[e] on G1 (group = 1, pts and out
 * [2][4][n]) or G2 (group = 2, [4][4][n]) - what `G1Affine * Fr` / `G2Affine * Fr` compute in the reference's tests
 * (test_to_one, final_exp_native.rs:245-250) and what a verifier folding many pairing checks into one by a random linear
 * combination needs in front of the pairing batch.  scalars: [4][n], each a PLAIN 256-bit integer in four little-endian
 * 64-bit limbs (ark's `Fr::into_bigint()`), any value below 2^256.  inf[e] = 1 where the result is the point at
 * infinity (its coordinates come back as (0, 0), ark's encoding); input coordinates (0, 0) are the point at infinity.
 * The points are NOT validated (bnp_validate_batch does that): the group law is applied to whatever on-curve
 * coordinates arrive. */
int bnp_scalar_mul_batch(int group, const uint64_t* pts, const uint64_t* scalars, uint64_t* out, unsigned char* inf /* [n] */, size_t n);
/* MyFq12 `Mul`: out = a * b element-wise. */
int bnp_fq12_mul_batch(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);

/* ---- device-pointer variants (one device; no host<->device copies; asynchronous on `stream`) ---- */
int bnp_miller_loop_dev(int device, void* stream, const uint64_t* g1, const uint64_t* g2, uint64_t* out, size_t n, int k);
/* Miller values up to a proper-subfield factor: only valid as input to a final exponentiation
 * (what bnp_pairing_product reduces and gathers). */
int bnp_miller_loop_fused_dev(int device, void* stream, const uint64_t* g1, const uint64_t* g2, uint64_t* out, size_t n);
int bnp_final_exp_dev(int device, void* stream, const uint64_t* in, uint64_t* out, size_t n, int variant);
int bnp_final_exp_witness_dev(int device, void* stream, const uint64_t* in, uint64_t* out /* [60][4][n] */, size_t n);
int bnp_pairing_dev(int device, void* stream, const uint64_t* g1, const uint64_t* g2, uint64_t* out, size_t n, int k, int variant);
int bnp_frobenius_dev(int device, void* stream, const uint64_t* in, uint64_t* out, size_t n, size_t power);
int bnp_pow_u64_dev(int device, void* stream, const uint64_t* in, uint64_t* out, size_t n, const uint64_t* exp, size_t n_limbs);
int bnp_validate_dev(int device, void* stream, const uint64_t* g1, const uint64_t* g2, unsigned char* ok /* device, [n] */, size_t n);
int bnp_scalar_mul_dev(int device, void* stream, int group, const uint64_t* pts, const uint64_t* scalars, uint64_t* out,
                       unsigned char* inf /* device, [n] */, size_t n);
int bnp_fq12_mul_dev(int device, void* stream, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* In-place tree product of n MyFq12 values (buf[12][4][n], destroyed) -> out[12][4][1]. */
int bnp_fq12_product_dev(int device, void* stream, uint64_t* buf, uint64_t* out, size_t n);

/* ---- introspection / measurement ---- */
/* Algorithmic work of one element of the named program ("pairing_v0", "miller", "final_exp_v0", ...):
 * 32x32->64 multiply-accumulates (64 per Fp product + 72 per Montgomery reduction), 0 if unknown. */
uint64_t bnp_program_macs(const char* program);
/* Multiply-accumulates the kernel actually issues for one element.  One thread runs a whole Karatsuba Fq2 operation,
 * so this equals bnp_program_macs (round 1's component-split kernel issued 15 % more). */
uint64_t bnp_program_macs_executed(const char* program);

/* ---- Prepared G2 points ------------------------------------------------------------------------------
 * The engine's `G2Prepared`: the 91 line-coefficient triples of one G2 point (BNP_PREP_FQ = 546 Fq per point, the same
 * list ark-ec's `G2Prepared::from` builds, in this engine's normalisation).  A pairing whose G2 point is prepared skips
 * all point arithmetic of the Miller loop; a Groth16 verifier prepares the three G2 points of its verifying key once.
 *   bnp_g2_prepare_batch:       g2 [4][4][n] -> coeffs [BNP_PREP_FQ][4][n]
 *   bnp_pairing_prepared_batch: out[i] = final_exp( prod_{j < kv} miller(g1[j], g2[j])  *  prod_{j < kp} miller(g1[kv + j], prepared[j]) )
 *       g1 [2 (kv + kp)][4][n], g2 [4 kv][4][n] (NULL when kv = 0); `prepared` is ONE element, [kp * BNP_PREP_FQ][4][1],
 *       shared by the whole batch.  Shapes in the library: (kv, kp) = (0, 1), (1, 2), (1, 3). */
#define BNP_PREP_FQ 546
int bnp_g2_prepare_batch(const uint64_t* g2, uint64_t* coeffs, size_t n);
int bnp_pairing_prepared_batch(const uint64_t* g1, const uint64_t* g2, const uint64_t* prepared, uint64_t* out, size_t n,
                               int kv, int kp, int variant);
int bnp_pairing_prepared_dev(int device, void* stream, const uint64_t* g1, const uint64_t* g2, const uint64_t* prepared,
                             uint64_t* out, size_t n, int kv, int kp, int variant);

/* ---- Wire formats decoded / encoded on the device -------------------------------------------------
 * The reference takes ark-bn254 values; what a caller holds are bytes.  `fmt`:
 *   BNP_WIRE_ARK_UNCOMPRESSED / BNP_WIRE_ARK_COMPRESSED  ark-serialize 0.4 CanonicalSerialize of G1Affine / G2Affine
 *       (little-endian canonical integers, Fq2 as c0 || c1, flags in the two top bits of the last byte: bit 7 "y is the
 *       larger of {y, -y}", bit 6 point at infinity); the compressed form recovers y by a square root on the device;
 *   BNP_WIRE_EIP197  EIP-196/197 words: 32-byte big-endian, G1 = x || y, G2 = x.c1 || x.c0 || y.c1 || y.c0, zero = infinity.
 * Element sizes: G1 64 / 32 / 64 bytes, G2 128 / 64 / 128 bytes.  Outputs are the SoA arrays the pairing entry points
 * take; status[i] is one of BNP_POINT_*, and the coordinates of an element whose status is not BNP_POINT_OK are zero.
 * `check_subgroup` adds the r-torsion test the reference hides in G2Affine::new (miller_loop_native.rs:303,311).
 * Host pointers; the work runs on the first device of bnp_init. */
#define BNP_WIRE_ARK_UNCOMPRESSED 0
#define BNP_WIRE_ARK_COMPRESSED 1
#define BNP_WIRE_EIP197 2
#define BNP_POINT_OK 0
#define BNP_POINT_INFINITY 1
#define BNP_POINT_NOT_CANONICAL 2   /* a coordinate >= p, or both flag bits set */
#define BNP_POINT_NOT_ON_CURVE 3    /* y^2 != x^3 + b, or (compressed) x^3 + b is not a square */
#define BNP_POINT_NOT_IN_SUBGROUP 4 /* G2: on the twist, outside the r-torsion */
int bnp_decode_g1_batch(int fmt, const uint8_t* in, size_t n, uint64_t* g1 /* [2][4][n] */, uint8_t* status /* [n] */);
int bnp_decode_g2_batch(int fmt, const uint8_t* in, size_t n, uint64_t* g2 /* [4][4][n] */, uint8_t* status /* [n] */,
                        int check_subgroup);
/* MyFq12 SoA [12][4][n] <-> ark-serialize bytes of the ark Fq12 the reference's `Into<Fq12>` gives (pairing.rs:21):
 * 384 bytes per element, c0.c0.c0, c0.c0.c1, c0.c1.c0, ... little-endian canonical integers. */
int bnp_encode_fq12_batch(const uint64_t* f12, size_t n, uint8_t* out /* 384 n bytes */);
int bnp_decode_fq12_batch(const uint8_t* in /* 384 n bytes */, size_t n, uint64_t* f12, uint8_t* status /* [n] */);
/* The Ethereum pairing precompile (address 0x08, EIP-197) on `k` pairs of 192 bytes: *result = 1 iff
 * e(P_1, Q_1) ... e(P_k, Q_k) == 1.  Decoding, curve and subgroup checks, Miller loops, product and final exponentiation
 * all run on the device; pairs with a point at infinity drop out; returns BNP_EMALFORMED (precompile failure) when a
 * point is not a canonical encoding, not on its curve, or (G2) not in the subgroup. */
int bnp_eip197_pairing_check(const uint8_t* in, size_t k, int* result);

/* Kernel launches issued by this library since bnp_init (for bench.py's gpu_launches). */
uint64_t bnp_launch_count(void);
/* Dependency-free IMAD.WIDE.U32 throughput microbenchmark on `device`: writes multiply-accumulates
 * per second (the roofline denominator) to *macs_per_s. */
int bnp_imad_peak(int device, double* macs_per_s);
/* Same loop with plain 32-bit IMAD (informational: the FMA pipe's nominal integer issue rate). */
int bnp_imad32_peak(int device, double* imads_per_s);
/* Run an arbitrary sequencer program by name on device arrays (test hook for op-level parity). */
int bnp_run_program_dev(int device, void* stream, const char* program, const uint64_t* g1, const uint64_t* g2,
                        const uint64_t* f12, const uint64_t* aux, uint64_t* out, size_t n);
/* Tuning knobs (0 keeps the current setting): threads per block (32/64/128/256/384/512; one thread per pairing; default 384 =
 * one block per SM, whose shared memory plus tensor memory hold 18 Fq2 slots per pairing), and whether
 * programs are run as
 * phase-split task queues (1 = automatic: when the unsplit batch would leave the last round of warp-tasks badly
 * filled, 2 = never, 3 = whenever the library has a split variant of the program). */
int bnp_set_launch_config(int threads_per_block, int phase_mode);
/* Threads per block the sequencer kernel is launched with (bench.py reports the kernel instance it timed). */
int bnp_threads_per_block(void);

#ifdef __cplusplus
}
#endif
#endif /* BNP_H */
