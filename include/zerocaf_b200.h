/*
 * zerocaf_b200.h -- C ABI of libzerocaf_b200.so: the batched, B200-native (sm_100a) backend for the arithmetic hot
 * path of dusk-network/dusk-zerocaf.
 *
 * The reference crate has NO FFI / plugin boundary: its seam is the backend type alias
 *     pub type FieldElement = backend::u64::field::FieldElement;     /root/reference/src/field.rs:83-91
 *     pub type Scalar       = backend::u64::scalar::Scalar;          /root/reference/src/scalar.rs:68-76
 * selected by a cargo feature (/root/reference/Cargo.toml:41-45, src/backend/mod.rs:9-16).  A `cuda_backend` feature
 * binds the entry points below (see INTEGRATION.md for the Rust `extern "C"` block and the trait impls).
 *
 * Data layout at the boundary == the reference's in-memory types:
 *   FieldElement / Scalar : uint64_t[5], radix-2^52 limbs, little-endian limb order, canonical (< modulus, limbs < 2^52)
 *                           /root/reference/src/backend/u64/field.rs:31-32, scalar.rs:26-27
 *   EdwardsPoint / RistrettoPoint : uint64_t[20] = X|Y|Z|T          /root/reference/src/edwards.rs:336-342,
 *                                                                    /root/reference/src/ristretto.rs:157-158
 * Arrays are AoS with 40-byte (field/scalar) and 160-byte (point) stride.  Inputs must be canonical; outputs are.
 *
 * Every function returns a status: 0 = ok, > 0 = argument error (ZC_ERR_*), < 0 = -(cudaError_t) / NCCL failure.
 * Nothing aborts or throws across the boundary (the reference panics instead: scalar.rs:465, field.rs:285).
 * There is no CPU fallback: without a CUDA device zc_ctx_create fails with a negative status.
 *
 * Plain functions take HOST pointers (synchronous; internally a chunked pipeline: H2D of chunk k+1, the kernel on chunk k
 * and D2H of chunk k-1 overlap when the host memory is pinned -- zc_host_alloc / zc_host_register).  `_dev` twins take DEVICE pointers,
 * enqueue on the context's stream and return without synchronising (zc_ctx_sync waits).  `out` may alias an input
 * exactly (in place) but must not partially overlap.  A context is bound to one device and one stream and is not
 * thread-safe; distinct contexts are independent.
 */
#ifndef ZEROCAF_B200_H
#define ZEROCAF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZC_OK                0
#define ZC_ERR_NULL          1   /* null pointer argument */
#define ZC_ERR_SIZE          2   /* n too large / zero where not allowed */
#define ZC_ERR_MODE          3   /* unknown mode / window */
#define ZC_ERR_NONCANONICAL  4   /* validation requested and an input is not canonical */
#define ZC_ERR_STATE         5   /* context misuse (e.g. sharded call without a communicator) */

#define ZC_SCALAR_MUL_STRICT 0   /* LSB-first double-and-add with add-as-double: limb-exact vs edwards.rs:102-120 */
#define ZC_SCALAR_MUL_FAST   1   /* signed fixed-window, dedicated doubling: same group element, other representative */

typedef struct zc_ctx zc_ctx;

/* ---- context ------------------------------------------------------------------------------------------- */
const char *zc_version(void);
/* stream: a cudaStream_t to enqueue on, or NULL for a stream owned by the context. */
int32_t zc_ctx_create(int32_t device, void *stream, zc_ctx **out);
int32_t zc_ctx_destroy(zc_ctx *ctx);
int32_t zc_ctx_sync(zc_ctx *ctx);
const char *zc_last_error_string(zc_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t zc_ctx_launch_count(zc_ctx *ctx);
/* pinned host memory for callers that want full-speed copies */
int32_t zc_host_alloc(size_t bytes, void **out);
int32_t zc_host_free(void *p);
/* pin / unpin memory the caller already owns (e.g. a Rust Vec<FieldElement>) so the host-pointer entry points can
 * overlap their H2D and D2H copies with the kernels */
int32_t zc_host_register(void *p, size_t bytes);
int32_t zc_host_unregister(void *p);

/* ---- FieldElement batch ops: out[i] = a[i] (op) b[i] mod p ------------------------------------------------ */
/* replaces Mul  field.rs:250-275 */
int32_t zc_fe_mul_batch(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
/* replaces Square field.rs:302-315 */
int32_t zc_fe_square_batch(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
/* replaces Add  field.rs:191-215 */
int32_t zc_fe_add_batch(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
/* replaces Sub  field.rs:217-248 */
int32_t zc_fe_sub_batch(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
/* replaces Neg  field.rs:170-189 */
int32_t zc_fe_neg_batch(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
/* BASELINE config 2: prod[i] = a[i]*b[i], sq[i] = a[i]^2, both fully reduced, one launch */
int32_t zc_fe_mul_square_batch(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *prod, uint64_t *sq, size_t n);

int32_t zc_fe_mul_batch_dev(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
int32_t zc_fe_square_batch_dev(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
int32_t zc_fe_add_batch_dev(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
int32_t zc_fe_sub_batch_dev(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
int32_t zc_fe_neg_batch_dev(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
int32_t zc_fe_mul_square_batch_dev(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *prod, uint64_t *sq, size_t n);

/* The same on the 32-byte wire format (FieldElement::to_bytes / from_bytes, field.rs:563-631: little-endian, canonical):
 * the kernel moves exactly the algorithmic 128 bytes per pair and a host call 20 % fewer PCIe bytes than the limb layout.
 * Arrays must be 16-byte aligned. */
int32_t zc_fe_mul_square_batch_packed(zc_ctx *ctx, const uint8_t *a, const uint8_t *b, uint8_t *prod, uint8_t *sq, size_t n);
int32_t zc_fe_mul_square_batch_packed_dev(zc_ctx *ctx, const uint8_t *a, const uint8_t *b, uint8_t *prod, uint8_t *sq, size_t n);
/* replaces Div field.rs:277-299: out[i] = a[i] * b[i]^-1 (the reference asserts b != 0, field.rs:285; here a / 0 = 0) */
int32_t zc_fe_div_batch(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
int32_t zc_fe_div_batch_dev(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);

/* ---- input validation (opt-in) -------------------------------------------------------------------------------------
 * The kernels assume canonical inputs (limbs < 2^52, value < modulus); the reference's types expose their limbs and its own
 * tests build non-canonical values (field.rs:1160-1167, 1193-1200).  zc_*_check_canonical_batch checks an array on the
 * device: ZC_OK, or ZC_ERR_NONCANONICAL with the index of the first offending element in *first_bad (may be NULL).
 * zc_ctx_set_validation(ctx, 1) makes the hot-path entry points do that check on their own inputs -- field / scalar /
 * point element-wise ops, mul_square, scalar multiplication, every MSM call -- and return ZC_ERR_NONCANONICAL (the outputs
 * are then unspecified); `_dev` calls become synchronous while it is on. */
int32_t zc_ctx_set_validation(zc_ctx *ctx, int32_t on);
int32_t zc_fe_check_canonical_batch(zc_ctx *ctx, const uint64_t *a, size_t n, uint64_t *first_bad);
int32_t zc_fe_check_canonical_batch_dev(zc_ctx *ctx, const uint64_t *a, size_t n, uint64_t *first_bad);
int32_t zc_scalar_check_canonical_batch(zc_ctx *ctx, const uint64_t *a, size_t n, uint64_t *first_bad);
int32_t zc_scalar_check_canonical_batch_dev(zc_ctx *ctx, const uint64_t *a, size_t n, uint64_t *first_bad);
int32_t zc_point_check_canonical_batch(zc_ctx *ctx, const uint64_t *p, size_t n, uint64_t *first_bad);
int32_t zc_point_check_canonical_batch_dev(zc_ctx *ctx, const uint64_t *p, size_t n, uint64_t *first_bad);

/* ---- Scalar batch ops (mod L): replaces scalar.rs Mul :247-270, Square :272-283, Add :184-208, Sub :210-245, Neg --- */
int32_t zc_scalar_mul_batch(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
int32_t zc_scalar_square_batch(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
int32_t zc_scalar_add_batch(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
int32_t zc_scalar_sub_batch(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
int32_t zc_scalar_neg_batch(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);

int32_t zc_scalar_mul_batch_dev(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
int32_t zc_scalar_square_batch_dev(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
int32_t zc_scalar_add_batch_dev(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
int32_t zc_scalar_sub_batch_dev(zc_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n);
int32_t zc_scalar_neg_batch_dev(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);

/* ---- EdwardsPoint / RistrettoPoint batch ops (limb-exact representatives) ---------------------------------- */
/* replaces Add edwards.rs:465-501 / ristretto.rs:248-276 */
int32_t zc_point_add_batch(zc_ctx *ctx, const uint64_t *p, const uint64_t *q, uint64_t *out, size_t n);
/* replaces Sub edwards.rs:503-545 / ristretto.rs:278-312 */
int32_t zc_point_sub_batch(zc_ctx *ctx, const uint64_t *p, const uint64_t *q, uint64_t *out, size_t n);
/* replaces Double edwards.rs:579-592 / ristretto.rs:314-328  (= P + P with the Add formulas) */
int32_t zc_point_double_batch(zc_ctx *ctx, const uint64_t *p, uint64_t *out, size_t n);
/* replaces Neg edwards.rs:440-463 / ristretto.rs:224-246 */
int32_t zc_point_neg_batch(zc_ctx *ctx, const uint64_t *p, uint64_t *out, size_t n);
/* replaces Mul<&Scalar> edwards.rs:547-577 / ristretto.rs:330-392 (double_and_add edwards.rs:102-120) */
int32_t zc_point_scalar_mul_batch(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, uint64_t *out, size_t n, int32_t mode);
/* Ristretto equality ristretto.rs:166-176: eq[i] = 1 iff X1*Y2 == Y1*X2 or X1*X2 == Y1*Y2 */
int32_t zc_ristretto_eq_batch(zc_ctx *ctx, const uint64_t *p, const uint64_t *q, uint8_t *eq, size_t n);

int32_t zc_point_add_batch_dev(zc_ctx *ctx, const uint64_t *p, const uint64_t *q, uint64_t *out, size_t n);
int32_t zc_point_sub_batch_dev(zc_ctx *ctx, const uint64_t *p, const uint64_t *q, uint64_t *out, size_t n);
int32_t zc_point_double_batch_dev(zc_ctx *ctx, const uint64_t *p, uint64_t *out, size_t n);
int32_t zc_point_neg_batch_dev(zc_ctx *ctx, const uint64_t *p, uint64_t *out, size_t n);
int32_t zc_point_scalar_mul_batch_dev(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, uint64_t *out, size_t n, int32_t mode);
int32_t zc_ristretto_eq_batch_dev(zc_ctx *ctx, const uint64_t *p, const uint64_t *q, uint8_t *eq, size_t n);

/* Fixed-base scalar multiplication out[i] = [s_i] B for the basepoint (constants.rs:188-211): replaces
 * `&BASEPOINT * &scalar` (edwards.rs:547-577) and is the working equivalent of the reference's untested fixed-base
 * window_naf_mul (edwards.rs:155-171).  Signed radix-16 digits over a 48 KiB table of affine cached multiples built on the
 * device on first use; same group element as double_and_add, other representative (compare canonically). */
int32_t zc_basepoint_mul_batch(zc_ctx *ctx, const uint64_t *scalars, uint64_t *out, size_t n);
int32_t zc_basepoint_mul_batch_dev(zc_ctx *ctx, const uint64_t *scalars, uint64_t *out, size_t n);

/* ---- canonicalisation: the wire-format step right after the hot path ----------------------
 * The reference computes these with data-dependent loops (Savas-Koc inverse, Tonelli-Shanks); each returns a uniquely
 * defined value, evaluated here with fixed exponent chains (a^(p-2); the p = 5 mod 8 square-root-ratio recipe), so the
 * outputs are bit-identical.  out_bytes of the _dev variant must be 16-byte aligned. */
/* replaces FieldElement::inverse field.rs:854-925 (the reference panics on 0, field.rs:864; here inverse(0) = 0) */
int32_t zc_fe_invert_batch(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
int32_t zc_fe_invert_batch_dev(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
/* replaces AffinePoint::from(EdwardsPoint) edwards.rs:1085-1092: out_xy[i] = (X/Z, Y/Z) as uint64_t[10] */
int32_t zc_point_to_affine_batch(zc_ctx *ctx, const uint64_t *p, uint64_t *out_xy, size_t n);
int32_t zc_point_to_affine_batch_dev(zc_ctx *ctx, const uint64_t *p, uint64_t *out_xy, size_t n);
/* replaces RistrettoPoint::compress ristretto.rs:398-425: out_bytes[i] = the 32-byte CompressedRistretto */
int32_t zc_ristretto_compress_batch(zc_ctx *ctx, const uint64_t *p, uint8_t *out_bytes, size_t n);
int32_t zc_ristretto_compress_batch_dev(zc_ctx *ctx, const uint64_t *p, uint8_t *out_bytes, size_t n);
/* replaces CompressedRistretto::decompress ristretto.rs:96-154 (the step right before the path):
 * ok[i] = 1 and out_points[i] = (x, y, 1, xy) where the reference returns Some(point); ok[i] = 0 and a zeroed point where
 * it returns None (negative or non-canonical s, non-square, negative t, y = 0).  in_bytes must be 4-byte aligned (_dev). */
int32_t zc_ristretto_decompress_batch(zc_ctx *ctx, const uint8_t *in_bytes, uint64_t *out_points, uint8_t *ok, size_t n);
int32_t zc_ristretto_decompress_batch_dev(zc_ctx *ctx, const uint8_t *in_bytes, uint64_t *out_points, uint8_t *ok, size_t n);
/* Hash to group, limb-exact: the returned (X:Y:Z:T) are the reference's own products.
 * replaces RistrettoPoint::elligator_ristretto_flavor ristretto.rs:430-471 (r0 as [u64;5] limbs, value < 2^256) */
int32_t zc_ristretto_elligator_batch(zc_ctx *ctx, const uint64_t *r0, uint64_t *out_points, size_t n);
int32_t zc_ristretto_elligator_batch_dev(zc_ctx *ctx, const uint64_t *r0, uint64_t *out_points, size_t n);
/* replaces RistrettoPoint::from_uniform_bytes ristretto.rs:493-507: 64 bytes per point, Elligator on each half, R_1 + R_2 */
int32_t zc_ristretto_from_uniform_bytes_batch(zc_ctx *ctx, const uint8_t *in_bytes, uint64_t *out_points, size_t n);
int32_t zc_ristretto_from_uniform_bytes_batch_dev(zc_ctx *ctx, const uint8_t *in_bytes, uint64_t *out_points, size_t n);
/* replaces ValidityCheck for EdwardsPoint edwards.rs:393-400, 733-748: ok[i] = 1 iff (aX^2 + Y^2) Z^2 == Z^4 + d X^2 Y^2 */
int32_t zc_point_is_valid_batch(zc_ctx *ctx, const uint64_t *p, uint8_t *ok, size_t n);
int32_t zc_point_is_valid_batch_dev(zc_ctx *ctx, const uint64_t *p, uint8_t *ok, size_t n);

/* ---- vector operations around the MSM: what an inner-product argument does to its scalar vectors
 * between MSMs, the signed-digit recodings, and the 32-byte wire format of FieldElement / Scalar -------------------------
 * pow:  out[i] = a[i]^e[i] mod m, 0^0 = 1                   (Pow: field.rs:334-354, scalar.rs:293-322)
 * half: out[i] = a[i] * 2^-1 mod m                          (Half: field.rs:317-323, scalar.rs:285-291)
 * to_bytes / from_bytes: [u64;5] limbs <-> 32 little-endian bytes (field.rs:563-631, scalar.rs:445-516).  FieldElement
 *   from_bytes keeps all 256 bits like the reference; Scalar::from_bytes asserts value <= L - 1 -- here ok[i] = 0 instead
 *   of a panic (the limbs are still written).  Byte pointers must be 4-byte aligned on the device.
 * window_naf: 256 signed digits (i8, least significant first) per scalar, width in 2..7; width 2 is compute_NAF
 *   (scalar.rs:370-415).
 * sqrt_ratio_i: (was_square[i], out[i]) = FieldElement::sqrt_ratio_i(u[i], v[i])  (field.rs:443-491). */
int32_t zc_fe_pow_batch(zc_ctx *ctx, const uint64_t *a, const uint64_t *e, uint64_t *out, size_t n);
int32_t zc_fe_pow_batch_dev(zc_ctx *ctx, const uint64_t *a, const uint64_t *e, uint64_t *out, size_t n);
int32_t zc_scalar_pow_batch(zc_ctx *ctx, const uint64_t *a, const uint64_t *e, uint64_t *out, size_t n);
int32_t zc_scalar_pow_batch_dev(zc_ctx *ctx, const uint64_t *a, const uint64_t *e, uint64_t *out, size_t n);
int32_t zc_fe_half_batch(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
int32_t zc_fe_half_batch_dev(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
int32_t zc_scalar_half_batch(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
int32_t zc_scalar_half_batch_dev(zc_ctx *ctx, const uint64_t *a, uint64_t *out, size_t n);
int32_t zc_fe_to_bytes_batch(zc_ctx *ctx, const uint64_t *a, uint8_t *out_bytes, size_t n);
int32_t zc_fe_to_bytes_batch_dev(zc_ctx *ctx, const uint64_t *a, uint8_t *out_bytes, size_t n);
int32_t zc_scalar_to_bytes_batch(zc_ctx *ctx, const uint64_t *a, uint8_t *out_bytes, size_t n);
int32_t zc_scalar_to_bytes_batch_dev(zc_ctx *ctx, const uint64_t *a, uint8_t *out_bytes, size_t n);
int32_t zc_fe_from_bytes_batch(zc_ctx *ctx, const uint8_t *in_bytes, uint64_t *out, size_t n);
int32_t zc_fe_from_bytes_batch_dev(zc_ctx *ctx, const uint8_t *in_bytes, uint64_t *out, size_t n);
int32_t zc_scalar_from_bytes_batch(zc_ctx *ctx, const uint8_t *in_bytes, uint64_t *out, uint8_t *ok, size_t n);
int32_t zc_scalar_from_bytes_batch_dev(zc_ctx *ctx, const uint8_t *in_bytes, uint64_t *out, uint8_t *ok, size_t n);
int32_t zc_scalar_window_naf_batch(zc_ctx *ctx, const uint64_t *a, int32_t width, int8_t *out_digits, size_t n);
int32_t zc_scalar_window_naf_batch_dev(zc_ctx *ctx, const uint64_t *a, int32_t width, int8_t *out_digits, size_t n);
/* into_bits: the 256 bits of each scalar, least significant first, one byte per bit (Scalar::into_bits, scalar.rs:352-366);
 * out_bits must be 16-byte aligned on the device */
int32_t zc_scalar_into_bits_batch(zc_ctx *ctx, const uint64_t *a, uint8_t *out_bits, size_t n);
int32_t zc_scalar_into_bits_batch_dev(zc_ctx *ctx, const uint64_t *a, uint8_t *out_bits, size_t n);
int32_t zc_fe_sqrt_ratio_i_batch(zc_ctx *ctx, const uint64_t *u, const uint64_t *v, uint64_t *out, uint8_t *was_square, size_t n);
int32_t zc_fe_sqrt_ratio_i_batch_dev(zc_ctx *ctx, const uint64_t *u, const uint64_t *v, uint64_t *out, uint8_t *was_square, size_t n);

/* ---- multi-scalar multiplication  out = sum_i [s_i] P_i  (new capability; the reference has none) ---
 * Semantics = fold(Add, identity, [double_and_add(P_i, s_i)]) (edwards.rs:102-120, 465-489) as a GROUP ELEMENT: the
 * returned (X:Y:Z:T) is a valid representative, compare with affine / Ristretto equality.  Pippenger with signed
 * `window_bits`-bit digits (8..16); n = 0 returns the identity (0,1,1,0). */
int32_t zc_msm(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits, uint64_t *out_point);
int32_t zc_msm_dev(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits, uint64_t *out_point_dev);

/* Fixed generators (bulletproofs G_i, H_i): an opaque handle that OWNS everything derived from the points.
 *   kind = ZC_GEN_PREPARED    the points normalised to Z = 1 in cached form (one field inversion per point, 128 B per
 *                             point): every bucket addition of a later MSM is a 7-multiplication mixed addition and the
 *                             per-call operand pass disappears.  window_bits / rank / nranks are ignored; the handle
 *                             serves any window size and any sharding.
 *   kind = ZC_GEN_FIXED_BASE  memory for time: the FIXED-BASE TABLES  2^(window_bits * w) * P_i  (same 128-B record) for
 *                             the windows the rank owns (zc_msm_plan_query) -- 16 x n rows at window_bits = 16 on one GPU (2 GiB for
 *                             2^20 points), 2 x n rows per rank on 8.  An MSM of exactly that shape then treats every
 *                             (window, point) digit as an entry of ONE bucket set: a single bucket reduction per rank and
 *                             no doubling chain at all (the serial tail that limits the multi-GPU scaling of plain
 *                             Pippenger).  One-time cost: window_bits * w_max doublings + one inversion per row.  A call
 *                             with another (window_bits, rank, nranks) returns ZC_ERR_MODE.
 * The point array is read during creation only (the call synchronises the stream before it returns): the caller may
 * free, reuse or overwrite it afterwards -- nothing is keyed by its address, so a recycled allocation cannot alias an
 * old generator set.  Several handles can be alive at once (G and H vectors).  A handle belongs to the context that
 * created it; destroy it before the context.  Results are the same group element as the plain calls. */
#define ZC_GEN_PREPARED    1
#define ZC_GEN_FIXED_BASE  2
typedef struct zc_msm_generators zc_msm_generators;
int32_t zc_msm_generators_create_dev(zc_ctx *ctx, const uint64_t *points_dev, size_t n, int32_t kind, int32_t window_bits,
                                     int32_t rank, int32_t nranks, zc_msm_generators **out);
int32_t zc_msm_generators_destroy(zc_ctx *ctx, zc_msm_generators *gens);
/* n, kind and the device bytes the handle owns (any out pointer may be NULL) */
int32_t zc_msm_generators_info(const zc_msm_generators *gens, size_t *n, int32_t *kind, size_t *device_bytes);
/* sum_i [s_i] G_i over the handle's n generators; scalars_dev holds n scalars.  _partial / _sharded as below. */
int32_t zc_msm_gen_dev(zc_ctx *ctx, const zc_msm_generators *gens, const uint64_t *scalars_dev, int32_t window_bits, uint64_t *out_point_dev);
int32_t zc_msm_gen_partial_dev(zc_ctx *ctx, const zc_msm_generators *gens, const uint64_t *scalars_dev, int32_t window_bits,
                               int32_t rank, int32_t nranks, uint64_t *out_point_dev);
int32_t zc_msm_gen_sharded_dev(zc_ctx *ctx, const zc_msm_generators *gens, const uint64_t *scalars_dev, int32_t window_bits, uint64_t *out_point_dev);

/* The sharding plan as a host function (no device needed): out[0] = the rank that owns `window` among nranks (boustrophedon:
 * windows 0..R-1 -> ranks 0..R-1, R..2R-1 -> ranks R-1..0, ... so the rank with the highest window also has the lowest),
 * out[1] = sub-bucket bits of a short top window (plain path), out[2] = spread bits of a short window in the fixed-base
 * path, out[3] = s such that a fixed-base table row of this window is 2^s P_i. */
int32_t zc_msm_plan_query(int32_t window_bits, int32_t window, int32_t nranks, int32_t out[4]);

/* Bucket-window-sharded MSM: a collective, every rank calls it with the same (points, scalars, n, window_bits), all
 * resident on its own device.  Rank r accumulates the windows it owns (zc_msm_plan_query), scales its window sums and folds them to
 * one partial point; the partial points are exchanged ONCE -- over NVLink peer memory (zc_peer_mailbox_*) or with one
 * ncclAllGather on the context's stream -- and folded in a fixed order on every rank, so all ranks return identical bits.  `nccl_comm` is an ncclComm_t created by the caller
 * (e.g. from an ncclUniqueId distributed with torch.distributed); libnccl is resolved at run time with dlopen. */
int32_t zc_ctx_set_nccl(zc_ctx *ctx, void *nccl_comm, int32_t rank, int32_t nranks);
int32_t zc_nccl_unique_id(uint8_t id_out[128]);
int32_t zc_nccl_comm_init(const uint8_t id[128], int32_t rank, int32_t nranks, void **comm_out);
int32_t zc_nccl_comm_destroy(void *comm);
/* Exchange over NVLink peer memory instead of NCCL: every rank creates a mailbox in its own HBM (zc_peer_mailbox_create
 * returns its 64-byte CUDA IPC handle), the handles of all ranks are gathered with the host framework's transport and
 * passed, in rank order, to zc_peer_mailbox_connect.  zc_msm_sharded_dev then delivers the partial points with peer
 * stores, waits on flags and folds them in one kernel (a fixed tree: identical bits on all ranks).  Slots and flags are
 * double-buffered by sequence parity, so back-to-back exchanges cannot overwrite a partial a slower peer still reads; the
 * flag wait is bounded (ZC_PEER_TIMEOUT_MS, default 5000): if a rank never arrives the kernel returns the identity and
 * zc_ctx_sync / zc_msm_sharded report ZC_ERR_STATE naming the missing rank instead of hanging. */
int32_t zc_peer_mailbox_create(zc_ctx *ctx, uint8_t handle_out[64]);
int32_t zc_peer_mailbox_connect(zc_ctx *ctx, const uint8_t *handles /* nranks x 64 */, int32_t rank, int32_t nranks);
/* One process driving several contexts (several GPUs with peer access, or several streams of one GPU): the mailboxes are
 * plain device pointers -- zc_peer_mailbox_create(ctx, scratch), zc_peer_mailbox_ptr on every context, then
 * zc_peer_mailbox_connect_local with the nranks pointers in rank order.  The calls of one collective may be enqueued from
 * a single host thread, rank after rank (they do not block the host). */
int32_t zc_peer_mailbox_ptr(zc_ctx *ctx, void **out);
int32_t zc_peer_mailbox_connect_local(zc_ctx *ctx, void *const *mailboxes /* nranks */, int32_t rank, int32_t nranks);
int32_t zc_msm_sharded_dev(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits, uint64_t *out_point_dev);
/* the same collective with HOST pointers (every rank passes the same arrays; synchronous; the copies are pipelined with
 * the operand pass) */
int32_t zc_msm_sharded(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits, uint64_t *out_point);
/* the local half only (no exchange): rank r's partial point, for tests of the sharding logic without NCCL */
int32_t zc_msm_partial_dev(zc_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, int32_t window_bits,
                           int32_t rank, int32_t nranks, uint64_t *out_point_dev);
/* fold k partial points in index order with the Add formulas (edwards.rs:465-489) */
int32_t zc_point_fold_dev(zc_ctx *ctx, const uint64_t *points_dev, size_t k, uint64_t *out_point_dev);

#ifdef __cplusplus
}
#endif
#endif /* ZEROCAF_B200_H */
