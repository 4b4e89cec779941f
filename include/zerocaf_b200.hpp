// zerocaf_b200.hpp -- C++ host mirror of the reference crate's type and operator surface, over the C ABI of
// libzerocaf_b200.so (include/zerocaf_b200.h).  Header-only; link with -lzerocaf_b200.
//
// The reference is a compiled (Rust) library whose toolchain is not present in this image, so the host side above the C
// ABI is written in C++ with the reference's names, argument meaning and error behaviour:
//   FieldElement    /root/reference/src/backend/u64/field.rs:31-32   (+ - * neg :170-275, Square :302-315, Identity :78-87,
//                                                                    from_bytes / to_bytes :563-631, inverse :854-925)
//   Scalar          /root/reference/src/backend/u64/scalar.rs:26-27  (+ - * :184-270, Square :272-283, from_bytes :445-467
//                                                                    panics above L-1 -> throws here)
//   EdwardsPoint    /root/reference/src/edwards.rs:336-342           (Identity :381-391, Neg :440-455, Add :465-501,
//                                                                    Sub :503-545, Mul<Scalar> :547-577, Double :579-592,
//                                                                    PartialEq = affine equality :360-364, is_valid :393-400)
//   RistrettoPoint  /root/reference/src/ristretto.rs:157-158         (forwarding operators :224-392, ct_eq :166-176,
//                                                                    compress :398-425, decompress :96-154,
//                                                                    from_uniform_bytes :493-507)
// The element types have the reference's in-memory layout ([u64;5] limbs, X|Y|Z|T), so arrays of them are passed to the
// ABI as they are.  Single-element operators run a one-element batch on the default device context -- faithful, but a
// kernel launch per call; hot loops use the slice functions in zerocaf::batch (the capability the reference lacks).
// Where the reference panics (assert! / unwrap) these throw zerocaf::Error; where it returns Option, std::optional.
#pragma once
#include <array>
#include <utility>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "zerocaf_b200.h"

namespace zerocaf {

struct Error : std::runtime_error {
  int32_t status;
  Error(int32_t st, const std::string& what) : std::runtime_error(what), status(st) {}
};

// One device context (one GPU, one stream).  Not thread-safe, like the C context it owns.
class Gpu {
 public:
  explicit Gpu(int device = 0) {
    int32_t st = zc_ctx_create(device, nullptr, &ctx_);
    if (st != ZC_OK) throw Error(st, "zc_ctx_create failed (no CUDA device? there is no CPU fallback), status " + std::to_string(st));
  }
  ~Gpu() { if (ctx_) zc_ctx_destroy(ctx_); }
  Gpu(const Gpu&) = delete;
  Gpu& operator=(const Gpu&) = delete;
  zc_ctx* ctx() const { return ctx_; }
  void check(int32_t st) const {
    if (st != ZC_OK) throw Error(st, std::string("zerocaf_b200: ") + zc_last_error_string(ctx_) + " (status " + std::to_string(st) + ")");
  }
  static Gpu& instance() { static Gpu g(0); return g; }

 private:
  zc_ctx* ctx_ = nullptr;
};

struct FieldElement;
struct Scalar;
struct EdwardsPoint;
struct RistrettoPoint;

namespace detail {
inline void limbs_from_bytes(const uint8_t b[32], uint64_t l[5]) {     // field.rs:563-587: keeps all 256 bits
  auto load8 = [&](int off) { uint64_t v = 0; for (int i = 7; i >= 0; i--) v = (v << 8) | (off + i < 32 ? b[off + i] : 0); return v; };
  const uint64_t m = (1ull << 52) - 1;
  l[0] = load8(0) & m; l[1] = (load8(6) >> 4) & m; l[2] = (load8(12) >> 8) & m; l[3] = (load8(19) >> 4) & m; l[4] = (load8(24) >> 16) & m;
}
inline void limbs_to_bytes(const uint64_t l[5], uint8_t out[32]) {      // field.rs:591-631
  unsigned __int128 acc = 0; int bits = 0, k = 0;
  for (int i = 0; i < 5; i++) {
    acc |= (unsigned __int128)l[i] << bits; bits += 52;
    while (bits >= 8 && k < 32) { out[k++] = (uint8_t)acc; acc >>= 8; bits -= 8; }
  }
  while (k < 32) { out[k++] = (uint8_t)acc; acc >>= 8; }
}
}  // namespace detail

// ---- FieldElement ----------------------------------------------------------------------------------------------
struct FieldElement {
  uint64_t l[5];
  static FieldElement zero() { return {{0, 0, 0, 0, 0}}; }
  static FieldElement one() { return {{1, 0, 0, 0, 0}}; }
  static FieldElement identity() { return one(); }                       // field.rs:78-87
  static FieldElement minus_one() { return -one(); }
  static FieldElement from_bytes(const uint8_t b[32]) { FieldElement r; detail::limbs_from_bytes(b, r.l); return r; }
  std::array<uint8_t, 32> to_bytes() const { std::array<uint8_t, 32> o; detail::limbs_to_bytes(l, o.data()); return o; }
  friend FieldElement operator+(const FieldElement& a, const FieldElement& b) { FieldElement r; Gpu::instance().check(zc_fe_add_batch(Gpu::instance().ctx(), a.l, b.l, r.l, 1)); return r; }
  friend FieldElement operator-(const FieldElement& a, const FieldElement& b) { FieldElement r; Gpu::instance().check(zc_fe_sub_batch(Gpu::instance().ctx(), a.l, b.l, r.l, 1)); return r; }
  friend FieldElement operator*(const FieldElement& a, const FieldElement& b) { FieldElement r; Gpu::instance().check(zc_fe_mul_batch(Gpu::instance().ctx(), a.l, b.l, r.l, 1)); return r; }
  // Div (field.rs:277-299): x * y^-1; the reference asserts y != 0 (:285)
  friend FieldElement operator/(const FieldElement& a, const FieldElement& b) {
    if (b == zero()) throw Error(ZC_ERR_NONCANONICAL, "Cannot divide by zero.");
    FieldElement r; Gpu::instance().check(zc_fe_div_batch(Gpu::instance().ctx(), a.l, b.l, r.l, 1)); return r;
  }
  FieldElement operator-() const { FieldElement r; Gpu::instance().check(zc_fe_neg_batch(Gpu::instance().ctx(), l, r.l, 1)); return r; }
  FieldElement square() const { FieldElement r; Gpu::instance().check(zc_fe_square_batch(Gpu::instance().ctx(), l, r.l, 1)); return r; }
  FieldElement pow(const FieldElement& e) const { FieldElement r; Gpu::instance().check(zc_fe_pow_batch(Gpu::instance().ctx(), l, e.l, r.l, 1)); return r; }   // field.rs:334-354
  FieldElement half() const { FieldElement r; Gpu::instance().check(zc_fe_half_batch(Gpu::instance().ctx(), l, r.l, 1)); return r; }                          // field.rs:317-323
  // field.rs:443-491: (was_square, root)
  static std::pair<bool, FieldElement> sqrt_ratio_i(const FieldElement& u, const FieldElement& v) {
    FieldElement r; uint8_t sq = 0;
    Gpu::instance().check(zc_fe_sqrt_ratio_i_batch(Gpu::instance().ctx(), u.l, v.l, r.l, &sq, 1));
    return {sq != 0, r};
  }
  FieldElement inverse() const {                                          // field.rs:854-925, panics on zero (:864)
    if (*this == zero()) throw Error(ZC_ERR_NONCANONICAL, "FieldElement::inverse of zero");
    FieldElement r; Gpu::instance().check(zc_fe_invert_batch(Gpu::instance().ctx(), l, r.l, 1)); return r;
  }
  friend bool operator==(const FieldElement& a, const FieldElement& b) { return a.to_bytes() == b.to_bytes(); }   // src/field.rs:93-106
  friend bool operator!=(const FieldElement& a, const FieldElement& b) { return !(a == b); }
};
static_assert(sizeof(FieldElement) == 40, "FieldElement must be [u64;5]");

// ---- Scalar ----------------------------------------------------------------------------------------------------
struct Scalar {
  uint64_t l[5];
  static Scalar zero() { return {{0, 0, 0, 0, 0}}; }
  static Scalar one() { return {{1, 0, 0, 0, 0}}; }
  static Scalar from_u64(uint64_t v) { return {{v & ((1ull << 52) - 1), v >> 52, 0, 0, 0}}; }
  // scalar.rs:445-467: asserts the value is <= L - 1
  static Scalar from_bytes(const uint8_t b[32]) {
    static const uint8_t LM1[32] = {0x62, 0xc8, 0x5f, 0x75, 0x6f, 0x03, 0xb4, 0x6a, 0x93, 0xd5, 0x2f, 0x82, 0x4d, 0xc7, 0xe6, 0x0a,
                                    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x02};
    for (int i = 31; i >= 0; i--) { if (b[i] < LM1[i]) break; if (b[i] > LM1[i]) throw Error(ZC_ERR_NONCANONICAL, "Scalar::from_bytes: value above L - 1"); }
    Scalar r; detail::limbs_from_bytes(b, r.l); return r;
  }
  std::array<uint8_t, 32> to_bytes() const { std::array<uint8_t, 32> o; detail::limbs_to_bytes(l, o.data()); return o; }
  friend Scalar operator+(const Scalar& a, const Scalar& b) { Scalar r; Gpu::instance().check(zc_scalar_add_batch(Gpu::instance().ctx(), a.l, b.l, r.l, 1)); return r; }
  friend Scalar operator-(const Scalar& a, const Scalar& b) { Scalar r; Gpu::instance().check(zc_scalar_sub_batch(Gpu::instance().ctx(), a.l, b.l, r.l, 1)); return r; }
  friend Scalar operator*(const Scalar& a, const Scalar& b) { Scalar r; Gpu::instance().check(zc_scalar_mul_batch(Gpu::instance().ctx(), a.l, b.l, r.l, 1)); return r; }
  Scalar operator-() const { Scalar r; Gpu::instance().check(zc_scalar_neg_batch(Gpu::instance().ctx(), l, r.l, 1)); return r; }
  Scalar square() const { Scalar r; Gpu::instance().check(zc_scalar_square_batch(Gpu::instance().ctx(), l, r.l, 1)); return r; }
  Scalar pow(const Scalar& e) const { Scalar r; Gpu::instance().check(zc_scalar_pow_batch(Gpu::instance().ctx(), l, e.l, r.l, 1)); return r; }   // scalar.rs:293-322
  Scalar half() const { Scalar r; Gpu::instance().check(zc_scalar_half_batch(Gpu::instance().ctx(), l, r.l, 1)); return r; }                    // scalar.rs:285-291
  // scalar.rs:396-415 (width 2 = compute_NAF :370-390): 256 signed digits, least significant first
  std::array<int8_t, 256> compute_window_NAF(uint8_t width) const {
    alignas(4) std::array<int8_t, 256> d;
    Gpu::instance().check(zc_scalar_window_naf_batch(Gpu::instance().ctx(), l, width, d.data(), 1));
    return d;
  }
  std::array<int8_t, 256> compute_NAF() const { return compute_window_NAF(2); }
  // scalar.rs:352-366: the 256 bits of the value, least significant first
  std::array<uint8_t, 256> into_bits() const {
    alignas(16) std::array<uint8_t, 256> b;
    Gpu::instance().check(zc_scalar_into_bits_batch(Gpu::instance().ctx(), l, b.data(), 1));
    return b;
  }
  friend bool operator==(const Scalar& a, const Scalar& b) { return a.to_bytes() == b.to_bytes(); }
  friend bool operator!=(const Scalar& a, const Scalar& b) { return !(a == b); }
};
static_assert(sizeof(Scalar) == 40, "Scalar must be [u64;5]");

// ---- EdwardsPoint ----------------------------------------------------------------------------------------------
struct EdwardsPoint {
  FieldElement X, Y, Z, T;
  const uint64_t* limbs() const { return X.l; }
  uint64_t* limbs() { return X.l; }
  static EdwardsPoint identity() { return {FieldElement::zero(), FieldElement::one(), FieldElement::one(), FieldElement::zero()}; }   // edwards.rs:381-391
  static EdwardsPoint basepoint() {                                      // constants.rs:188-211
    return {{{276718085098056ull, 1646536057461434ull, 2704687245600312ull, 2630386667454967ull, 13476148227069ull}},
            {{1303868825475266ull, 3250718520537114ull, 2702159777242978ull, 2702159776422297ull, 10555311626649ull}},
            {{1, 0, 0, 0, 0}},
            {{3634527586288175ull, 2006028620404053ull, 3424252198034825ull, 2478951925947079ull, 4567251727358ull}}};
  }
  friend EdwardsPoint operator+(const EdwardsPoint& a, const EdwardsPoint& b) { EdwardsPoint r; Gpu::instance().check(zc_point_add_batch(Gpu::instance().ctx(), a.limbs(), b.limbs(), r.limbs(), 1)); return r; }
  friend EdwardsPoint operator-(const EdwardsPoint& a, const EdwardsPoint& b) { EdwardsPoint r; Gpu::instance().check(zc_point_sub_batch(Gpu::instance().ctx(), a.limbs(), b.limbs(), r.limbs(), 1)); return r; }
  EdwardsPoint operator-() const { EdwardsPoint r; Gpu::instance().check(zc_point_neg_batch(Gpu::instance().ctx(), limbs(), r.limbs(), 1)); return r; }
  EdwardsPoint double_() const { EdwardsPoint r; Gpu::instance().check(zc_point_double_batch(Gpu::instance().ctx(), limbs(), r.limbs(), 1)); return r; }   // = self + self, edwards.rs:589-591
  // Mul<&Scalar>: double_and_add, limb-exact (edwards.rs:102-120, 547-577)
  friend EdwardsPoint operator*(const EdwardsPoint& p, const Scalar& s) {
    EdwardsPoint r; Gpu::instance().check(zc_point_scalar_mul_batch(Gpu::instance().ctx(), p.limbs(), s.l, r.limbs(), 1, ZC_SCALAR_MUL_STRICT)); return r;
  }
  friend EdwardsPoint operator*(const Scalar& s, const EdwardsPoint& p) { return p * s; }
  bool is_valid() const { uint8_t ok = 0; Gpu::instance().check(zc_point_is_valid_batch(Gpu::instance().ctx(), limbs(), &ok, 1)); return ok != 0; }   // edwards.rs:393-400
  // AffinePoint::from(EdwardsPoint) (edwards.rs:1085-1092)
  std::pair<FieldElement, FieldElement> to_affine() const {
    uint64_t xy[10]; Gpu::instance().check(zc_point_to_affine_batch(Gpu::instance().ctx(), limbs(), xy, 1));
    std::pair<FieldElement, FieldElement> r; std::memcpy(r.first.l, xy, 40); std::memcpy(r.second.l, xy + 5, 40); return r;
  }
  // PartialEq: equality of the affine points (edwards.rs:360-364)
  friend bool operator==(const EdwardsPoint& a, const EdwardsPoint& b) { return a.to_affine() == b.to_affine(); }
  friend bool operator!=(const EdwardsPoint& a, const EdwardsPoint& b) { return !(a == b); }
};
static_assert(sizeof(EdwardsPoint) == 160, "EdwardsPoint must be X|Y|Z|T of [u64;5]");

// ---- RistrettoPoint --------------------------------------------------------------------------------------------
struct RistrettoPoint {
  EdwardsPoint p;                                                        // pub struct RistrettoPoint(pub EdwardsPoint)
  static RistrettoPoint identity() { return {EdwardsPoint::identity()}; }   // ristretto.rs:186-193
  static RistrettoPoint basepoint() { return {EdwardsPoint::basepoint()}; } // constants.rs:214
  friend RistrettoPoint operator+(const RistrettoPoint& a, const RistrettoPoint& b) { return {a.p + b.p}; }
  friend RistrettoPoint operator-(const RistrettoPoint& a, const RistrettoPoint& b) { return {a.p - b.p}; }
  RistrettoPoint operator-() const { return {-p}; }
  RistrettoPoint double_() const { return {p.double_()}; }
  friend RistrettoPoint operator*(const RistrettoPoint& a, const Scalar& s) { return {a.p * s}; }
  friend RistrettoPoint operator*(const Scalar& s, const RistrettoPoint& a) { return {a.p * s}; }
  // ct_eq: X1 Y2 == Y1 X2  or  X1 X2 == Y1 Y2 (ristretto.rs:166-176)
  friend bool operator==(const RistrettoPoint& a, const RistrettoPoint& b) {
    uint8_t eq = 0; Gpu::instance().check(zc_ristretto_eq_batch(Gpu::instance().ctx(), a.p.limbs(), b.p.limbs(), &eq, 1)); return eq != 0;
  }
  friend bool operator!=(const RistrettoPoint& a, const RistrettoPoint& b) { return !(a == b); }
  std::array<uint8_t, 32> compress() const {                              // ristretto.rs:398-425
    alignas(16) uint8_t out[32]; Gpu::instance().check(zc_ristretto_compress_batch(Gpu::instance().ctx(), p.limbs(), out, 1));
    std::array<uint8_t, 32> r; std::memcpy(r.data(), out, 32); return r;
  }
  static std::optional<RistrettoPoint> decompress(const uint8_t enc[32]) {   // ristretto.rs:96-154
    RistrettoPoint r; uint8_t ok = 0;
    Gpu::instance().check(zc_ristretto_decompress_batch(Gpu::instance().ctx(), enc, r.p.limbs(), &ok, 1));
    if (!ok) return std::nullopt;
    return r;
  }
  static RistrettoPoint from_uniform_bytes(const uint8_t bytes[64]) {     // ristretto.rs:493-507
    RistrettoPoint r; Gpu::instance().check(zc_ristretto_from_uniform_bytes_batch(Gpu::instance().ctx(), bytes, r.p.limbs(), 1)); return r;
  }
};
static_assert(sizeof(RistrettoPoint) == 160, "RistrettoPoint is a newtype over EdwardsPoint");

// ---- slice functions: the batched capability (the reference has none) --------------------------------
namespace batch {
inline void fe_mul(Gpu& g, const FieldElement* a, const FieldElement* b, FieldElement* out, size_t n) { g.check(zc_fe_mul_batch(g.ctx(), a->l, b->l, out->l, n)); }
inline void fe_square(Gpu& g, const FieldElement* a, FieldElement* out, size_t n) { g.check(zc_fe_square_batch(g.ctx(), a->l, out->l, n)); }
inline void fe_mul_square(Gpu& g, const FieldElement* a, const FieldElement* b, FieldElement* prod, FieldElement* sq, size_t n) { g.check(zc_fe_mul_square_batch(g.ctx(), a->l, b->l, prod->l, sq->l, n)); }
inline void point_add(Gpu& g, const EdwardsPoint* p, const EdwardsPoint* q, EdwardsPoint* out, size_t n) { g.check(zc_point_add_batch(g.ctx(), p->limbs(), q->limbs(), out->limbs(), n)); }
inline void point_double(Gpu& g, const EdwardsPoint* p, EdwardsPoint* out, size_t n) { g.check(zc_point_double_batch(g.ctx(), p->limbs(), out->limbs(), n)); }
inline void scalar_mul(Gpu& g, const RistrettoPoint* p, const Scalar* s, RistrettoPoint* out, size_t n, bool strict = true) {
  g.check(zc_point_scalar_mul_batch(g.ctx(), p->p.limbs(), s->l, out->p.limbs(), n, strict ? ZC_SCALAR_MUL_STRICT : ZC_SCALAR_MUL_FAST));
}
inline void basepoint_mul(Gpu& g, const Scalar* s, RistrettoPoint* out, size_t n) { g.check(zc_basepoint_mul_batch(g.ctx(), s->l, out->p.limbs(), n)); }
inline void compress(Gpu& g, const RistrettoPoint* p, uint8_t* out_bytes /* n x 32 */, size_t n) { g.check(zc_ristretto_compress_batch(g.ctx(), p->p.limbs(), out_bytes, n)); }
// sum_i scalars[i] * points[i]: what a caller writes today as zip / map / fold over Mul and Add
inline RistrettoPoint msm(Gpu& g, const RistrettoPoint* points, const Scalar* scalars, size_t n, int window_bits = 16) {
  RistrettoPoint r = RistrettoPoint::identity();
  g.check(zc_msm(g.ctx(), n ? points->p.limbs() : nullptr, n ? scalars->l : nullptr, n, window_bits, r.p.limbs()));
  return r;
}
inline void fe_div(Gpu& g, const FieldElement* a, const FieldElement* b, FieldElement* out, size_t n) { g.check(zc_fe_div_batch(g.ctx(), a->l, b->l, out->l, n)); }
// fixed generators (bulletproofs G_i, H_i): an owning handle over zc_msm_generators.  The handle keeps its own copy of
// everything derived from the points, so the caller's array may be dropped or reused after construction.
class Generators {
 public:
  // points_dev: n points resident on the device (the reference layout); kind ZC_GEN_PREPARED or ZC_GEN_FIXED_BASE
  Generators(Gpu& g, const uint64_t* points_dev, size_t n, int kind = ZC_GEN_PREPARED, int window_bits = 16, int rank = 0, int nranks = 1) : g_(g) {
    g.check(zc_msm_generators_create_dev(g.ctx(), points_dev, n, kind, window_bits, rank, nranks, &h_));
  }
  ~Generators() { if (h_) zc_msm_generators_destroy(g_.ctx(), h_); }
  Generators(const Generators&) = delete;
  Generators& operator=(const Generators&) = delete;
  // sum_i scalars_dev[i] * G_i, result on the device
  void msm_dev(const uint64_t* scalars_dev, uint64_t* out_point_dev, int window_bits = 16) const { g_.check(zc_msm_gen_dev(g_.ctx(), h_, scalars_dev, window_bits, out_point_dev)); }
  void msm_sharded_dev(const uint64_t* scalars_dev, uint64_t* out_point_dev, int window_bits = 16) const { g_.check(zc_msm_gen_sharded_dev(g_.ctx(), h_, scalars_dev, window_bits, out_point_dev)); }
  zc_msm_generators* raw() const { return h_; }
 private:
  Gpu& g_;
  zc_msm_generators* h_ = nullptr;
};
}  // namespace batch

}  // namespace zerocaf
