// The reference's examples/basic_ops.rs (a scalar times a point, printed) with the C++ host mirror over libzerocaf_b200.so,
// followed by the batched form of the same computation and an MSM -- what the reference cannot express.
//   g++ -std=c++17 -I include examples/basic_ops.cpp -L dusk_zerocaf_b200 -lzerocaf_b200 -Wl,-rpath,$PWD/dusk_zerocaf_b200
#include <cstdio>
#include <vector>

#include "zerocaf_b200.hpp"

using namespace zerocaf;

static void print_point(const char* name, const RistrettoPoint& p) {
  auto c = p.compress();
  std::printf("%s = ", name);
  for (uint8_t b : c) std::printf("%02x", b);
  std::printf("\n");
}

int main() {
  // one scalar, one point: the reference's Mul<&Scalar> for &RistrettoPoint (double_and_add)
  const Scalar s = Scalar::from_u64(0x123456789abcdefull);
  const RistrettoPoint B = RistrettoPoint::basepoint();
  const RistrettoPoint P = B * s;
  print_point("[s]B", P);
  std::printf("[s]B + [s]B == [2s]B : %s\n", (P + P == B * (s + s)) ? "true" : "false");

  // the same for 2^16 scalars at once, then sum_i [t_i]P_i as one MSM
  Gpu& gpu = Gpu::instance();
  const size_t n = 1 << 16;
  std::vector<Scalar> t(n);
  for (size_t i = 0; i < n; i++) t[i] = Scalar::from_u64(0x9e3779b97f4a7c15ull * (i + 1) >> 4);
  std::vector<RistrettoPoint> points(n), out(n);
  batch::basepoint_mul(gpu, t.data(), points.data(), n);
  batch::scalar_mul(gpu, points.data(), t.data(), out.data(), n, /*strict=*/false);
  print_point("[t_0]P_0", out[0]);
  print_point("sum_i [t_i]P_i", batch::msm(gpu, points.data(), t.data(), n));
  return 0;
}
