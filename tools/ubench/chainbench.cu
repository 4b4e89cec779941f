// chainbench.cu -- round-2 experiment, NOT part of the product: single-warp latency of the operations the MSM's serial
// tail is made of (window chain: dependent point doublings, four lanes per point; reduction trees: four-lane additions).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o chainbench chainbench.cu
//   ./chainbench
// Every probe runs ITERS dependent operations in one warp and reports cycles (clock64) and ns (%globaltimer) per operation.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../dusk_zerocaf_b200/csrc/zc_quad.cuh"

using namespace zc;

#ifndef ITERS
#define ITERS 1024
#endif

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// round-1 versions (nested ?: selects that compile to divergent branches, canonical values, 2d T2 as a full product)
__device__ __forceinline__ Fe quad_stage2_r1(const Fe& E, const Fe& F, const Fe& G, const Fe& H, int q) {
  typedef ModP M;
  // X3 = E F, Y3 = G H, Z3 = F G, T3 = E H
  Fe u, v;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u.w[k] = (q == 0 || q == 3) ? E.w[k] : (q == 1 ? G.w[k] : F.w[k]);
    v.w[k] = (q == 0) ? F.w[k] : (q == 2 ? G.w[k] : H.w[k]);
  }
  return mont_mul<M>(u, v);
}
__device__ __noinline__ Fe quad_double_r1(Fe c, int q, int qbase) {
  typedef ModP M;
  // stage 1: X^2, Y^2, Z^2 on lanes 0..2 and T Z (= X Y, so E = 2 T Z) on lane 3
  Fe z = shfl_fe(c, qbase + 2);
  Fe in2 = c;
  if (q == 3) in2 = z;
  Fe s = mont_mul<M>(c, in2);
  Fe A = shfl_fe(s, qbase), B = shfl_fe(s, qbase + 1), ZZ = shfl_fe(s, qbase + 2), TZ = shfl_fe(s, qbase + 3);
  Fe E = fe_dbl_lazy(TZ);                      // < 2m
  Fe C = fe_dbl_lazy(ZZ);                      // < 2m
  Fe G = fe_sub_lazy<1>(B, A);                 // B - A + m      in (0, 2m)
  Fe F = fe_sub_lazy<2>(G, C);                 // G - C + 2m     in (0, 4m)
  Fe ApB = fe_add<M>(A, B);                    // canonical
  Fe zero{{0, 0, 0, 0, 0, 0, 0, 0}};
  Fe H = fe_sub_lazy<1>(zero, ApB);            // m - (A + B)    in (0, m]
  return quad_stage2_r1(E, F, G, H, q);
}
// c (distributed over the quad) += the full point p (every lane holds all of p; both canonical).  All linear combinations
// are lazy (< 2m, see above) and computed by every lane before the select: a quad's lanes never branch apart except for
// lane 3's 2d T2 product.
__device__ __noinline__ Fe quad_add_r1(Fe c, Pt p, int q, int qbase) {
  typedef ModP M;
  const Fe x1 = shfl_fe(c, qbase), y1 = shfl_fe(c, qbase + 1);
  const Fe d1 = fe_sub_lazy<1>(y1, x1), s1 = fe_add_lazy(y1, x1);
  const Fe d2 = fe_sub_lazy<1>(p.Y, p.X), s2 = fe_add_lazy(p.Y, p.X);
  const Fe z2 = fe_dbl_lazy(p.Z);
  Fe t2 = p.T;
  if (q == 3) t2 = mont_mul<M>(p.T, D2_MONT());
  Fe u, v;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u.w[k] = q == 0 ? d1.w[k] : (q == 1 ? s1.w[k] : c.w[k]);
    v.w[k] = q == 0 ? d2.w[k] : (q == 1 ? s2.w[k] : (q == 2 ? z2.w[k] : t2.w[k]));
  }
  const Fe s = mont_mul<M>(u, v);                              // A, B, D = 2 Z1 Z2, C = T1 2d T2
  const Fe A = shfl_fe(s, qbase), B = shfl_fe(s, qbase + 1), D = shfl_fe(s, qbase + 2), C = shfl_fe(s, qbase + 3);
  const Fe E = fe_sub_lazy<1>(B, A);
  const Fe F = fe_sub_lazy<1>(D, C);
  const Fe G = fe_add_lazy(D, C);
  const Fe H = fe_add_lazy(B, A);
  return quad_stage2_r1(E, F, G, H, q);
}


__device__ __forceinline__ Fe quad_select_uv_xor(const Fe& E, const Fe& F, const Fe& G, const Fe& H, int q) {
  Fe r;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const uint32_t u = (q == 0 || q == 3) ? E.w[k] : (q == 1 ? G.w[k] : F.w[k]);
    const uint32_t v = (q == 0) ? F.w[k] : (q == 2 ? G.w[k] : H.w[k]);
    r.w[k] = (u ^ v) & (k == 7 ? 0x0fffffffu : 0xffffffffu);
  }
  return r;
}

struct Res { long long cycles; unsigned long long ns; uint32_t words[8]; };

template <int P>
__global__ void __launch_bounds__(32) probe(const uint32_t* __restrict__ in, Res* __restrict__ res) {
  __shared__ __align__(16) uint32_t sp[32];
  const int lane = threadIdx.x, q = lane & 3, qbase = lane & ~3;
  Fe x, y;
#pragma unroll
  for (int k = 0; k < 8; k++) { x.w[k] = in[16 * lane + k]; y.w[k] = in[16 * lane + 8 + k]; }
  x.w[7] &= 0x0fffffffu; y.w[7] &= 0x0fffffffu;
  if (lane < 32) sp[lane] = in[lane] & 0x0fffffffu;
  __syncwarp();
  Pt pp = ld_pt(sp);
  const unsigned long long t0 = gtimer();
  const long long c0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
    if (P == 0) x = mont_mul<ModP>(x, y);
    else if (P == 1) x = mont_mul_lazy<ModP>(x, y);
    else if (P == 2) x = quad_double_r1(x, q, qbase);
    else if (P == 3) x = quad_add_r1(x, pp, q, qbase);
    else if (P == 4) x = quad_double_inl(x, q, qbase);
    else if (P == 5) x = quad_add_inl(x, pp, q, qbase);
    else if (P == 6) {            // the shuffles of one doubling alone (z to lane 3, then the four stage-1 results to all)
      Fe z = shfl_fe(x, qbase + 2);
      Fe a = shfl_fe(z, qbase), b = shfl_fe(z, qbase + 1), c = shfl_fe(z, qbase + 2), d = shfl_fe(z, qbase + 3);
#pragma unroll
      for (int k = 0; k < 8; k++) x.w[k] = a.w[k] ^ b.w[k] ^ c.w[k] ^ (d.w[k] + 1u);
    } else if (P == 7) {          // the lazy linear combinations of one doubling alone
      Fe E = fe_dbl_lazy(x), C = fe_dbl_lazy(y), G = fe_sub_lazy<1>(x, y), F = fe_sub_lazy<2>(G, C);
      Fe ApB = fe_add<ModP>(x, y);
      Fe zero{{0, 0, 0, 0, 0, 0, 0, 0}};
      Fe H = fe_sub_lazy<1>(zero, ApB);
      x = quad_select_uv_xor(E, F, G, H, q);
    } else if (P == 8) x = quad_double(x, q, qbase);
    else if (P == 9) x = quad_add(x, pp, q, qbase);
    else if (P == 10) x = fe_mul_small(x, 126297u);
    else if (P == 11) {          // the lazy linear combinations + branch-free selection of one doubling
      Fe E = fe_dbl_lazy(x), C = fe_dbl_lazy(y), G = fe_sub_lazy<2>(x, y), F = fe_sub_lazy<3>(G, C);
      Fe zero{{0, 0, 0, 0, 0, 0, 0, 0}};
      Fe H = fe_sub_lazy<3>(zero, fe_add_lazy(x, y));
      Fe u, v; quad_pick(u, v, E, F, G, H, q);
#pragma unroll
      for (int k = 0; k < 8; k++) x.w[k] = (u.w[k] ^ v.w[k]) & (k == 7 ? 0x0fffffffu : 0xffffffffu);
    }
  }
  const long long c1 = clock64();
  const unsigned long long t1 = gtimer();
  if (lane == 0) { res->cycles = c1 - c0; res->ns = t1 - t0; }
  if (lane < 4) {
    // fold the quad's coordinates into a checksum the host can compare between variants (as a projective point the lean
    // variants return other representatives; the host only prints the words)
#pragma unroll
    for (int k = 0; k < 8; k++) if (lane == 0) res->words[k] = x.w[k];
  }
}

// affine check of a doubling variant against quad_double: X/Z and Y/Z after N doublings must agree.  Done on the device
// with cross products: X1 Z2 == X2 Z1 and Y1 Z2 == Y2 Z1 (canonical after mont_mul).
template <int V>
__global__ void __launch_bounds__(32) check(const uint32_t* __restrict__ in, int n, int* __restrict__ bad) {
  const int lane = threadIdx.x, q = lane & 3, qbase = lane & ~3;
  // start from a valid curve point: the caller passes the basepoint in Montgomery form in in[0..31]
  Fe c0; ld_fe(in + 8 * q, c0);
  Fe a = c0, b = c0;
  Pt base; base = ld_pt(in);
  for (int i = 0; i < n; i++) {
    a = quad_double_r1(a, q, qbase);
    b = V == 0 ? quad_double_inl(b, q, qbase) : b;
    if ((i & 3) == 3) { a = quad_add_r1(a, base, q, qbase); b = quad_add_inl(b, base, q, qbase); }
  }
  Pt A, B;
  A.X = shfl_fe(a, qbase); A.Y = shfl_fe(a, qbase + 1); A.Z = shfl_fe(a, qbase + 2); A.T = shfl_fe(a, qbase + 3);
  B.X = shfl_fe(b, qbase); B.Y = shfl_fe(b, qbase + 1); B.Z = shfl_fe(b, qbase + 2); B.T = shfl_fe(b, qbase + 3);
  Fe l1 = mont_mul<ModP>(A.X, B.Z), r1 = mont_mul<ModP>(B.X, A.Z);
  Fe l2 = mont_mul<ModP>(A.Y, B.Z), r2 = mont_mul<ModP>(B.Y, A.Z);
  Fe l3 = mont_mul<ModP>(A.T, B.Z), r3 = mont_mul<ModP>(B.T, A.Z);
  // the lean results are lazily reduced: mont_mul of values < 8m by canonical ones is < 2m before its reduce_once, so
  // one more reduce_once after it makes both sides canonical
  reduce_once<ModP>(l1); reduce_once<ModP>(r1); reduce_once<ModP>(l2); reduce_once<ModP>(r2); reduce_once<ModP>(l3); reduce_once<ModP>(r3);
  if (lane == 0) *bad = (fe_eq(l1, r1) ? 0 : 1) + (fe_eq(l2, r2) ? 0 : 2) + (fe_eq(l3, r3) ? 0 : 4);
}

int main() {
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  uint32_t h[512];
  for (int k = 0; k < 512; k++) h[k] = 0x9e3779b9u * (k + 1) ^ (k << 7);
  // basepoint (x, 3/5, 1, x*3/5) in Montgomery form R = 2^256 for the correctness check: computed on the host with
  // __int128-free schoolbook bigints would be long; instead take the identity-safe route: P = (0, 1, 1, 0) doubled stays
  // the identity, which would not exercise anything -- so the check kernel gets the words of a valid point from argv
  // (tools/ubench/chainbench_point.txt, written by the python driver) or falls back to the probes only.
  uint32_t pt[32]; bool have_pt = false;
  if (FILE* f = fopen("chainbench_point.txt", "r")) {
    have_pt = true;
    for (int k = 0; k < 32; k++) if (fscanf(f, "%x", &pt[k]) != 1) have_pt = false;
    fclose(f);
  }
  uint32_t *in, *pin; Res* res; int* bad;
  cudaMalloc(&in, sizeof(h)); cudaMalloc(&pin, sizeof(pt)); cudaMalloc(&res, sizeof(Res)); cudaMalloc(&bad, 4);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  if (have_pt) {
    cudaMemcpy(pin, pt, sizeof(pt), cudaMemcpyHostToDevice);
    int hb = -1;
    check<0><<<1, 32>>>(pin, 200, bad);
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
    printf("lean doubling/addition vs quad_double/quad_add after 200 doublings + 50 additions: %s (code %d)\n", hb == 0 ? "same point" : "MISMATCH", hb);
  }
  const char* names[] = {"mont_mul (dependent chain)", "mont_mul_lazy", "quad_double (round 1)", "quad_add (round 1)",
                         "quad_double_inl", "quad_add_inl", "shuffles of one doubling", "linear ops of one doubling (r1)",
                         "quad_double (noinline)", "quad_add (noinline)", "fe_mul_small", "linear ops + selp picks"};
  for (int p = 0; p < 12; p++) {
    Res r;
    for (int rep = 0; rep < 3; rep++) {
      switch (p) {
        case 0: probe<0><<<1, 32>>>(in, res); break; case 1: probe<1><<<1, 32>>>(in, res); break;
        case 2: probe<2><<<1, 32>>>(in, res); break; case 3: probe<3><<<1, 32>>>(in, res); break;
        case 4: probe<4><<<1, 32>>>(in, res); break; case 5: probe<5><<<1, 32>>>(in, res); break;
        case 6: probe<6><<<1, 32>>>(in, res); break; case 7: probe<7><<<1, 32>>>(in, res); break;
        case 8: probe<8><<<1, 32>>>(in, res); break; case 9: probe<9><<<1, 32>>>(in, res); break;
        case 10: probe<10><<<1, 32>>>(in, res); break; case 11: probe<11><<<1, 32>>>(in, res); break;
      }
      cudaMemcpy(&r, res, sizeof(r), cudaMemcpyDeviceToHost);
    }
    printf("%-32s %7.1f cycles  %7.1f ns per op  (SM clock %.0f MHz)\n", names[p], (double)r.cycles / ITERS, (double)r.ns / ITERS,
           (double)r.cycles / (double)r.ns * 1e3);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
