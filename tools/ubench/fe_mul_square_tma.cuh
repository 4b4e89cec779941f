// Experiment record (not part of libzerocaf_b200.so): TMA-staged persistent variant of fe_mul_square_kernel.
// Measured on B200 (round 1): 0.544 ms per 2^24 pairs vs 0.532 ms for the direct kernel -- the kernel is bound by the
// integer-multiply pipe and the 1000 W power cap (sm clocks 1.78-1.88 GHz under this load), not by its 40-byte-stride
// access pattern, so the product keeps the simpler direct kernel.  To try it again: include this file in zc_kernels.cu
// after fe_mul_square_kernel and launch it for the 16-byte-aligned, non-aliased full tiles.
// ---- K1 fused, TMA-staged: the same arithmetic, operands streamed through shared memory by the bulk-copy engine ------
// The AoS [u64;5] layout makes every per-thread limb load a 40-byte-stride access: 32 sectors per warp request and each
// sector requested five times.  Here a CTA owns tiles of TILE consecutive elements; one thread issues two
// cp.async.bulk loads (5 KiB each, fully contiguous) per tile into a three-stage ring completed on mbarriers, threads
// read / write their limbs in shared memory (conflict-free at 40-byte stride), and the results leave with two
// cp.async.bulk stores from the same stage buffer.  The grid is persistent (a few CTAs per SM), so the load of tile
// k+1 is in flight while tile k is multiplied.
constexpr int TMA_TILE = 128;
constexpr int TMA_STAGES = 3;
constexpr uint32_t TMA_ARR_BYTES = TMA_TILE * 40;          // one operand array of one tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <class M>
__global__ void __launch_bounds__(TMA_TILE, 7) fe_mul_square_tma_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b,
                                                                        uint64_t* __restrict__ prod, uint64_t* __restrict__ sq,
                                                                        uint32_t ntiles) {
  __shared__ __align__(128) uint64_t stage[TMA_STAGES][2][TMA_TILE * 5];   // [stage][a|b -> prod|sq][limbs]
  __shared__ __align__(8) uint64_t mbar[TMA_STAGES];
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < TMA_STAGES; k++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[k])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue_load = [&](uint32_t tile, int s) {
    const uint32_t mb = smem_u32(&mbar[s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(2u * TMA_ARR_BYTES) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(&stage[s][0][0])), "l"(a + (size_t)tile * TMA_TILE * 5), "r"(TMA_ARR_BYTES), "r"(mb) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(&stage[s][1][0])), "l"(b + (size_t)tile * TMA_TILE * 5), "r"(TMA_ARR_BYTES), "r"(mb) : "memory");
  };
  uint32_t it = 0;
  int s = 0;                                 // stage of the current tile; its use count is it / TMA_STAGES
  if (tid == 0 && blockIdx.x < ntiles) issue_load(blockIdx.x, 0);
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++, s = (s + 1 == TMA_STAGES ? 0 : s + 1)) {
    const uint32_t parity = (it / TMA_STAGES) & 1u;
    const uint32_t next = tile + gridDim.x;
    if (tid == 0 && next < ntiles) {
      // the next stage was the source of the bulk stores issued two iterations ago: all but the latest group must have
      // finished reading shared memory
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      issue_load(next, s + 1 == TMA_STAGES ? 0 : s + 1);
    }
    {  // wait for this tile's bytes
      const uint32_t mb = smem_u32(&mbar[s]);
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mb), "r"(parity) : "memory");
      }
    }
    const uint64_t* sa = &stage[s][0][5 * tid];
    const uint64_t* sb = &stage[s][1][5 * tid];
    Fe x = fe_from_limbs52(sa[0], sa[1], sa[2], sa[3], sa[4]);
    Fe y = fe_from_limbs52(sb[0], sb[1], sb[2], sb[3], sb[4]);
    __syncthreads();                       // all inputs are in registers: the stage can take the results
    uint64_t l[5];
    fe_to_limbs52(fe_mul_normal<M>(x, y), l);
    uint64_t* sp = &stage[s][0][5 * tid];
    sp[0] = l[0]; sp[1] = l[1]; sp[2] = l[2]; sp[3] = l[3]; sp[4] = l[4];
    fe_to_limbs52(fe_sqr_normal<M>(x), l);
    uint64_t* sq_s = &stage[s][1][5 * tid];
    sq_s[0] = l[0]; sq_s[1] = l[1]; sq_s[2] = l[2]; sq_s[3] = l[3]; sq_s[4] = l[4];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the bulk-copy engine
    __syncthreads();
    if (tid == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(prod + (size_t)tile * TMA_TILE * 5), "r"(smem_u32(&stage[s][0][0])), "r"(TMA_ARR_BYTES) : "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(sq + (size_t)tile * TMA_TILE * 5), "r"(smem_u32(&stage[s][1][0])), "r"(TMA_ARR_BYTES) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the last stores
}

