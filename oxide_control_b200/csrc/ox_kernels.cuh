// Device helpers shared by the kernel translation units: TMA staging of the model tables and env indexing.
#pragma once
#include <cuda_runtime.h>

#include "ox_blob.h"

namespace ox {

// ---------------------------------------------------------------- model staging: one TMA bulk copy per CTA
// (cp.async.bulk global -> shared, completion on an mbarrier; shows up as UBLKCP in SASS)
__device__ __forceinline__ const unsigned char* stage_model(const unsigned char* __restrict__ gblob, int bytes) {
  extern __shared__ __align__(128) unsigned char ox_smem[];
  __shared__ __align__(8) unsigned long long mbar;
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ox_smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(gblob),
                 "r"(bytes), "r"(mb)
                 : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mb)
        : "memory");
  }
  return ox_smem;
}

// env owned by this thread, or -1. Warps may be deliberately under-filled (b.lanes < 32 active lanes per warp): at small
// batch sizes the step is latency-bound with far fewer warps than SM sub-partitions (8192 envs = 256 full warps for 592
// schedulers), so spreading the envs over more, thinner warps uses the idle schedulers and shortens every warp's
// divergent max-over-lanes critical path.
template <typename T>
__device__ __forceinline__ int env_index(const DevBatch<T>& b) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (lane >= b.lanes) return -1;
  const int e = warp * b.lanes + lane;
  return e < b.nenv ? e : -1;
}


}  // namespace ox
