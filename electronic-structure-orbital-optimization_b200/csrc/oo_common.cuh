// Shared device helpers for the orbital-optimisation kernels (sm_100a only).
// PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), FP64 tensor-core MMA (DMMA.8x8x4).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace oo {

// ----------------------------------------------------------------------------------------------
// FP64 tensor-core MMA.  D(8x8) += A(8x4) * B(4x8).
// Fragment ownership (PTX ISA, mma.m8n8k4 .f64), lane = 4*g + c:
//   A[g][c]            one register
//   B[c][g]            one register
//   C/D[g][2c], [g][2c+1]  two registers
// On sm_100a every f64 mma shape lowers to SASS DMMA.8x8x4, so the native granule is used directly.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ----------------------------------------------------------------------------------------------
// TMA: 3-D tiled tensor load global -> shared, completion signalled on an mbarrier.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* tmap, int c0,
                                            int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst_smem), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// Same load with an L2 eviction-priority hint (policy from l2_policy_*()).
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst_smem, const CUtensorMap* tmap, int c0,
                                                 int c1, int c2, uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%2, %3, %4}], [%5], %6;"
      ::"r"(dst_smem), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy)
      : "memory");
}
// L2 eviction policies: data that is read exactly once (the ERI stream) should leave L2 first,
// small results that the next kernel re-reads should stay.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void st_global_hint(double* addr, double v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(addr), "d"(v), "l"(policy)
               : "memory");
}
// 16-byte read-only load with an L2 eviction-priority hint.
__device__ __forceinline__ double2 ldg128_hint(const double* addr, uint64_t policy) {
  double2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
               : "=d"(v.x), "=d"(v.y)
               : "l"(addr), "l"(policy));
  return v;
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// Programmatic dependent launch (PDL).  A kernel launched with the programmatic-stream-
// serialization attribute may start while its predecessor in the stream is still running:
// pdl_launch_dependents() (executed by every CTA of the predecessor, as early as possible) lets
// the successor's CTAs be scheduled as SMs drain; pdl_wait() in the successor blocks until the
// predecessor grid has completed and its memory is visible.  Everything a kernel reads from its
// predecessor -- including the optimiser's stop flag -- must come after pdl_wait().  Both are
// no-ops when the kernel was launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Nanosecond wall clock of the device (independent of the SM clock).
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Named barrier for a subset of the CTA (id 1..15; id 0 is __syncthreads).
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Non-blocking arrival on a named barrier (producer side of an arrive/sync pair).
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Shared-memory vector loads with explicit 32-bit shared addresses.
__device__ __forceinline__ void lds128(double& x, double& y, uint32_t addr) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double x;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(x) : "r"(addr));
  return x;
}

// Deterministic block-wide sum (fixed shuffle tree in both stages).  blockDim.x must be a multiple
// of 32, <= 1024.  `scratch` needs 33 doubles of shared memory.  Result is valid in every thread.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    double t = lane < nw ? scratch[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

}  // namespace oo
