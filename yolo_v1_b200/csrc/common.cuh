// common.cuh -- shared helpers of libyolo1_b200.so (sm_100a only).
//
// PTX wrappers for the bulk asynchronous copy engine (TMA, `cp.async.bulk`, SASS UBLKCP), mbarriers and
// the proxy fences that order generic shared-memory writes against it, plus small warp utilities.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "../../include/yolo1_b200.h"

namespace yolo1 {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

#define YOLO1_CUDA_TRY(expr)                      \
  do {                                            \
    cudaError_t _e = (expr);                      \
    if (_e != cudaSuccess) return (int)_e;        \
  } while (0)

// ---- launch hygiene: per (kernel, device) one-time work ---------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), cudaDeviceGetAttribute(MultiProcessorCount) and
// cudaOccupancyMaxActiveBlocksPerMultiprocessor are driver round trips of a few microseconds each; a small call
// (train.py:38-41: 2 352 cells) is shorter than the three together.  Each launcher owns one `static KernelPrep`
// (so: one per kernel instantiation) and asks it; the driver is consulted only when the device, the block size
// or the shared-memory size differ from what was seen last.
constexpr int kMaxDevices = 64;
struct KernelPrep {
  std::mutex mu;
  size_t attr_smem[kMaxDevices] = {};   // largest dynamic shared-memory size the kernel was opted in for
  int sms[kMaxDevices] = {};
  int occ_threads[kMaxDevices] = {};
  size_t occ_smem[kMaxDevices] = {};
  int per_sm[kMaxDevices] = {};
  bool occ_valid[kMaxDevices] = {};
};
// Returns 0 and fills sms / per_sm (resident CTAs per SM, >= 1; pass want_occupancy = false to skip the query).
template <typename K>
int prepare_kernel(KernelPrep& c, K kern, int threads, size_t smem, bool want_occupancy, int* sms, int* per_sm) {
  int dev = 0;
  YOLO1_CUDA_TRY(cudaGetDevice(&dev));
  const int d = (dev >= 0 && dev < kMaxDevices) ? dev : -1;
  std::lock_guard<std::mutex> lock(c.mu);
  if (d < 0 || smem > c.attr_smem[d]) {
    if (smem > 48 * 1024 || d < 0)
      YOLO1_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (d >= 0) c.attr_smem[d] = smem;
  }
  int n_sm = d >= 0 ? c.sms[d] : 0;
  if (n_sm == 0) {
    YOLO1_CUDA_TRY(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    if (d >= 0) c.sms[d] = n_sm;
  }
  if (sms) *sms = n_sm;
  if (want_occupancy) {
    int occ = 1;
    if (d >= 0 && c.occ_valid[d] && c.occ_threads[d] == threads && c.occ_smem[d] == smem) {
      occ = c.per_sm[d];
    } else {
      YOLO1_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
      if (occ < 1) occ = 1;
      if (d >= 0) c.occ_valid[d] = true, c.occ_threads[d] = threads, c.occ_smem[d] = smem, c.per_sm[d] = occ;
    }
    if (per_sm) *per_sm = occ;
  }
  return 0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- bulk async copies (1-D TMA) -------------------------------------------------------------------
// global -> shared, completion signalled on an mbarrier as transaction bytes. 16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// shared -> global, tracked by the per-thread bulk async-group.
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (before a bulk shared->global copy)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---- warp helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// The reference encoder's arithmetic for one box centre (utils/YOLODataLoader.py:218-224), every operation
// rounded on its own (explicit intrinsics: valid in translation units built with or without -fmad=false):
//   idx = ceil(c / fl32(1/S)) - 1 ;  delta = (c - idx * fl32(1/S)) / fl32(1/S)
__device__ __forceinline__ void encode_axis(float c, float cs, float& fidx, float& delta) {
  fidx = __fsub_rn(ceilf(__fdiv_rn(c, cs)), 1.0f);
  delta = __fdiv_rn(__fsub_rn(c, __fmul_rn(fidx, cs)), cs);
}

// element load/store with on-the-fly bf16 <-> fp32 conversion
__device__ __forceinline__ float ld_elem(const float* p) { return *p; }
__device__ __forceinline__ float ld_elem(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_elem(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_elem(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

}  // namespace yolo1
