// loss_sparse.cu -- K1 "confidence-first" form: read only the sectors a cell needs (sm_100a).
//
// With ~3 objects per 196 cells, 98.5 % of the cells of a call hold no object, and such a cell needs 12 bytes of
// input: target[0] (the object test, v1Loss.py:28) and pred[0:2] (the two confidences of the no-object term,
// v1Loss.py:91).  The dense streaming kernel (loss_nhwc.cu) nevertheless moves all 240 input bytes of every cell:
// it sits at 1.00 of the measured HBM rate for 360 B / cell and has no headroom left.
// This kernel changes the access pattern instead: one thread per cell pulls the first 8 bytes of its target row and
// of its pred row with two independent 8-byte loads (one 32-byte sector each), and only an object cell goes back
// for the rest of its two rows.  The gradient tile is still assembled in shared memory and leaves with one bulk
// (TMA) store, so the 120 B / cell of output stay fully coalesced.  DRAM then moves
//     32 + 32 + 120 = 184 B / cell  if the memory system fetches single sectors,
//     64 + 64 + 120 = 248 B / cell  at 64-byte fetch granularity (a 120-byte row stride touches ~every 64-byte block),
// against 360 dense; which of the two it is, is what tools/sparse_probe.py and the ncu capture measure
// (cudaLimitMaxL2FetchGranularity is the caller's knob).
//
// Latency, not bandwidth, is what a CTA sees here (two dependent-free sector loads, ~1 us), so the kernel relies on
// residency (8 CTAs x 128 threads per SM: 2 048 sector requests in flight per SM) plus a one-tile register
// prefetch: the loads of tile k+1 are issued before tile k is evaluated.
//
// LIST form (yolo1_loss_fwd_bwd_objects): the 4-byte ownership map replaces the target sector (coalesced: 128 B
// per warp), so a cell costs 4 + 32 + 120 B.
#include "loss_common.cuh"

namespace yolo1 {
namespace {

// a cell whose first channel pair is already in registers; the rest of the row is fetched on demand
struct SparseIn {
  float2 v01;
  const float* row;
  __device__ __forceinline__ float2 ld2(int c) const {
    return c == 0 ? v01 : __ldg(reinterpret_cast<const float2*>(row + c));
  }
};

__device__ __forceinline__ float2 ld_sector8(const float* p, uint64_t pol) {
  // streamed once: keep it out of L1, first in line for eviction in L2
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;"
               : "=f"(v.x), "=f"(v.y)
               : "l"(p), "l"(pol));
  return v;
}

template <bool HAS_GRAD, int TILE, bool LIST>
__global__ void __launch_bounds__(TILE) loss_sparse_kernel(const __grid_constant__ LossParams p) {
  constexpr int D = 30, NOUT = 2;
  constexpr uint32_t GB = TILE * D * sizeof(float);
  extern __shared__ __align__(128) unsigned char smem[];
  float* so = reinterpret_cast<float*>(smem);
  const int tid = threadIdx.x;
  const int64_t full = p.cells / TILE;
  const int64_t my_n = full > (int64_t)blockIdx.x ? (full - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const float* gp = reinterpret_cast<const float*>(p.pred);
  float* gg = reinterpret_cast<float*>(p.grad);
  const uint64_t pol = policy_evict_first();
  CellSums sums = {0.f, 0.f, 0.f, 0.f};
  uint32_t m1 = 0, m2 = 0;

  // register prefetch of the next tile's two sectors (LIST: the ownership slot, coalesced, and the object record)
  float2 nt = make_float2(0.f, 0.f), nc = make_float2(0.f, 0.f);
  ObjFetch nobj = {make_float4(0.f, 0.f, 0.f, 0.f), 0, -1};
  auto fetch = [&](int64_t k) {
    const int64_t q = ((int64_t)blockIdx.x + k * gridDim.x) * TILE + tid;
    nc = ld_sector8(gp + q * D, pol);
    if (LIST)
      nobj = fetch_object(p, __ldg(p.cellobj + q));
    else
      nt = ld_sector8(p.target + q * D, pol);
  };
  if (my_n > 0) fetch(0);
  for (int64_t k = 0; k < my_n; ++k) {
    const int64_t tile = (int64_t)blockIdx.x + k * gridDim.x, q = tile * TILE + tid;
    const int o = (int)(k % NOUT);
    const SparseIn P{nc, gp + q * D};
    const SparseIn T{nt, p.target + q * D};
    const ListTarget2 TL = LIST ? list_target2(p, nobj) : ListTarget2{0.f, 0.f, 0.f, 0.f, -1, false};
    if (k + 1 < my_n) fetch(k + 1);      // in flight while this tile is evaluated and stored
    const SmemOutF32 G{so + o * (TILE * D) + tid * D};
    bool obj;
    if (LIST)
      obj = cell_b2c20<HAS_GRAD>(P, TL, G, p, sums);
    else
      obj = cell_b2c20<HAS_GRAD>(P, T, G, p, sums);
    if (obj) note_object(m1, m2, q);
    if (HAS_GRAD) {
      fence_async_smem();
      if (tid == 0) bulk_wait_read<NOUT - 2>();   // buffer (k+1) % NOUT is free again
      __syncthreads();
      if (tid == 0) {
        bulk_s2g(gg + tile * (TILE * D), so + o * (TILE * D), GB, pol);
        bulk_commit();
      }
    }
  }
  // ragged tail (< TILE cells): one CTA, straight from / to global memory
  const int64_t tail0 = full * TILE;
  if ((int64_t)blockIdx.x == full % gridDim.x && tail0 + tid < p.cells) {
    const int64_t q = tail0 + tid;
    const GlobIn<float> P{gp + q * D, 1, false};
    const GlobOut<float> G{HAS_GRAD ? gg + q * D : nullptr, 1, nullptr, 0, false};
    bool obj;
    if (LIST) {
      obj = cell_generic<HAS_GRAD, false>(P, list_targetS(p, q), G, p, sums);
    } else {
      const GlobIn<float> T{p.target + q * D, 1, false};
      obj = cell_generic<HAS_GRAD, false>(P, T, G, p, sums);
    }
    if (obj) note_object(m1, m2, q);
  }
  block_epilogue<float, HAS_GRAD, true>(sums, m1, m2, p);
}

template <bool HAS_GRAD, int TILE, bool LIST>
int launch_sparse_t(const LossParams& p, cudaStream_t stream) {
  constexpr size_t smem = (size_t)2 * TILE * 30 * sizeof(float);
  auto kern = loss_sparse_kernel<HAS_GRAD, TILE, LIST>;
  static KernelPrep prep;
  int sms = kNumSMs, per_sm = 1;
  if (int rc = prepare_kernel(prep, kern, TILE, smem, true, &sms, &per_sm)) return rc;
  const int64_t tiles = p.cells / TILE;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > tiles) grid = tiles;
  if (grid > kMaxGrid) grid = kMaxGrid;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, TILE, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace

// shape: 0 = 128 cells per tile (30 KB of gradient buffers per CTA: 7 CTAs per SM), 1 = 64 (15 KB: 14 CTAs per SM),
// 2 = 256
int launch_loss_sparse(const LossParams& p, bool has_grad, int shape, cudaStream_t stream) {
  const bool list = p.list_mode != 0;
#define YOLO1_SPARSE(T)                                                                           \
  (list ? (has_grad ? launch_sparse_t<true, T, true>(p, stream) : launch_sparse_t<false, T, true>(p, stream)) \
        : (has_grad ? launch_sparse_t<true, T, false>(p, stream) : launch_sparse_t<false, T, false>(p, stream)))
  switch (shape) {
    case 0: return YOLO1_SPARSE(128);
    case 1: return YOLO1_SPARSE(64);
    case 2: return YOLO1_SPARSE(256);
    default: return YOLO1_ERR_ARG;
  }
#undef YOLO1_SPARSE
}

}  // namespace yolo1
