// loss_nhwc.cu -- K1 streaming kernel for contiguous [N,S,S,30] tensors (see loss_common.cuh for the design).
#include "loss_common.cuh"

namespace yolo1 {
namespace {

// ---- LIST 2: the owners of a tile's cells straight from the object lists (warp 0 of loss_tma_kernel) -----------
struct OwnStage {
  uint32_t q0;        // first cell of the tile (a launch holds fewer than 2^32 cells)
  int n_lo, span;     // first image the tile touches, how many it touches
  int off;            // stage A: lane j holds offsets[n_lo + j], j <= span
  int k, img, label;  // stage B: this lane's object (index, image, label); k = -1: none
  float4 box;
  int rest, hi;       // objects [rest, hi) of the tile's images did not fit the 32 lanes
};
template <int TILE>
__device__ __forceinline__ void own_stage_a(const LossParams& p, int64_t tile, int lane, OwnStage& a) {
  const uint32_t SS = (uint32_t)(p.S * p.S);
  a.q0 = (uint32_t)tile * TILE;
  a.n_lo = (int)(a.q0 / SS);
  a.span = (int)((a.q0 + TILE - 1) / SS) - a.n_lo + 1;   // <= 31 images: the launcher asks for S >= 3
  a.off = lane <= a.span ? (int)__ldg(p.offsets + a.n_lo + lane) : 0;
}
__device__ __forceinline__ void own_stage_b(const LossParams& p, int lane, const OwnStage& a, OwnStage& b) {
  b.q0 = a.q0, b.n_lo = a.n_lo, b.span = a.span;
  const int lo = __shfl_sync(0xffffffffu, a.off, 0);
  b.hi = __shfl_sync(0xffffffffu, a.off, a.span);
  b.rest = lo + 32;
  b.k = lo + lane;
  b.img = a.n_lo;
  for (int j = 1; j < a.span; ++j) b.img += b.k >= __shfl_sync(0xffffffffu, a.off, j) ? 1 : 0;
  if (b.k < b.hi) {
    b.box = __ldg(reinterpret_cast<const float4*>(p.boxes + 4 * (int64_t)b.k));
    b.label = __ldg(p.labels + b.k);
  } else {
    b.k = -1;
  }
}
template <int TILE>
__device__ __forceinline__ void own_place(const LossParams& p, const OwnStage& b, int k, int img, const float4& box,
                                          int label, int32_t* own) {
  int cell;
  if (!object_cell(p, box, label, cell)) {
    atomicExch(p.status, 1);
    return;
  }
  const uint32_t q = (uint32_t)img * (uint32_t)(p.S * p.S) + (uint32_t)cell;
  if (q >= b.q0 && q - b.q0 < (uint32_t)TILE) atomicMax(own + (q - b.q0), k);
}
template <int TILE>
__device__ __forceinline__ void own_stage_c(const LossParams& p, int lane, const OwnStage& b, int32_t* own) {
  if (b.k >= 0) own_place<TILE>(p, b, b.k, b.img, b.box, b.label, own);
  for (int k = b.rest + lane; k < b.hi; k += 32) {   // images with many objects: the rest without the pipeline
    int img = b.n_lo;
    for (int j = 1; j < b.span; ++j) img += (int64_t)k >= __ldg(p.offsets + b.n_lo + j) ? 1 : 0;
    own_place<TILE>(p, b, k, img, __ldg(reinterpret_cast<const float4*>(p.boxes + 4 * (int64_t)k)), __ldg(p.labels + k),
                    own);
  }
}

// ---- K1 fast kernel: contiguous layout, TMA in / TMA out ----------------------------------------------
template <typename E, bool HAS_GRAD, int TILE, int STAGES, int NOUT, bool SIG = false, int LIST = 0>
__global__ void __launch_bounds__(TILE + (LIST == 2 ? 32 : 0)) loss_tma_kernel(const __grid_constant__ LossParams p) {
  constexpr int D = 30;
  // LIST 1: the target stage holds one int per cell (the owning object's index) instead of 30 floats.
  // LIST 2: no target stage at all -- warp 0 finds the owners of the tiles ahead from the object lists themselves
  // (own_stage_* below), so the call needs neither the pre-pass kernel nor its 4-byte-per-cell map.
  constexpr uint32_t PB = TILE * D * sizeof(E),
                     TB = LIST == 1 ? TILE * sizeof(int32_t) : (LIST == 2 ? 0 : TILE * D * sizeof(float)), GB = PB;
  static_assert(PB % 16 == 0 && TB % 16 == 0, "bulk copies move multiples of 16 bytes");
  static_assert(NOUT == 0 || NOUT >= 2, "NOUT = 0: gradient tile overwrites the pred stage in place; else >= 2 buffers");
  constexpr bool INPLACE = NOUT == 0;
  extern __shared__ __align__(128) unsigned char smem[];
  E* sp = reinterpret_cast<E*>(smem);
  float* st = reinterpret_cast<float*>(smem + STAGES * PB);
  E* so = reinterpret_cast<E*>(smem + STAGES * (PB + TB));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (PB + TB) + NOUT * GB);
  int32_t* own = reinterpret_cast<int32_t*>(bars + STAGES);   // LIST 2: [2][TILE] owners of the next two tiles
  (void)own;

  const int tid = threadIdx.x;
  // LIST 2: one more warp than the tile has cells; it owns no cell and finds the owners of the tiles ahead
  const bool helper = LIST == 2 && tid >= TILE;
  const int hl = tid - TILE;   // its lane
  const int64_t full = p.cells / TILE;  // tiles moved by the copy engine; the ragged tail goes direct
  const int64_t my_n = full > (int64_t)blockIdx.x ? (full - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const E* gp = reinterpret_cast<const E*>(p.pred);
  E* gg = reinterpret_cast<E*>(p.grad);
  uint64_t pol = 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
    pol = policy_evict_first();
  }
  __syncthreads();
  auto issue = [&](int64_t k) {
    const int s = (int)(k % STAGES);
    const int64_t off = ((int64_t)blockIdx.x + k * gridDim.x) * (TILE * D);
    mbar_arrive_expect_tx(&bars[s], PB + ((LIST == 1 && k + 1 >= my_n) ? 0u : TB));
    bulk_g2s(sp + s * (TILE * D), gp + off, PB, &bars[s], pol);
    if (LIST == 2) {
    } else if (LIST == 1) {
      // the stage of tile k carries the ownership map of tile k+1 (see the software pipeline below)
      if (k + 1 < my_n)
        bulk_g2s(reinterpret_cast<unsigned char*>(st) + s * TB,
                 p.cellobj + ((int64_t)blockIdx.x + (k + 1) * gridDim.x) * TILE, TB, &bars[s], pol);
    } else {
      bulk_g2s(st + s * (TILE * D), p.target + off, TB, &bars[s], pol);
    }
  };
  if (tid == 0)
    for (int64_t k = 0; k < my_n && k < STAGES; ++k) issue(k);

  CellSums sums = {0.f, 0.f, 0.f, 0.f};
  uint32_t m1 = 0, m2 = 0;
  // LIST: the object record of my cell in the NEXT tile, fetched while the current tile is being finished
  ObjFetch nxt = {make_float4(0.f, 0.f, 0.f, 0.f), 0, -1};
  if (LIST == 1 && my_n > 0) nxt = fetch_object(p, p.cellobj[(int64_t)blockIdx.x * TILE + tid]);
  // LIST 2.  The helper warp runs three stages ahead of the tile being evaluated, each a trip to memory further
  // along, and every stage has a whole tile period to land:
  //   A (tile k+4): lane j loads offsets[n_lo + j], the object ranges of the images the tile touches;
  //   B (tile k+3): lane t loads object lo + t of those images (box, label) and derives its image from A's offsets;
  //   C (tile k+2): the lanes drop their object's index into own[k & 1][cell - tile start] with atomicMax (the last
  //                 object of a cell wins, utils/YOLODataLoader.py:220); more than 32 objects: the rest, blocking;
  //   cell threads (tile k+1): read own[(k + 1) & 1][tid], put -1 back, fetch that object's record as LIST 1 does.
  // The CTA barrier every tile already has orders C's writes before the reads one iteration later, and the reads /
  // resets before C returns to the same buffer.  (With the stages inside warp 0 instead of a warp of their own the
  // tile barrier waits for them: 0.558 against 0.532 ms with the pre-pass at config-3 size.)
  OwnStage oa = {}, ob = {};
  const auto tile_at = [&](int64_t k) { return (int64_t)blockIdx.x + k * gridDim.x; };
  if (LIST == 2) {
    if (!helper) own[tid] = -1, own[TILE + tid] = -1;
    __syncthreads();
    if (helper && my_n > 0) {
      for (int j = 0; j < 2 && j < my_n; ++j) {   // tiles 0 and 1 straight through
        own_stage_a<TILE>(p, tile_at(j), hl, oa);
        own_stage_b(p, hl, oa, ob);
        own_stage_c<TILE>(p, hl, ob, own + j * TILE);
      }
      if (my_n > 2) own_stage_a<TILE>(p, tile_at(2), hl, oa), own_stage_b(p, hl, oa, ob);
      if (my_n > 3) own_stage_a<TILE>(p, tile_at(3), hl, oa);
    }
    __syncthreads();
    if (!helper) {
      if (my_n > 0) nxt = fetch_object(p, own[tid]);
      own[tid] = -1;
    }
    __syncthreads();   // buffer 0 is clean before stage C of tile 2 (first iteration) writes into it
  }
  using PIn = typename SmemIn<E>::type;
  using GOut = typename SmemOut<E>::type;
  for (int64_t k = 0; k < my_n; ++k) {
    const int s = (int)(k % STAGES), o = INPLACE ? 0 : (int)(k % (NOUT > 0 ? NOUT : 1));
    // in-place: every thread reads its own cell's 30 values before it overwrites them with the gradient
    E* gtile = INPLACE ? sp + s * (TILE * D) : so + o * (TILE * D);
    if (helper) {
      if (k + 2 < my_n) own_stage_c<TILE>(p, hl, ob, own + (k & 1) * TILE);
      if (k + 3 < my_n) own_stage_b(p, hl, oa, ob);
      if (k + 4 < my_n) own_stage_a<TILE>(p, tile_at(k + 4), hl, oa);
    } else {
      mbar_wait(&bars[s], (uint32_t)((k / STAGES) & 1));
      const PIn P{sp + s * (TILE * D) + tid * D};
      const SmemInF32 T{st + s * (TILE * D) + tid * D};
      const GOut G{gtile + tid * D};
      bool obj;
      if (LIST) {
        const ListTarget2 TL = list_target2(p, nxt);
        if (LIST == 2) {
          int32_t* o1 = own + ((k + 1) & 1) * TILE;
          if (k + 1 < my_n) nxt = fetch_object(p, o1[tid]);
          o1[tid] = -1;
        } else if (k + 1 < my_n) {  // gather for tile k+1 now; it lands behind the barrier and the next wait
          const int32_t* slots = reinterpret_cast<const int32_t*>(reinterpret_cast<const unsigned char*>(st) + s * TB);
          nxt = fetch_object(p, slots[tid]);
        }
        if (SIG)
          obj = cell_b2c20<HAS_GRAD>(SigIn<PIn>{P}, TL, SigOut<GOut>{G}, p, sums);
        else
          obj = cell_b2c20<HAS_GRAD>(P, TL, G, p, sums);
      } else if (SIG) {
        obj = cell_b2c20<HAS_GRAD>(SigIn<PIn>{P}, T, SigOut<GOut>{G}, p, sums);
      } else {
        obj = cell_b2c20<HAS_GRAD>(P, T, G, p, sums);
      }
      if (obj) note_object(m1, m2, ((int64_t)blockIdx.x + k * gridDim.x) * TILE + tid);
    }
    if (HAS_GRAD) {
      fence_async_smem();  // my shared-memory gradient writes -> visible to the copy engine
      if (!INPLACE && tid == 0) bulk_wait_read<(NOUT >= 2 ? NOUT - 2 : 0)>();  // buffer (k+1) % NOUT is free again
    }
    __syncthreads();
    if (tid == 0) {
      if (HAS_GRAD) {
        bulk_s2g(gg + ((int64_t)blockIdx.x + k * gridDim.x) * (TILE * D), gtile, GB, pol);
        bulk_commit();
      }
      if (k + STAGES < my_n) {
        if (HAS_GRAD && INPLACE) bulk_wait_read<0>();  // the store has drained stage s: it may be refilled
        issue(k + STAGES);
      }
    }
  }
  // ragged tail (< TILE cells): one CTA, straight from / to global memory
  const int64_t tail0 = full * TILE;
  if ((int64_t)blockIdx.x == full % gridDim.x && tail0 + tid < p.cells) {
    const int64_t q = tail0 + tid;
    const GlobIn<E> P{gp + q * D, 1, SIG};
    const GlobOut<E> G{HAS_GRAD ? gg + q * D : nullptr, 1, gp + q * D, 1, SIG};
    bool obj;
    if (LIST) {
      obj = cell_generic<HAS_GRAD, false>(P, list_targetS(p, q), G, p, sums);
    } else {
      const GlobIn<float> T{p.target + q * D, 1, false};
      obj = cell_generic<HAS_GRAD, false>(P, T, G, p, sums);
    }
    if (obj) note_object(m1, m2, q);
  }
  block_epilogue<E, HAS_GRAD, true>(sums, m1, m2, p);
}

template <typename E, bool HAS_GRAD, int TILE, int STAGES, int NOUT, bool SIG = false, int LIST = 0>
int launch_tma(const LossParams& p, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * TILE * (30 * sizeof(E) + (LIST == 1 ? 4 : (LIST == 2 ? 0 : 120))) +
                          (size_t)NOUT * TILE * 30 * sizeof(E) + STAGES * sizeof(uint64_t) +
                          (LIST == 2 ? 2 * TILE * sizeof(int32_t) : 0);
  auto kern = loss_tma_kernel<E, HAS_GRAD, TILE, STAGES, NOUT, SIG, LIST>;
  constexpr int threads = TILE + (LIST == 2 ? 32 : 0);
  static KernelPrep prep;   // one per kernel instantiation: attribute / occupancy queries once per device
  int sms = kNumSMs, per_sm = 1;
  if (int rc = prepare_kernel(prep, kern, threads, smem, true, &sms, &per_sm)) return rc;
  const int64_t tiles = p.cells / TILE;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > tiles) grid = tiles;
  if (grid > kMaxGrid) grid = kMaxGrid;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, threads, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

template <typename E, bool HAS_GRAD>
int launch_tma_variant(const LossParams& p, int variant, cudaStream_t stream) {
  if (p.list_mode == 2) {  // object-list targets, owners found in the kernel
    if (p.logits) return launch_tma<E, HAS_GRAD, 128, 2, 2, true, 2>(p, stream);
    return launch_tma<E, HAS_GRAD, 128, 2, 2, false, 2>(p, stream);
  }
  if (p.list_mode) {  // object-list targets behind the ownership map of the pre-pass
    if (p.logits) return launch_tma<E, HAS_GRAD, 128, 2, 2, true, 1>(p, stream);
    switch (variant) {  // other shapes measured slower
      default: return launch_tma<E, HAS_GRAD, 128, 2, 2, false, 1>(p, stream);
    }
  }
  if (p.logits) return launch_tma<E, HAS_GRAD, 128, 2, 2, true>(p, stream);  // one launch shape with the fused head
  switch (variant) {
    case 0:
    case 1: return launch_tma<E, HAS_GRAD, 128, 2, 2>(p, stream);
    case 2: return launch_tma<E, HAS_GRAD, 128, 3, 2>(p, stream);
    case 3: return launch_tma<E, HAS_GRAD, 64, 3, 2>(p, stream);
    case 5: return launch_tma<E, HAS_GRAD, 256, 2, 2>(p, stream);
    case 8: return launch_tma<E, HAS_GRAD, 128, 2, 0>(p, stream);   // in-place gradient tile: 61 KB, 3 CTAs/SM
    case 13: return launch_tma<E, HAS_GRAD, 256, 2, 0>(p, stream);  // 123 KB, 1 CTA/SM
    default: return YOLO1_ERR_ARG;
  }
}

// ---- the same pipeline for any (B, C): contiguous [N,S,S,5B+C] tensors, runtime channel count -----------------
// (other detectors' heads, e.g. 80 classes).  One thread per cell through cell_generic on scalar shared-memory
// accessors; tile = blockDim cells, sized by the host so that two stages + two output buffers fit three times per SM.
template <typename E, bool HAS_GRAD>
__global__ void __launch_bounds__(128) loss_tma_any_kernel(const __grid_constant__ LossParams p) {
  constexpr int STAGES = 2, NOUT = 2;
  const int D = 5 * p.B + p.C, TILE = blockDim.x, tile_elems = TILE * D;
  const uint32_t PB = tile_elems * sizeof(E), TB = tile_elems * sizeof(float), GB = PB;
  extern __shared__ __align__(128) unsigned char smem[];
  E* sp = reinterpret_cast<E*>(smem);
  float* st = reinterpret_cast<float*>(smem + STAGES * PB);
  E* so = reinterpret_cast<E*>(smem + STAGES * (PB + TB));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (PB + TB) + NOUT * GB);
  int32_t* own = reinterpret_cast<int32_t*>(bars + STAGES);   // LIST 2: [2][TILE] owners of the next two tiles
  (void)own;
  const int tid = threadIdx.x;
  const int64_t full = p.cells / TILE;
  const int64_t my_n = full > (int64_t)blockIdx.x ? (full - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const E* gp = reinterpret_cast<const E*>(p.pred);
  E* gg = reinterpret_cast<E*>(p.grad);
  const bool sig = p.logits != 0;
  uint64_t pol = 0;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
    pol = policy_evict_first();
  }
  __syncthreads();
  auto issue = [&](int64_t k) {
    const int s = (int)(k % STAGES);
    const int64_t off = ((int64_t)blockIdx.x + k * gridDim.x) * tile_elems;
    mbar_arrive_expect_tx(&bars[s], PB + TB);
    bulk_g2s(sp + s * tile_elems, gp + off, PB, &bars[s], pol);
    bulk_g2s(st + s * tile_elems, p.target + off, TB, &bars[s], pol);
  };
  if (tid == 0)
    for (int64_t k = 0; k < my_n && k < STAGES; ++k) issue(k);
  CellSums sums = {0.f, 0.f, 0.f, 0.f};
  uint32_t m1 = 0, m2 = 0;
  for (int64_t k = 0; k < my_n; ++k) {
    const int s = (int)(k % STAGES), o = (int)(k % NOUT);
    mbar_wait(&bars[s], (uint32_t)((k / STAGES) & 1));
    if (HAS_GRAD) {
      // most of the gradient tile is zero: clear it with 16-byte stores, the cells then write what is not
      // (buffer o was released by the wait_group.read before the previous iteration's barrier)
      uint4* z4 = reinterpret_cast<uint4*>(so + o * tile_elems);
      for (int t = tid; t < (int)(GB / 16); t += TILE) z4[t] = make_uint4(0u, 0u, 0u, 0u);
      __syncthreads();
    }
    const E* zc = sp + s * tile_elems + tid * D;
    const SmemInS<E> P{zc, sig};
    const SmemInS<float> T{st + s * tile_elems + tid * D, false};
    const SmemOutS<E> G{so + o * tile_elems + tid * D, zc, sig};
    if (cell_generic<HAS_GRAD, false, true>(P, T, G, p, sums))
      note_object(m1, m2, ((int64_t)blockIdx.x + k * gridDim.x) * TILE + tid);
    if (HAS_GRAD) {
      fence_async_smem();
      if (tid == 0) bulk_wait_read<NOUT - 2>();
    }
    __syncthreads();
    if (tid == 0) {
      if (HAS_GRAD) {
        bulk_s2g(gg + ((int64_t)blockIdx.x + k * gridDim.x) * tile_elems, so + o * tile_elems, GB, pol);
        bulk_commit();
      }
      if (k + STAGES < my_n) issue(k + STAGES);
    }
  }
  const int64_t tail0 = full * TILE;
  if ((int64_t)blockIdx.x == full % gridDim.x && tail0 + tid < p.cells) {
    const int64_t q = tail0 + tid;
    const GlobIn<E> P{gp + q * D, 1, sig};
    const GlobIn<float> T{p.target + q * D, 1, false};
    const GlobOut<E> G{HAS_GRAD ? gg + q * D : nullptr, 1, gp + q * D, 1, sig};
    if (cell_generic<HAS_GRAD, false>(P, T, G, p, sums)) note_object(m1, m2, q);
  }
  block_epilogue<E, HAS_GRAD, true>(sums, m1, m2, p);
}

template <typename E, bool HAS_GRAD>
int launch_tma_any(const LossParams& p, cudaStream_t stream) {
  const int D = 5 * p.B + p.C;
  int tile = 128;
  while (tile > 32 && (size_t)tile * D * (4 * sizeof(E) + 8) > 72 * 1024) tile >>= 1;   // 3 CTAs per SM
  const size_t smem = (size_t)tile * D * (4 * sizeof(E) + 8) + 2 * sizeof(uint64_t);
  if (smem > 200 * 1024) return YOLO1_ERR_UNSUPPORTED;
  auto kern = loss_tma_any_kernel<E, HAS_GRAD>;
  static KernelPrep prep;
  int sms = kNumSMs, per_sm = 1;
  if (int rc = prepare_kernel(prep, kern, tile, smem, true, &sms, &per_sm)) return rc;
  const int64_t tiles = p.cells / tile;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > tiles) grid = tiles;
  if (grid > kMaxGrid) grid = kMaxGrid;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, tile, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace

int launch_loss_nhwc_any(const LossParams& p, bool bf16, bool has_grad, cudaStream_t stream) {
  if (bf16) return has_grad ? launch_tma_any<__nv_bfloat16, true>(p, stream) : launch_tma_any<__nv_bfloat16, false>(p, stream);
  return has_grad ? launch_tma_any<float, true>(p, stream) : launch_tma_any<float, false>(p, stream);
}

int launch_loss_nhwc(const LossParams& p, bool bf16, bool has_grad, int variant, cudaStream_t stream) {
  if (bf16) return has_grad ? launch_tma_variant<__nv_bfloat16, true>(p, variant, stream)
                            : launch_tma_variant<__nv_bfloat16, false>(p, variant, stream);
  return has_grad ? launch_tma_variant<float, true>(p, variant, stream)
                  : launch_tma_variant<float, false>(p, variant, stream);
}

}  // namespace yolo1
