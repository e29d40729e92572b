// loss_planar_sparse.cu -- K1 "confidence-first" form for the channel-planar view (sm_100a).
//
// The backbone hands the loss a permuted NCHW view (OriginResNet.py:189): an image is 30 planes of S*S values.  A
// cell without object -- 98.5 % of them -- needs pred[0:2] and nothing else of pred, and in THIS layout those two
// values live in the image's first two planes: one contiguous, fully used run of 2*S*S elements.  (On the NHWC layout
// the same idea fetches every 128-byte line anyway and gains nothing: profiles/ncu_sparse_r2.md.)  So:
//
//   pred   : every thread loads its cell's two confidences straight from planes 0 and 1 (coalesced 4-byte loads, lanes
//            <-> consecutive cells); only a cell that holds an object gathers its other 28 channels (28 sectors);
//   target : dense NHWC rows by bulk (TMA) copy, 2 stages -- or, for object-list targets, the 4-byte ownership slot
//            (coalesced) and the object record;
//   grad   : the planar gradient tile of whole images is kept in shared memory BETWEEN tiles: planes 2..29 are zero
//            except at object cells, so a thread writes its two confidence gradients, an object cell its 30 values,
//            and the cell that was an object two tiles ago (same buffer) clears its 28 values again.  One bulk (TMA)
//            store per tile writes the 30 planes of the tile's images.
//
// Traffic per cell, fp32: 8 (pred) + 120 (target) + 120 (grad) = 248 B against 360 for the dense planar kernels
// (loss_planar.cu, loss_ws.cu); with object lists 8 + 4 + 120 = 132 B against 248.  bf16 pred / grad: 184 against 240.
#include "loss_common.cuh"

namespace yolo1 {
namespace {

// pred of one cell in the planar layout: everything in registers.  The confidences come from the prefetch; the other
// 28 channels are gathered up front, all loads in flight together, and only for a cell that holds an object.  (They
// must not be fetched on demand inside cell_b2c20: the compiler cannot prove that the global pred pointer and the
// shared-memory gradient pointer do not alias, so every load behind a gradient store waits for it -- ten DRAM round
// trips in a row per object cell; measured 1.19 ms instead of 0.52 ms at config-3 size.)
struct PlanarRegsIn {
  float v[30];
  __device__ __forceinline__ float2 ld2(int c) const { return make_float2(v[c], v[c + 1]); }
};
// gradient tile that is zero wherever nobody wrote: the zero stores of a cell without object are dropped
template <typename E>
struct PlanarKeepZeroOut {
  E* cell;
  int plane;
  __device__ __forceinline__ void st2(int, float, float) const {}   // called with zeros only (cell_b2c20)
  __device__ __forceinline__ void st2p(int c, float x, float y, float, float) const {
    st_elem(cell + c * plane, x);
    st_elem(cell + (c + 1) * plane, y);
  }
};

template <typename E, bool HAS_GRAD, bool SIG, bool LIST>
__global__ void __launch_bounds__(256) loss_planar_sparse_kernel(const __grid_constant__ LossParams p, int tile_imgs) {
  constexpr int D = 30, STAGES = 2, NOUT = 2;
  const int SS = p.S * p.S, tile_cells = tile_imgs * SS, tile_elems = tile_cells * D;
  const uint32_t GB = tile_elems * sizeof(E), TB = LIST ? 0u : tile_elems * sizeof(float);
  extern __shared__ __align__(128) unsigned char smem[];
  E* so = reinterpret_cast<E*>(smem);                                   // NOUT gradient tiles
  float* st = reinterpret_cast<float*>(smem + (HAS_GRAD ? NOUT : 0) * (size_t)GB);   // STAGES target tiles (dense)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (HAS_GRAD ? NOUT : 0) * (size_t)GB + STAGES * (size_t)TB);

  const int tid = threadIdx.x;
  const int64_t n_imgs = p.cells / SS;
  const int64_t full = n_imgs / tile_imgs;
  const int64_t my_n = full > (int64_t)blockIdx.x ? (full - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const E* gp = reinterpret_cast<const E*>(p.pred);
  E* gg = reinterpret_cast<E*>(p.grad);
  uint64_t pol = 0;
  if (tid == 0) {
    if (!LIST) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
      mbar_fence_init();
    }
    pol = policy_evict_first();
  }
  if (HAS_GRAD) {   // both gradient tiles start out zero
    uint4* z = reinterpret_cast<uint4*>(so);
    for (int t = tid; t < (int)(NOUT * GB / 16); t += blockDim.x) z[t] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  auto issue = [&](int64_t k) {   // dense target: the tile's rows, one bulk copy
    const int s = (int)(k % STAGES);
    mbar_arrive_expect_tx(&bars[s], TB);
    bulk_g2s(st + s * tile_elems, p.target + ((int64_t)blockIdx.x + k * gridDim.x) * tile_elems, TB, &bars[s], pol);
  };
  if (!LIST && tid == 0)
    for (int64_t k = 0; k < my_n && k < STAGES; ++k) issue(k);

  const bool active = tid < tile_cells;
  const int img = active ? tid / SS : 0, r = active ? tid - img * SS : 0;
  const int poff = img * (D * SS) + r;   // my cell's channel 0 inside a planar tile
  CellSums sums = {0.f, 0.f, 0.f, 0.f};
  uint32_t m1 = 0, m2 = 0;
  uint32_t was_obj = 0;                  // bit o: my cell of gradient tile o holds an object's 28 extra values
  // Software pipeline, one tile deep, in registers: while tile k's gradient tile is fenced, synchronised and stored,
  // the loads of tile k+1 -- its two confidences and, if its target says "object", its other 28 channels -- are
  // already in flight (issued right after tile k was evaluated, so the registers are the same ones).  Without it the
  // whole CTA waits at the tile barrier for the one lane that gathers (95 % of the 196-cell tiles hold an object):
  // 0.75 ms instead of 0.50 ms.  LIST: the ownership slot is read two tiles ahead, the object record one.
  PlanarRegsIn V;
  bool v_obj = false;
  ObjFetch v_rec = {make_float4(0.f, 0.f, 0.f, 0.f), 0, -1};
  int32_t slot_ahead = -1;               // LIST: ownership slot of my cell in tile k+1
  auto tile_of = [&](int64_t k) { return (int64_t)blockIdx.x + k * gridDim.x; };
  auto load_tile = [&](int64_t k, bool obj) {   // pred of my cell in tile k -> V (asynchronous: first use next iteration)
    const E* cell = gp + tile_of(k) * tile_elems + poff;
    V.v[0] = ld_elem(cell), V.v[1] = ld_elem(cell + SS);
    if (obj) {
#pragma unroll
      for (int c = 2; c < D; ++c) V.v[c] = ld_elem(cell + c * SS);
    } else {
#pragma unroll
      for (int c = 2; c < D; ++c) V.v[c] = 0.f;
    }
  };
  if (my_n > 0) {
    if (LIST) {
      if (active) {
        v_rec = fetch_object(p, __ldg(p.cellobj + tile_of(0) * tile_cells + tid));
        v_obj = v_rec.k >= 0;
        if (my_n > 1) slot_ahead = __ldg(p.cellobj + tile_of(1) * tile_cells + tid);
      }
    } else {
      mbar_wait(&bars[0], 0);
      v_obj = active && st[tid * D] == 1.0f;   // v1Loss.py:28
    }
    if (active) load_tile(0, v_obj);
  }
  for (int64_t k = 0; k < my_n; ++k) {
    const int64_t tile = tile_of(k);
    const int s = (int)(k % STAGES), o = (int)(k % NOUT);
    if (active) {
      E* gcell = so + o * tile_elems + poff;
      if (HAS_GRAD && ((was_obj >> o) & 1u)) {   // the store that read this buffer two tiles ago has drained
#pragma unroll
        for (int c = 2; c < D; ++c) st_elem(gcell + c * SS, 0.f);
      }
      const PlanarKeepZeroOut<E> G{gcell, SS};
      bool obj;
      // (SIG: SigIn applies the sigmoid to whatever the cell reads -- two values for a cell without object --, SigOut
      //  multiplies by p (1 - p) with the probabilities cell_b2c20 hands back)
      if (LIST) {
        const ListTarget2 TL = list_target2(p, v_rec);
        if (SIG)
          obj = cell_b2c20<HAS_GRAD>(SigIn<PlanarRegsIn>{V}, TL, SigOut<PlanarKeepZeroOut<E>>{G}, p, sums);
        else
          obj = cell_b2c20<HAS_GRAD>(V, TL, G, p, sums);
      } else {
        const SmemInF32 T{st + s * tile_elems + tid * D};
        if (SIG)
          obj = cell_b2c20<HAS_GRAD>(SigIn<PlanarRegsIn>{V}, T, SigOut<PlanarKeepZeroOut<E>>{G}, p, sums);
        else
          obj = cell_b2c20<HAS_GRAD>(V, T, G, p, sums);
      }
      was_obj = (was_obj & ~(1u << o)) | ((obj ? 1u : 0u) << o);
      if (obj) note_object(m1, m2, tile * tile_cells + tid);
    }
    // tile k+1: is my cell an object there, and its loads
    if (k + 1 < my_n) {
      if (LIST) {
        if (active) {
          v_rec = fetch_object(p, slot_ahead);
          v_obj = slot_ahead >= 0;
          if (k + 2 < my_n) slot_ahead = __ldg(p.cellobj + tile_of(k + 2) * tile_cells + tid);
        }
      } else {
        mbar_wait(&bars[s ^ 1], (uint32_t)(((k + 1) / STAGES) & 1));   // issued one iteration ago
        v_obj = active && st[(s ^ 1) * tile_elems + tid * D] == 1.0f;
      }
      if (active) load_tile(k + 1, v_obj);
    }
    if (HAS_GRAD) {
      fence_async_smem();                              // my gradient writes -> visible to the copy engine
      if (tid == 0) bulk_wait_read<NOUT - 2>();        // tile (k+1) % NOUT has left: it may be written next
    }
    __syncthreads();
    if (tid == 0) {
      if (HAS_GRAD) {
        bulk_s2g(gg + tile * tile_elems, so + o * tile_elems, GB, pol);
        bulk_commit();
      }
      if (!LIST && k + STAGES < my_n) issue(k + STAGES);   // every thread is done with target stage s
    }
  }
  // ragged tail (< tile_imgs images): one CTA, strided global accesses
  const int64_t tail0 = full * tile_cells;
  if ((int64_t)blockIdx.x == full % gridDim.x && tail0 + tid < p.cells && active) {
    const int64_t q = tail0 + tid;
    const E* zq = gp + cell_offset(p.ps, q, p.S);
    const GlobIn<E> P{zq, p.ps[3], SIG};
    const GlobOut<E> G{HAS_GRAD ? gg + cell_offset(p.gs, q, p.S) : nullptr, p.gs[3], zq, p.ps[3], SIG};
    bool obj;
    if (LIST) {
      obj = cell_generic<HAS_GRAD, false>(P, list_targetS(p, q), G, p, sums);
    } else {
      const GlobIn<float> T{p.target + cell_offset(p.ts, q, p.S), p.ts[3], false};
      obj = cell_generic<HAS_GRAD, false>(P, T, G, p, sums);
    }
    if (obj) note_object(m1, m2, q);
  }
  block_epilogue<E, HAS_GRAD, true>(sums, m1, m2, p);
}

template <typename E, bool HAS_GRAD, bool SIG, bool LIST>
int launch_planar_sparse_t(const LossParams& p, int tile_imgs, cudaStream_t stream) {
  const int tile_cells = tile_imgs * p.S * p.S;
  const size_t smem = (HAS_GRAD ? 2 : 0) * (size_t)tile_cells * 30 * sizeof(E) + (LIST ? 0 : 2 * (size_t)tile_cells * 120) +
                      2 * sizeof(uint64_t);
  const int threads = (tile_cells + 31) / 32 * 32;
  if (threads > 256 || smem > 200 * 1024) return YOLO1_ERR_UNSUPPORTED;
  auto kern = loss_planar_sparse_kernel<E, HAS_GRAD, SIG, LIST>;
  static KernelPrep prep;
  int sms = kNumSMs, per_sm = 1;
  if (int rc = prepare_kernel(prep, kern, threads, smem, true, &sms, &per_sm)) return rc;
  const int64_t tiles = p.cells / tile_cells;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > tiles) grid = tiles;
  if (grid > kMaxGrid) grid = kMaxGrid;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, threads, smem, stream>>>(p, tile_imgs);
  return (int)cudaGetLastError();
}

template <typename E, bool HAS_GRAD>
int launch_planar_sparse_e(const LossParams& p, int tile_imgs, cudaStream_t stream) {
  if (p.list_mode)
    return p.logits ? launch_planar_sparse_t<E, HAS_GRAD, true, true>(p, tile_imgs, stream)
                    : launch_planar_sparse_t<E, HAS_GRAD, false, true>(p, tile_imgs, stream);
  return p.logits ? launch_planar_sparse_t<E, HAS_GRAD, true, false>(p, tile_imgs, stream)
                  : launch_planar_sparse_t<E, HAS_GRAD, false, false>(p, tile_imgs, stream);
}

}  // namespace

int launch_loss_planar_sparse(const LossParams& p, bool bf16, bool has_grad, int tile_imgs, cudaStream_t stream) {
  if (bf16) return has_grad ? launch_planar_sparse_e<__nv_bfloat16, true>(p, tile_imgs, stream)
                            : launch_planar_sparse_e<__nv_bfloat16, false>(p, tile_imgs, stream);
  return has_grad ? launch_planar_sparse_e<float, true>(p, tile_imgs, stream)
                  : launch_planar_sparse_e<float, false>(p, tile_imgs, stream);
}

}  // namespace yolo1
