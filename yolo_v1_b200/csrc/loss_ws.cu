// loss_ws.cu -- K1, warp-specialised form: a dedicated producer warp drives the copy engine, the consumer warps
// only compute.  Same arithmetic and the same results as loss_nhwc.cu / loss_planar.cu (loss_common.cuh).
//
// Why: with the gradient tile written IN PLACE over the pred stage (the only way two CTAs fit an SM once a tile
// is a whole 14x14 image, 47 KB per stage), the thread that issues the bulk store must wait for the store to
// drain the stage before it can refill it.  In the single-role kernels that thread also owns a cell, so every
// other thread of the CTA waits for it at the next barrier.  Here lane 0 of an extra warp does nothing but
//     wait done[s] -> bulk store tile k-STAGES -> wait until the store has read the stage -> bulk load tile k
// while the consumers run   wait full[s] -> compute in place -> fence.proxy.async -> arrive done[s]
// with no CTA-wide barrier in the loop.  full[] are transaction barriers (expect_tx), done[] count one arrival
// per consumer warp.
#include "loss_common.cuh"

namespace yolo1 {
namespace {

template <typename E, bool HAS_GRAD, bool PLANAR, int STAGES, bool SIG, bool LIST>
__global__ void __launch_bounds__(288) loss_ws_kernel(const __grid_constant__ LossParams p, int tile_cells) {
  constexpr int D = 30;
  const int SS = p.S * p.S, tile_elems = tile_cells * D;
  const uint32_t PB = tile_elems * sizeof(E), TB = LIST ? tile_cells * sizeof(int32_t) : tile_elems * sizeof(float);
  extern __shared__ __align__(128) unsigned char smem[];
  E* sp = reinterpret_cast<E*>(smem);
  unsigned char* st = smem + STAGES * PB;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * (PB + TB));
  uint64_t* done = full + STAGES;

  const int tid = threadIdx.x, lane = tid & 31;
  const int n_cons = blockDim.x - 32;           // consumer threads; the last warp is the producer
  const bool producer = tid >= n_cons;
  const int64_t full_tiles = p.cells / tile_cells;
  const int64_t my_n = full_tiles > (int64_t)blockIdx.x ? (full_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const E* gp = reinterpret_cast<const E*>(p.pred);
  E* gg = reinterpret_cast<E*>(p.grad);
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], n_cons / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  CellSums sums = {0.f, 0.f, 0.f, 0.f};
  uint32_t m1 = 0, m2 = 0;
  if (producer) {
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      auto tile_of = [&](int64_t k) { return (int64_t)blockIdx.x + k * gridDim.x; };
      auto store = [&](int64_t j) {
        bulk_s2g(gg + tile_of(j) * tile_elems, sp + (j % STAGES) * tile_elems, PB, pol);
        bulk_commit();
      };
      for (int64_t k = 0; k < my_n; ++k) {
        const int s = (int)(k % STAGES);
        if (k >= STAGES) {
          mbar_wait(&done[s], (uint32_t)(((k / STAGES) - 1) & 1));   // tile k-STAGES is finished in this stage
          if (HAS_GRAD) {
            store(k - STAGES);
            bulk_wait_read<0>();                                     // ... and has left it
          }
        }
        const bool with_t = !(LIST && k + 1 >= my_n);
        mbar_arrive_expect_tx(&full[s], PB + (with_t ? TB : 0u));
        bulk_g2s(sp + s * tile_elems, gp + tile_of(k) * tile_elems, PB, &full[s], pol);
        if (LIST) {
          if (with_t) bulk_g2s(st + s * TB, p.cellobj + tile_of(k + 1) * tile_cells, TB, &full[s], pol);
        } else {
          bulk_g2s(st + s * TB, p.target + tile_of(k) * tile_elems, TB, &full[s], pol);
        }
      }
      for (int64_t j = my_n > STAGES ? my_n - STAGES : 0; j < my_n; ++j) {   // drain
        mbar_wait(&done[j % STAGES], (uint32_t)((j / STAGES) & 1));
        if (HAS_GRAD) store(j);
      }
      if (HAS_GRAD) {
        bulk_wait_all<0>();
        fence_async_all();
      }
    }
  } else {
    ObjFetch nxt = {make_float4(0.f, 0.f, 0.f, 0.f), 0, -1};
    if (LIST && my_n > 0 && tid < tile_cells) nxt = fetch_object(p, p.cellobj[(int64_t)blockIdx.x * tile_cells + tid]);
    const int img = tid / SS, r = tid - img * SS;
    const int poff = PLANAR ? img * (D * SS) + r : tid * D;   // my cell's first channel inside a pred tile
    const int pstep = PLANAR ? SS : 1;                        // distance between its channels
    for (int64_t k = 0; k < my_n; ++k) {
      const int s = (int)(k % STAGES);
      mbar_wait(&full[s], (uint32_t)((k / STAGES) & 1));
      if (tid < tile_cells) {
        E* cell = sp + s * tile_elems + poff;
        const PlanarIn<E> P{cell, pstep};
        const PlanarOut<E> G{cell, pstep};       // in place: a thread reads its cell before it overwrites it
        bool obj;
        if (LIST) {
          const ListTarget2 TL = list_target2(p, nxt);
          if (k + 1 < my_n) nxt = fetch_object(p, reinterpret_cast<const int32_t*>(st + s * TB)[tid]);
          if (SIG)
            obj = cell_b2c20<HAS_GRAD>(SigIn<PlanarIn<E>>{P}, TL, SigOut<PlanarOut<E>>{G}, p, sums);
          else
            obj = cell_b2c20<HAS_GRAD>(P, TL, G, p, sums);
        } else {
          const SmemInF32 T{reinterpret_cast<const float*>(st + s * TB) + tid * D};
          if (SIG)
            obj = cell_b2c20<HAS_GRAD>(SigIn<PlanarIn<E>>{P}, T, SigOut<PlanarOut<E>>{G}, p, sums);
          else
            obj = cell_b2c20<HAS_GRAD>(P, T, G, p, sums);
        }
        if (obj) note_object(m1, m2, ((int64_t)blockIdx.x + k * gridDim.x) * tile_cells + tid);
      }
      if (HAS_GRAD) fence_async_smem();   // my gradient writes -> visible to the copy engine
      __syncwarp();
      if (lane == 0) mbar_arrive(&done[s]);
    }
  }
  // ragged tail (< one tile): one CTA, strided global accesses
  const int64_t tail0 = full_tiles * tile_cells;
  if ((int64_t)blockIdx.x == full_tiles % gridDim.x && tail0 + tid < p.cells && tid < tile_cells) {
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
  block_epilogue<E, HAS_GRAD, false>(sums, m1, m2, p);   // the producer already waited for its own bulk stores
}

template <typename E, bool HAS_GRAD, bool PLANAR, int STAGES, bool SIG, bool LIST>
int launch_ws_t(const LossParams& p, int tile_cells, cudaStream_t stream) {
  const size_t smem = (size_t)STAGES * tile_cells * (30 * sizeof(E) + (LIST ? 4 : 120)) + 2 * STAGES * sizeof(uint64_t);
  const int threads = (tile_cells + 31) / 32 * 32 + 32;
  if (threads > 288 || smem > 227 * 1024) return YOLO1_ERR_UNSUPPORTED;
  auto kern = loss_ws_kernel<E, HAS_GRAD, PLANAR, STAGES, SIG, LIST>;
  static KernelPrep prep;   // one per kernel instantiation: attribute / occupancy queries once per device
  int sms = kNumSMs, per_sm = 1;
  if (int rc = prepare_kernel(prep, kern, threads, smem, true, &sms, &per_sm)) return rc;
  const int64_t tiles = p.cells / tile_cells;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > tiles) grid = tiles;
  if (grid > kMaxGrid) grid = kMaxGrid;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, threads, smem, stream>>>(p, tile_cells);
  return (int)cudaGetLastError();
}

template <typename E, bool HAS_GRAD, bool PLANAR, int STAGES>
int launch_ws_flags(const LossParams& p, int tile_cells, cudaStream_t stream) {
  if (p.logits)
    return p.list_mode ? launch_ws_t<E, HAS_GRAD, PLANAR, STAGES, true, true>(p, tile_cells, stream)
                       : launch_ws_t<E, HAS_GRAD, PLANAR, STAGES, true, false>(p, tile_cells, stream);
  return p.list_mode ? launch_ws_t<E, HAS_GRAD, PLANAR, STAGES, false, true>(p, tile_cells, stream)
                     : launch_ws_t<E, HAS_GRAD, PLANAR, STAGES, false, false>(p, tile_cells, stream);
}

template <bool PLANAR, int STAGES>
int launch_ws_types(const LossParams& p, bool bf16, bool has_grad, int tile_cells, cudaStream_t stream) {
  if (bf16) return has_grad ? launch_ws_flags<__nv_bfloat16, true, PLANAR, STAGES>(p, tile_cells, stream)
                            : launch_ws_flags<__nv_bfloat16, false, PLANAR, STAGES>(p, tile_cells, stream);
  return has_grad ? launch_ws_flags<float, true, PLANAR, STAGES>(p, tile_cells, stream)
                  : launch_ws_flags<float, false, PLANAR, STAGES>(p, tile_cells, stream);
}

}  // namespace

int launch_loss_ws(const LossParams& p, bool bf16, bool has_grad, bool is_planar, int tile_cells, int stages,
                   cudaStream_t stream) {
  // Instantiated for the channel-planar view only: for contiguous NHWC tensors the single-role kernel with separate
  // output buffers (loss_nhwc.cu) measured faster than every in-place shape (tools/tune_loss.py).
  if (!is_planar) return YOLO1_ERR_UNSUPPORTED;
  // three in-place stages measured best (fp32 0.753 vs 0.766 ms with two, bf16 0.494 vs 0.502); only that shape is built
  (void)stages;
  return launch_ws_types<true, 3>(p, bf16, has_grad, tile_cells, stream);
}

}  // namespace yolo1
