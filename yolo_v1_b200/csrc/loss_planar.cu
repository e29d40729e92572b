// loss_planar.cu -- K1 streaming kernel for the channel-planar (permuted NCHW) view (see loss_common.cuh).
#include "loss_common.cuh"

namespace yolo1 {
namespace {

// ---- K1 fast kernel, channel-planar pred/grad: the backbone's permuted NCHW view (OriginResNet.py:189) ------
// pred / grad are [N][30][S*S] in memory (element strides (30 S^2, S, 1, S^2)), target is contiguous NHWC.
// An image's 30 planes are one contiguous block, so a tile of `tile_imgs` whole images still moves with one
// bulk copy per tensor; one thread per cell reads its channels S*S elements apart (conflict-free) and writes
// the gradient tile in the same planar layout, so `permute`'s backward stays a free view.
template <typename E, bool HAS_GRAD, int STAGES, int NOUT, bool SIG = false, bool LIST = false>
__global__ void __launch_bounds__(256) loss_tma_planar_kernel(const __grid_constant__ LossParams p, int tile_imgs) {
  constexpr int D = 30;
  const int SS = p.S * p.S, tile_cells = tile_imgs * SS, tile_elems = tile_cells * D;
  const uint32_t PB = tile_elems * sizeof(E), TB = LIST ? tile_cells * sizeof(int32_t) : tile_elems * sizeof(float), GB = PB;
  extern __shared__ __align__(128) unsigned char smem[];
  E* sp = reinterpret_cast<E*>(smem);
  float* st = reinterpret_cast<float*>(smem + STAGES * PB);
  E* so = reinterpret_cast<E*>(smem + STAGES * (PB + TB));   // NOUT == 0: the gradient overwrites the pred stage
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (PB + TB) + NOUT * GB);
  constexpr bool INPLACE = NOUT == 0;

  const int tid = threadIdx.x;
  const int64_t n_imgs = p.cells / SS;
  const int64_t full = n_imgs / tile_imgs;
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
    const int64_t off = ((int64_t)blockIdx.x + k * gridDim.x) * tile_elems;
    mbar_arrive_expect_tx(&bars[s], PB + ((LIST && k + 1 >= my_n) ? 0u : TB));
    bulk_g2s(sp + s * tile_elems, gp + off, PB, &bars[s], pol);
    if (LIST) {
      if (k + 1 < my_n)  // the stage of tile k carries the ownership map of tile k+1
        bulk_g2s(reinterpret_cast<unsigned char*>(st) + s * TB,
                 p.cellobj + ((int64_t)blockIdx.x + (k + 1) * gridDim.x) * tile_cells, TB, &bars[s], pol);
    } else {
      bulk_g2s(st + s * tile_elems, p.target + off, TB, &bars[s], pol);
    }
  };
  if (tid == 0)
    for (int64_t k = 0; k < my_n && k < STAGES; ++k) issue(k);

  CellSums sums = {0.f, 0.f, 0.f, 0.f};
  uint32_t m1 = 0, m2 = 0;
  ObjFetch nxt = {make_float4(0.f, 0.f, 0.f, 0.f), 0, -1};
  if (LIST && my_n > 0 && tid < tile_cells) nxt = fetch_object(p, p.cellobj[(int64_t)blockIdx.x * tile_cells + tid]);
  const int img = tid / SS, r = tid - img * SS;   // my cell inside a tile
  const int poff = img * (D * SS) + r;
  for (int64_t k = 0; k < my_n; ++k) {
    const int s = (int)(k % STAGES);
    mbar_wait(&bars[s], (uint32_t)((k / STAGES) & 1));
    // in place: a thread reads its cell's 30 values before it overwrites them with the gradient
    E* gtile = INPLACE ? sp + s * tile_elems : so + (int)(k % (NOUT > 0 ? NOUT : 1)) * tile_elems;
    if (tid < tile_cells) {
      const PlanarIn<E> P{sp + s * tile_elems + poff, SS};
      const SmemInF32 T{st + s * tile_elems + tid * D};
      const PlanarOut<E> G{gtile + poff, SS};
      bool obj;
      if (LIST) {
        const ListTarget2 TL = list_target2(p, nxt);
        if (k + 1 < my_n) {
          const int32_t* slots = reinterpret_cast<const int32_t*>(reinterpret_cast<const unsigned char*>(st) + s * TB);
          nxt = fetch_object(p, slots[tid]);
        }
        if (SIG)
          obj = cell_b2c20<HAS_GRAD>(SigIn<PlanarIn<E>>{P}, TL, SigOut<PlanarOut<E>>{G}, p, sums);
        else
          obj = cell_b2c20<HAS_GRAD>(P, TL, G, p, sums);
      } else if (SIG) {
        obj = cell_b2c20<HAS_GRAD>(SigIn<PlanarIn<E>>{P}, T, SigOut<PlanarOut<E>>{G}, p, sums);
      } else {
        obj = cell_b2c20<HAS_GRAD>(P, T, G, p, sums);
      }
      if (obj) note_object(m1, m2, ((int64_t)blockIdx.x + k * gridDim.x) * tile_cells + tid);
    }
    if (HAS_GRAD) {
      fence_async_smem();
      if (!INPLACE && tid == 0) bulk_wait_read<(NOUT >= 2 ? NOUT - 2 : 0)>();
    }
    __syncthreads();
    if (tid == 0) {
      if (HAS_GRAD) {
        bulk_s2g(gg + ((int64_t)blockIdx.x + k * gridDim.x) * tile_elems, gtile, GB, pol);
        bulk_commit();
      }
      if (k + STAGES < my_n) {
        if (HAS_GRAD && INPLACE) bulk_wait_read<0>();  // the store has drained stage s: it may be refilled
        issue(k + STAGES);
      }
    }
  }
  // ragged tail (< tile_imgs images): one CTA, strided global accesses
  const int64_t tail0 = full * tile_cells;
  if ((int64_t)blockIdx.x == full % gridDim.x && tail0 + tid < p.cells && tid < tile_cells) {
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

template <typename E, bool HAS_GRAD, int NOUT, bool SIG, bool LIST = false>
int launch_planar_n(const LossParams& p, int tile_imgs, cudaStream_t stream) {
  constexpr int STAGES = 2;
  const int tile_cells = tile_imgs * p.S * p.S;
  const size_t smem = (size_t)STAGES * tile_cells * (30 * sizeof(E) + (LIST ? 4 : 120)) +
                      (size_t)NOUT * tile_cells * 30 * sizeof(E) + STAGES * sizeof(uint64_t);
  const int threads = (tile_cells + 31) / 32 * 32;
  auto kern = loss_tma_planar_kernel<E, HAS_GRAD, STAGES, NOUT, SIG, LIST>;
  static KernelPrep prep;   // one per kernel instantiation: attribute / occupancy queries once per device
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

// This kernel (separate output buffers, two or more CTAs per SM) for small tiles; the warp-specialised kernel with the
// gradient written in place over the pred stage (loss_ws.cu) for large ones.
template <typename E, bool HAS_GRAD>
int launch_planar(const LossParams& p, int tile_imgs, cudaStream_t stream) {
  const size_t tile_cells = (size_t)tile_imgs * p.S * p.S;
  const size_t separate = 2 * tile_cells * (30 * sizeof(E) + (p.list_mode ? 4 : 120)) + 2 * tile_cells * 30 * sizeof(E);
  // measured (tools/tune_loss.py): whole-image tiles of a 14x14 grid (196 cells) run best warp-specialised with three
  // in-place stages (fp32 0.753 ms vs 0.80, bf16 0.494 ms vs 0.537 at config-3 size); 98-cell tiles (S = 7) run best
  // here with separate output buffers (0.717 ms vs 0.80)
  if (separate > 110 * 1024 || tile_cells >= 160)
    return launch_loss_ws(p, sizeof(E) == 2, HAS_GRAD, true, (int)tile_cells, 3, stream);
  if (p.list_mode)
    return p.logits ? launch_planar_n<E, HAS_GRAD, 2, true, true>(p, tile_imgs, stream)
                    : launch_planar_n<E, HAS_GRAD, 2, false, true>(p, tile_imgs, stream);
  return p.logits ? launch_planar_n<E, HAS_GRAD, 2, true, false>(p, tile_imgs, stream)
                  : launch_planar_n<E, HAS_GRAD, 2, false, false>(p, tile_imgs, stream);
}

}  // namespace

// whole images per tile so that both tiles are multiples of 16 bytes and hold about `target_cells` cells; 0 = no fit
int planar_tile_imgs(int S, size_t esz, int target_cells, bool list_mode) {
  const int SS = S * S;
  int m = 0;
  for (int k = 1; k <= 16; ++k)
    if (((size_t)k * SS * 30 * esz) % 16 == 0 && ((size_t)k * SS * 120) % 16 == 0 &&
        (!list_mode || ((size_t)k * SS * 4) % 16 == 0)) {
      m = k;
      break;
    }
  if (m == 0 || m * SS > 256) return 0;
  int t = m;
  while ((t + m) * SS <= target_cells) t += m;
  return t;
}


int launch_loss_planar(const LossParams& p, bool bf16, bool has_grad, int tile_imgs, cudaStream_t stream) {
  if (bf16) return has_grad ? launch_planar<__nv_bfloat16, true>(p, tile_imgs, stream)
                            : launch_planar<__nv_bfloat16, false>(p, tile_imgs, stream);
  return has_grad ? launch_planar<float, true>(p, tile_imgs, stream) : launch_planar<float, false>(p, tile_imgs, stream);
}

}  // namespace yolo1
