// loss.cu -- K1: YOLO v1 loss forward + backward in one fused pass (sm_100a).
//
// Replaces YOLOLossV1.forward (reference v1Loss.py:22-118) and the autograd backward that train.py:171
// runs over its graph.  Specification: SURVEY.md Appendix A.  Arithmetic fp32 (reference dtype).
//
// Data path of the fast kernel (contiguous [N,S,S,30] pred/target/grad, the layout the reference's
// DataLoader and benchmarks produce):
//   HBM --cp.async.bulk (1-D TMA) + mbarrier--> shared tile (TILE cells x 120 B, pred and target)
//   one thread per grid cell computes the terms and the 30 gradient values into a shared output tile
//   shared --cp.async.bulk store--> HBM.
// Every byte crosses HBM exactly once (360 B per cell: 120 pred + 120 target + 120 grad); no thread
// issues an uncoalesced global access.  CTAs are persistent (grid = resident CTAs) and walk tiles
// round-robin with a STAGES-deep input ring and NOUT output buffers.
//
// The reference's row-slice behaviour (v1Loss.py:101: the first two responsible boxes of the CALL get
// plain squared error on x,y,w,h, all later ones squared error of square roots) is an order dependent,
// global predicate.  The streaming pass therefore evaluates every object cell in square-root form and
// records the two smallest object-cell indices; the last CTA to finish (ticket) re-evaluates those two
// cells in plain form, patches their 4 coordinate gradients and the location sum, reduces the per-CTA
// partial sums in a fixed order (deterministic) and writes the five loss terms.  No host sync, no second
// launch.  Chunked calls (host-buffer pipeline) carry the number of objects seen so far in the workspace.
#include "loss_common.cuh"

namespace yolo1 {
namespace {

// ---- K1 host-resident variant: pred / target / grad are pinned, mapped HOST memory ---------------------------
// For callers that hold host buffers (yolo1_loss_fwd_bwd_host).  Shipping the tensors to HBM first costs
// 240 B per cell over PCIe; but a cell without object (94-98 % of them) only needs target[0] and pred[0:2].
// Here every thread pulls exactly those two 8-byte pieces straight from host memory (one 32-byte sector
// each over PCIe; object cells then read their full 240 B), builds the 120-byte gradient row in a shared tile
// and the tile leaves with one bulk store directly into the caller's host gradient buffer.  One launch, no
// staging buffers, PCIe reads and writes overlap inside the kernel.  Measured rates: tools/zc_probe.cu.
struct GlobIn2 {
  const float* p;
  __device__ __forceinline__ float2 ld2(int c) const { return *reinterpret_cast<const float2*>(p + c); }
};

template <bool HAS_GRAD, int TILE>
__global__ void __launch_bounds__(TILE) loss_hostmapped_kernel(const __grid_constant__ LossParams p) {
  constexpr int D = 30, NOUT = 2;
  constexpr uint32_t GB = TILE * D * sizeof(float);
  extern __shared__ __align__(128) unsigned char smem[];
  float* so = reinterpret_cast<float*>(smem);
  const int tid = threadIdx.x;
  const int64_t full = p.cells / TILE;
  const int64_t my_n = full > (int64_t)blockIdx.x ? (full - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const float* gp = reinterpret_cast<const float*>(p.pred);
  float* gg = reinterpret_cast<float*>(p.grad);
  uint64_t pol = 0;
  if (tid == 0) pol = policy_evict_first();
  CellSums sums = {0.f, 0.f, 0.f, 0.f};
  uint32_t m1 = 0, m2 = 0;
  for (int64_t k = 0; k < my_n; ++k) {
    const int64_t tile = (int64_t)blockIdx.x + k * gridDim.x, q = tile * TILE + tid;
    const int o = (int)(k % NOUT);
    const GlobIn2 P{gp + q * D};
    const GlobIn2 T{p.target + q * D};
    const SmemOutF32 G{so + o * (TILE * D) + tid * D};
    if (cell_b2c20<HAS_GRAD>(P, T, G, p, sums)) note_object(m1, m2, q);
    if (HAS_GRAD) {
      fence_async_smem();
      if (tid == 0) bulk_wait_read<NOUT - 2>();
      __syncthreads();
      if (tid == 0) {
        bulk_s2g(gg + tile * (TILE * D), so + o * (TILE * D), GB, pol);
        bulk_commit();
      }
    }
  }
  const int64_t tail0 = full * TILE;
  if ((int64_t)blockIdx.x == full % gridDim.x && tail0 + tid < p.cells) {
    const int64_t q = tail0 + tid;
    const GlobIn<float> P{gp + q * D, 1, false};
    const GlobIn<float> T{p.target + q * D, 1, false};
    const GlobOut<float> G{HAS_GRAD ? gg + q * D : nullptr, 1, nullptr, 0, false};
    if (cell_generic<HAS_GRAD, false>(P, T, G, p, sums)) note_object(m1, m2, q);
  }
  block_epilogue<float, HAS_GRAD, true>(sums, m1, m2, p);
}

// ---- K1 generic kernel: any strides (e.g. the backbone's permuted NCHW view), any B, C --------------
// One thread per cell, grid-stride.  With the channel-planar view consecutive lanes read consecutive
// addresses of one channel plane, so every access is coalesced; channels of cells without object are
// never read beyond the B confidences.
template <typename E, bool HAS_GRAD>
__global__ void __launch_bounds__(kGenericThreads) loss_generic_kernel(const __grid_constant__ LossParams p) {
  CellSums sums = {0.f, 0.f, 0.f, 0.f};
  uint32_t m1 = 0, m2 = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < p.cells; q += stride) {
    const E* zq = reinterpret_cast<const E*>(p.pred) + cell_offset(p.ps, q, p.S);
    const GlobIn<E> P{zq, p.ps[3], p.logits != 0};
    const GlobOut<E> G{HAS_GRAD ? reinterpret_cast<E*>(p.grad) + cell_offset(p.gs, q, p.S) : nullptr, p.gs[3], zq,
                       p.ps[3], p.logits != 0};
    bool obj;
    if (p.list_mode) {
      obj = cell_generic<HAS_GRAD, false>(P, list_targetS(p, q), G, p, sums);
    } else {
      const GlobIn<float> T{p.target + cell_offset(p.ts, q, p.S), p.ts[3], false};
      obj = cell_generic<HAS_GRAD, false>(P, T, G, p, sums);
    }
    if (obj) note_object(m1, m2, q);
  }
  block_epilogue<E, HAS_GRAD, false>(sums, m1, m2, p);  // no bulk stores in flight here
}

// object lists -> cell ownership map (yolo1_loss_fwd_bwd_objects): per cell the LAST object of its image that falls
// into it (the reference encoder resets the cell before each write, utils/YOLODataLoader.py:220); cells without
// object hold -1.  4 bytes per cell instead of a 120-byte dense target row.
__global__ void __launch_bounds__(256) object_cells_kernel(const float* __restrict__ boxes,
                                                           const int32_t* __restrict__ labels,
                                                           const int64_t* __restrict__ offsets, int64_t N, int S, int C,
                                                           float cs, int G, int32_t* __restrict__ cellobj,
                                                           int32_t* __restrict__ status) {
  // A CTA builds the map of G whole images in shared memory and writes it out with coalesced 16-byte stores: one
  // launch, every map byte written once -- no memset of the map beforehand, no scattered 4-byte global stores.
  // One thread per OBJECT of those images (round 1: one thread per image walking its list -- three dependent trips
  // to memory per object, 20 busy threads per CTA): the images' offsets go to shared memory first, then every object
  // is one independent load; its image is the last one whose offset is <= its index (binary search over G + 1
  // entries), and "the last object wins" is an atomicMax on the object index.
  extern __shared__ __align__(16) int32_t own[];
  const int SS = S * S;
  const int64_t g0 = (int64_t)blockIdx.x * G;
  const int n_img = (int)((N - g0 < G) ? N - g0 : G);
  const int total = n_img * SS;
  int64_t* offs = reinterpret_cast<int64_t*>(own + (size_t)G * SS);   // G * SS * 4 is a multiple of 16 (host)
  for (int t = threadIdx.x; t <= n_img; t += blockDim.x) offs[t] = __ldg(offsets + g0 + t);
  for (int t = threadIdx.x; t < total; t += blockDim.x) own[t] = -1;
  __syncthreads();
  for (int64_t k = offs[0] + threadIdx.x, hi = offs[n_img]; k < hi; k += blockDim.x) {
    const float4 box = __ldg(reinterpret_cast<const float4*>(boxes) + k);
    const int lab = __ldg(labels + k);
    int a = 0, b = n_img;   // offs[a] <= k < offs[b]
    while (b - a > 1) {
      const int m = (a + b) >> 1;
      if (offs[m] <= k) a = m; else b = m;
    }
    float fi, fj, d;
    encode_axis(box.x, cs, fi, d);
    encode_axis(box.y, cs, fj, d);
    int col = (int)fi, row = (int)fj;
    if (col < -S || col >= S || row < -S || row >= S || lab < -C || lab >= C) {  // reference: IndexError
      atomicExch(status, 1);
      continue;
    }
    if (col < 0) col += S;  // Python indexing
    if (row < 0) row += S;
    atomicMax(&own[a * SS + row * S + col], (int32_t)k);
  }
  __syncthreads();
  int32_t* dst = cellobj + g0 * SS;   // 16-byte aligned: G * S * S is a multiple of 4 (host)
  for (int t = threadIdx.x; t < (total >> 2); t += blockDim.x)
    reinterpret_cast<int4*>(dst)[t] = reinterpret_cast<const int4*>(own)[t];
  for (int t = (total & ~3) + threadIdx.x; t < total; t += blockDim.x) dst[t] = own[t];
}

// grad *= *scale (autograd's backward(grad_output)); returns untouched when the scalar is exactly 1
template <typename E>
__global__ void __launch_bounds__(256) scale_grad_kernel(E* g, int64_t n, const float* scale) {
  const float s = __ldg(scale);
  if (s == 1.0f) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    st_elem(g + i, ld_elem(g + i) * s);
}
__global__ void __launch_bounds__(256) scale_grad_f32x4_kernel(float4* g, int64_t n4, const float* scale) {
  const float s = __ldg(scale);
  if (s == 1.0f) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = g[i];
    v.x *= s, v.y *= s, v.z *= s, v.w *= s;
    g[i] = v;
  }
}

// ---- host side ---------------------------------------------------------------------------------------
bool contiguous(const int64_t st[4], int S, int D) {
  return st[3] == 1 && st[2] == D && st[1] == (int64_t)S * D && st[0] == (int64_t)S * S * D;
}

bool planar(const int64_t st[4], int S, int D) {
  return st[3] == (int64_t)S * S && st[2] == 1 && st[1] == S && st[0] == (int64_t)S * S * D;
}

// whole images per tile so that both tiles are multiples of 16 bytes and hold 128..256 cells; 0 = no fit
template <bool HAS_GRAD>
int launch_hostmapped(const LossParams& p, cudaStream_t stream) {
  constexpr int TILE = 128;
  constexpr size_t smem = 2 * TILE * 30 * sizeof(float);
  auto kern = loss_hostmapped_kernel<HAS_GRAD, TILE>;
  static KernelPrep prep;
  int sms = kNumSMs;
  if (int rc = prepare_kernel(prep, kern, TILE, smem, false, &sms, nullptr)) return rc;
  const int64_t tiles = p.cells / TILE;
  int64_t grid = (int64_t)sms * 6;   // 31 KB and 128 threads per CTA: plenty of PCIe requests in flight
  if (grid > tiles) grid = tiles;
  if (grid > kMaxGrid) grid = kMaxGrid;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, TILE, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

template <typename E, bool HAS_GRAD>
int launch_generic(const LossParams& p, cudaStream_t stream) {
  static KernelPrep prep;
  int sms = kNumSMs;
  if (int rc = prepare_kernel(prep, loss_generic_kernel<E, HAS_GRAD>, kGenericThreads, 0, false, &sms, nullptr)) return rc;
  int64_t grid = (p.cells + kGenericThreads - 1) / kGenericThreads;
  const int64_t cap = (int64_t)sms * 8;
  if (grid > cap) grid = cap;
  if (grid > kMaxGrid) grid = kMaxGrid;
  if (grid < 1) grid = 1;
  loss_generic_kernel<E, HAS_GRAD><<<(unsigned)grid, kGenericThreads, 0, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace

// Launches one chunk of a loss call.  chunk_flags: bit 0 = first chunk (resets the workspace), bit 1 = last
// chunk (writes terms), bit 2 = pred holds pre-sigmoid logits (fused head epilogue).  variant < 0 forces the generic kernel.  Used by the public entry points below and by
// the host-buffer pipeline (host_ctx.cu).
struct ObjectLists {  // non-null: targets come as object lists
  const int32_t* cellobj;   // ownership map prepared by object_cells_kernel, or null: the streaming kernel finds the
                            // owners itself (list_mode 2; contiguous fast path only) from offsets / status
  const float* boxes;
  const int32_t* labels;
  const int64_t* offsets;
  int32_t* status;
};
constexpr int kVariantOwnersInKernel = 60;   // objects entry point: force list_mode 2; 61 = force the pre-pass

int loss_launch_chunk_impl(const void* pred, const int64_t ps[4], int pred_dtype, const float* target,
                           const int64_t ts_in[4], const ObjectLists* lists, void* grad, const int64_t gs[4],
                           float* terms, int64_t N, int S, int B, int C, float lambda_coord, float lambda_noobj,
                           float inv_batch_size, int coord_mode, void* workspace, size_t workspace_bytes,
                           int chunk_flags, int variant, cudaStream_t stream) {
  const int64_t dense_st[4] = {(int64_t)S * S * (5 * B + C), (int64_t)S * (5 * B + C), 5 * B + C, 1};
  const int64_t* ts = lists ? dense_st : ts_in;
  if (!terms || !workspace || !ps || !ts || N < 0) return YOLO1_ERR_ARG;
  if (N > 0 && (!pred || (!target && !lists))) return YOLO1_ERR_ARG;  // an empty batch may come with null pointers
  if (grad && !gs) return YOLO1_ERR_ARG;
  if ( S <= 0 || B <= 0 || C < 0) return YOLO1_ERR_ARG;
  if (pred_dtype != YOLO1_DTYPE_F32 && pred_dtype != YOLO1_DTYPE_BF16) return YOLO1_ERR_ARG;
  if (coord_mode != YOLO1_COORD_REFERENCE && coord_mode != YOLO1_COORD_PAPER) return YOLO1_ERR_ARG;
  if (B > kMaxB || 5 * B + C > 128) return YOLO1_ERR_UNSUPPORTED;
  if (workspace_bytes < sizeof(LossWs)) return YOLO1_ERR_ARG;
  const int64_t cells = N * S * S;
  if (cells >= 0xFFFFFFFFll) return YOLO1_ERR_UNSUPPORTED;  // cell indices are tracked in 32 bits per launch
  const size_t esz = pred_dtype == YOLO1_DTYPE_F32 ? 4 : 2;
  if ((uintptr_t)pred % esz || (uintptr_t)target % 4 || (grad && (uintptr_t)grad % esz) || (uintptr_t)terms % 4 ||
      (uintptr_t)workspace % 8)
    return YOLO1_ERR_ALIGN;
  if (lists)   // a 16-byte aligned stand-in for the checks below
    target = lists->cellobj ? reinterpret_cast<const float*>(lists->cellobj) : lists->boxes;

  LossParams p;
  p.pred = pred, p.target = target, p.grad = grad, p.terms = terms, p.ws = reinterpret_cast<LossWs*>(workspace);
  for (int d = 0; d < 4; ++d) p.ps[d] = ps[d], p.ts[d] = ts[d], p.gs[d] = grad ? gs[d] : 0;
  p.cells = cells, p.S = S, p.B = B, p.C = C;
  p.Sf = (float)S, p.lc = lambda_coord, p.ln = lambda_noobj, p.inv_bs = inv_batch_size;
  p.k2ln = 2.0f * lambda_noobj * inv_batch_size, p.k2 = 2.0f * inv_batch_size;
  p.coord_mode = coord_mode, p.last_chunk = (chunk_flags & 2) ? 1 : 0, p.logits = (chunk_flags & 4) ? 1 : 0;
  p.list_mode = lists ? (lists->cellobj ? 1 : 2) : 0;
  p.cellobj = lists ? lists->cellobj : nullptr, p.boxes = lists ? lists->boxes : nullptr;
  p.labels = lists ? lists->labels : nullptr, p.cs = (float)(1.0 / (double)S);
  p.offsets = lists ? lists->offsets : nullptr, p.status = lists ? lists->status : nullptr;
  if (lists) p.target = nullptr;

  // Small single-chunk calls (train.py:38-41 trains with 12 x 14 x 14 = 2 352 cells; BASELINE config 1 is 1 568):
  // one cluster, no workspace to reset, no fix-up pass (loss_small.cu).
  const int D = 5 * B + C;
  const bool bf = pred_dtype == YOLO1_DTYPE_BF16;
  const bool aligned16 = (uintptr_t)pred % 16 == 0 && (uintptr_t)target % 16 == 0 && (!grad || (uintptr_t)grad % 16 == 0);
  const bool one_chunk = (chunk_flags & 3) == 3;
  if (one_chunk && cells > 0 && cells <= kSmallTryCells && (variant == 0 || variant == kVariantSmall)) {
    int layout = 0;
    if (!lists && B == 2 && C == 20 && aligned16 && contiguous(ts, S, D)) {
      if (contiguous(ps, S, D) && (!grad || contiguous(gs, S, D))) layout = 1;
      else if (planar(ps, S, D) && (!grad || planar(gs, S, D))) layout = 2;
    }
    const int rc = launch_loss_small(p, bf, grad != nullptr, layout, stream);
    if (rc != YOLO1_ERR_UNSUPPORTED || variant == kVariantSmall) return rc;   // too large: the streaming kernels
  } else if (variant == kVariantSmall) {
    return YOLO1_ERR_UNSUPPORTED;
  }
  if (variant == kVariantNoSmall || variant == kVariantSmall) variant = 0;

  if (chunk_flags & 1) YOLO1_CUDA_TRY(cudaMemsetAsync(workspace, 0, offsetof(LossWs, partial), stream));

  const bool fast = variant >= 0 && B == 2 && C == 20 && contiguous(ps, S, D) && contiguous(ts, S, D) &&
                    (!grad || contiguous(gs, S, D)) && aligned16;
  if (variant == kVariantHostMapped) {  // pointers are device-visible HOST memory (host_ctx.cu)
    if (!fast || bf || p.logits || p.list_mode) return YOLO1_ERR_UNSUPPORTED;
    return grad ? launch_hostmapped<true>(p, stream) : launch_hostmapped<false>(p, stream);
  }
  if (variant >= kVariantSparse && variant < kVariantSparse + 3) {
    if (!fast || bf || p.logits) return YOLO1_ERR_UNSUPPORTED;
    return launch_loss_sparse(p, grad != nullptr, variant - kVariantSparse, stream);
  }
  if (fast) {
    return launch_loss_nhwc(p, bf, grad != nullptr, variant, stream);
  }
  if (p.list_mode == 2) return YOLO1_ERR_UNSUPPORTED;   // only the kernel above searches the lists itself
  // contiguous tensors of any (B, C): the same bulk-copy pipeline with a runtime channel count
  const bool fast_any = variant >= 0 && !lists && contiguous(ps, S, D) && contiguous(ts, S, D) &&
                        (!grad || contiguous(gs, S, D)) && (uintptr_t)pred % 16 == 0 && (uintptr_t)target % 16 == 0 &&
                        (!grad || (uintptr_t)grad % 16 == 0) && ((size_t)D * esz) % 4 == 0 && D <= 48;
  // (measured, tools/tune_any_channels.py: D = 25 runs 1.5x faster than the strided kernel this way; from D ~ 90 on
  // a tile leaves room for one warp per CTA only and the two kernels tie, so wide heads keep the strided kernel)
  if (fast_any && variant != kVariantHostMapped) return launch_loss_nhwc_any(p, bf, grad != nullptr, stream);
  const int tile_imgs = planar_tile_imgs(S, esz, variant == 1 ? 224 : 128, lists != nullptr);
  const bool fast_planar = variant >= 0 && B == 2 && C == 20 && tile_imgs > 0 && planar(ps, S, D) &&
                           contiguous(ts, S, D) && (!grad || planar(gs, S, D)) && (uintptr_t)pred % 16 == 0 &&
                           (uintptr_t)target % 16 == 0 && (!grad || (uintptr_t)grad % 16 == 0);
  if (fast_planar) {
    if (variant == 20)  // force the warp-specialised kernel whatever the tile size
      return launch_loss_ws(p, bf, grad != nullptr, true, tile_imgs * S * S, 3, stream);
    // confidence-first form (loss_planar_sparse.cu): reads planes 0-1 of pred and gathers the rest for object cells
    // only -- 248 instead of 360 B / cell.  Its tiles wait for the lane that gathers, so it is latency- rather than
    // byte-bound; measured (profiles/tune_planar_r2.log, config-3 size) it wins where the dense kernels are weakest:
    // fp32, whole 14x14 images per tile, dense target (0.686 vs 0.736 ms; fused head 0.731 vs 0.809), and loses for
    // bf16, object lists and 7x7 grids.  50 forces it, 51 forces the dense planar kernels (A/B runs).
    const int sparse_imgs = planar_tile_imgs(S, esz, 224, lists != nullptr);
    if (variant == kVariantPlanarSparse || (variant != 51 && !bf && !lists && sparse_imgs * S * S >= 160)) {
      const int rc = launch_loss_planar_sparse(p, bf, grad != nullptr, sparse_imgs, stream);
      if (rc != YOLO1_ERR_UNSUPPORTED || variant == kVariantPlanarSparse) return rc;
    }
    return launch_loss_planar(p, bf, grad != nullptr, tile_imgs, stream);
  }
  if (bf) return grad ? launch_generic<__nv_bfloat16, true>(p, stream) : launch_generic<__nv_bfloat16, false>(p, stream);
  return grad ? launch_generic<float, true>(p, stream) : launch_generic<float, false>(p, stream);
}

int loss_launch_chunk(const void* pred, const int64_t ps[4], int pred_dtype, const float* target,
                      const int64_t ts[4], void* grad, const int64_t gs[4], float* terms, int64_t N, int S,
                      int B, int C, float lambda_coord, float lambda_noobj, float inv_batch_size, int coord_mode,
                      void* workspace, size_t workspace_bytes, int chunk_flags, int variant, cudaStream_t stream) {
  return loss_launch_chunk_impl(pred, ps, pred_dtype, target, ts, nullptr, grad, gs, terms, N, S, B, C, lambda_coord,
                                lambda_noobj, inv_batch_size, coord_mode, workspace, workspace_bytes, chunk_flags,
                                variant, stream);
}

}  // namespace yolo1

extern "C" {

size_t yolo1_loss_objects_workspace_bytes(int64_t N, int S, int, int) {
  const size_t cells = (size_t)(N < 0 ? 0 : N) * S * S;
  return sizeof(yolo1::LossWs) + ((cells * sizeof(int32_t) + 15) & ~(size_t)15) + 16;
}

int yolo1_loss_fwd_bwd_objects(const void* pred, const int64_t pred_strides[4], int pred_dtype, int from_logits,
                               const float* boxes, const int32_t* labels, const int64_t* offsets, void* grad,
                               const int64_t grad_strides[4], float* terms, int64_t N, int S, int B, int C,
                               float lambda_coord, float lambda_noobj, float inv_batch_size, int coord_mode,
                               void* workspace, size_t workspace_bytes, int32_t* status, void* stream) {
  return yolo1_loss_fwd_bwd_objects_ex(pred, pred_strides, pred_dtype, from_logits, boxes, labels, offsets, grad,
                                       grad_strides, terms, N, S, B, C, lambda_coord, lambda_noobj, inv_batch_size,
                                       coord_mode, workspace, workspace_bytes, status, 0, stream);
}

int yolo1_loss_fwd_bwd_objects_ex(const void* pred, const int64_t pred_strides[4], int pred_dtype, int from_logits,
                                  const float* boxes, const int32_t* labels, const int64_t* offsets, void* grad,
                                  const int64_t grad_strides[4], float* terms, int64_t N, int S, int B, int C,
                                  float lambda_coord, float lambda_noobj, float inv_batch_size, int coord_mode,
                                  void* workspace, size_t workspace_bytes, int32_t* status, int variant,
                                  void* stream) {
  using namespace yolo1;
  if (N < 0 || S <= 0 || B <= 0 || C <= 0 || !workspace || !status || (N > 0 && !offsets)) return YOLO1_ERR_ARG;
  if (workspace_bytes < yolo1_loss_objects_workspace_bytes(N, S, B, C)) return YOLO1_ERR_ARG;
  if ((uintptr_t)workspace % 16 || (uintptr_t)boxes % 16 || (uintptr_t)labels % 4 || (uintptr_t)offsets % 8 ||
      (uintptr_t)status % 4)
    return YOLO1_ERR_ALIGN;
  static_assert(sizeof(LossWs) % 16 == 0, "the cell map follows the header 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  int32_t* cellobj = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(workspace) + sizeof(LossWs));
  const int64_t cells = N * S * S;
  YOLO1_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
  // Calls that will run the contiguous streaming kernel (same test as loss_launch_chunk_impl's `fast`, beyond the
  // small-call sizes) skip the pre-pass and its map: warp 0 of every CTA finds the owners of the tiles ahead.
  const int D = 5 * B + C;
  // fp32 only: the helper warp's registers cost the bf16 kernel three of its seven resident CTAs per SM (measured:
  // 0.419 against 0.333 ms at config-3 size), while fp32 tiles are limited by shared memory either way.
  const bool in_kernel = (variant == 0 || variant == kVariantOwnersInKernel) && B == 2 && C == 20 && S >= 3 &&
                         (pred_dtype == YOLO1_DTYPE_F32 || variant == kVariantOwnersInKernel) &&
                         cells > kSmallTryCells && pred_strides && contiguous(pred_strides, S, D) &&
                         (!grad || (grad_strides && contiguous(grad_strides, S, D))) && (uintptr_t)pred % 16 == 0 &&
                         (!grad || (uintptr_t)grad % 16 == 0);
  if (variant == kVariantOwnersInKernel && !in_kernel) return YOLO1_ERR_UNSUPPORTED;
  if (variant == kVariantOwnersInKernel || variant == kVariantOwnersInKernel + 1) variant = 0;
  if (in_kernel) {
    const ObjectLists lists = {nullptr, boxes, labels, offsets, status};
    return loss_launch_chunk_impl(pred, pred_strides, pred_dtype, nullptr, nullptr, &lists, grad, grad_strides, terms, N,
                                  S, B, C, lambda_coord, lambda_noobj, inv_batch_size, coord_mode, workspace,
                                  workspace_bytes, 3 | (from_logits ? 4 : 0), variant, s);
  }
  if (cells > 0) {
    int G = 4096 / (S * S);   // ~16 KB of map per CTA
    G = G < 4 ? 4 : (G > 256 ? 256 : (G & ~3));   // a multiple of 4 images keeps every CTA's slice 16-byte aligned
    const size_t smem = (size_t)G * S * S * sizeof(int32_t) + (size_t)(G + 1) * sizeof(int64_t);
    if (smem > 200 * 1024) return YOLO1_ERR_UNSUPPORTED;
    static KernelPrep prep;
    if (int rc = prepare_kernel(prep, object_cells_kernel, 256, smem, false, nullptr, nullptr)) return rc;
    object_cells_kernel<<<(unsigned)((N + G - 1) / G), 256, smem, s>>>(boxes, labels, offsets, N, S, C,
                                                                      (float)(1.0 / (double)S), G, cellobj, status);
    YOLO1_CUDA_TRY(cudaGetLastError());
  }
  const ObjectLists lists = {cellobj, boxes, labels, offsets, status};
  return loss_launch_chunk_impl(pred, pred_strides, pred_dtype, nullptr, nullptr, &lists, grad, grad_strides, terms, N, S,
                                B, C, lambda_coord, lambda_noobj, inv_batch_size, coord_mode, workspace,
                                workspace_bytes, 3 | (from_logits ? 4 : 0), variant, s);
}


size_t yolo1_loss_workspace_bytes(int64_t, int, int, int) { return sizeof(yolo1::LossWs); }

int yolo1_loss_fwd_bwd(const void* pred, const int64_t pred_strides[4], int pred_dtype, const float* target,
                       const int64_t target_strides[4], void* grad, const int64_t grad_strides[4], float* terms,
                       int64_t N, int S, int B, int C, float lambda_coord, float lambda_noobj,
                       float inv_batch_size, int coord_mode, void* workspace, size_t workspace_bytes,
                       void* stream) {
  return yolo1::loss_launch_chunk(pred, pred_strides, pred_dtype, target, target_strides, grad, grad_strides,
                                  terms, N, S, B, C, lambda_coord, lambda_noobj, inv_batch_size, coord_mode,
                                  workspace, workspace_bytes, 3, 0, (cudaStream_t)stream);
}

int yolo1_loss_fwd_bwd_ex(const void* pred, const int64_t pred_strides[4], int pred_dtype, const float* target,
                          const int64_t target_strides[4], void* grad, const int64_t grad_strides[4],
                          float* terms, int64_t N, int S, int B, int C, float lambda_coord, float lambda_noobj,
                          float inv_batch_size, int coord_mode, void* workspace, size_t workspace_bytes,
                          int variant, void* stream) {
  return yolo1::loss_launch_chunk(pred, pred_strides, pred_dtype, target, target_strides, grad, grad_strides,
                                  terms, N, S, B, C, lambda_coord, lambda_noobj, inv_batch_size, coord_mode,
                                  workspace, workspace_bytes, 3, variant, (cudaStream_t)stream);
}

int yolo1_loss_fwd_bwd_logits(const void* logits, const int64_t logit_strides[4], int dtype, const float* target,
                              const int64_t target_strides[4], void* grad, const int64_t grad_strides[4],
                              float* terms, int64_t N, int S, int B, int C, float lambda_coord, float lambda_noobj,
                              float inv_batch_size, int coord_mode, void* workspace, size_t workspace_bytes,
                              void* stream) {
  return yolo1::loss_launch_chunk(logits, logit_strides, dtype, target, target_strides, grad, grad_strides, terms, N,
                                  S, B, C, lambda_coord, lambda_noobj, inv_batch_size, coord_mode, workspace,
                                  workspace_bytes, 3 | 4, 0, (cudaStream_t)stream);
}

int yolo1_scale_grad(void* grad, int dtype, int64_t storage_numel, const float* grad_out_dev, void* stream) {
  if (!grad || !grad_out_dev || storage_numel < 0) return YOLO1_ERR_ARG;
  if (dtype != YOLO1_DTYPE_F32 && dtype != YOLO1_DTYPE_BF16) return YOLO1_ERR_ARG;
  if (storage_numel == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = yolo1::kNumSMs * 8;
  if (dtype == YOLO1_DTYPE_F32) {
    if ((uintptr_t)grad % 16 == 0 && storage_numel % 4 == 0)
      yolo1::scale_grad_f32x4_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<float4*>(grad), storage_numel / 4,
                                                          grad_out_dev);
    else
      yolo1::scale_grad_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<float*>(grad), storage_numel,
                                                           grad_out_dev);
  } else {
    yolo1::scale_grad_kernel<__nv_bfloat16>
        <<<grid, 256, 0, s>>>(reinterpret_cast<__nv_bfloat16*>(grad), storage_numel, grad_out_dev);
  }
  return (int)cudaGetLastError();
}

}  // extern "C"
