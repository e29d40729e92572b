// encode.cu -- target encoder on the GPU.
//
// Replaces yoloDataset.encoder (reference utils/YOLODataLoader.py:200-230) for a whole batch: ragged lists of
// (cx, cy, w, h) boxes and labels (CSR offsets) -> the dense target tensor [N,S,S,5B+C] the loss consumes.
// The reference builds a 94-98 % zero tensor on the host and ships it over PCIe; here only the object lists
// cross (a few dozen bytes per image) and the tensor is produced at HBM write speed.  Built with -fmad=false:
// the cell index and the in-cell offsets are bit-exact with the reference's fp32 arithmetic.
//
// A CTA owns a group of G whole images: the tile is zeroed in shared memory, warp i scatters image i's objects
// in input order, lanes <-> channels (the reference resets the cell before writing, so the LAST object in a cell
// wins), and the
// finished tile leaves with one bulk store (cp.async.bulk, SASS UBLKCP) -- a write-only stream of 120 B per cell.
#include "common.cuh"

namespace yolo1 {
namespace {

struct EncodeParams {
  const float* boxes;
  const int32_t* labels;
  const int64_t* offsets;
  float* target;
  int32_t* status;  // device int: set to 1 when a centre or label falls outside the grid / class range
  int64_t N;
  int S, B, C, G;
  float cs;  // fl32(1/S)
};

__global__ void __launch_bounds__(256) encode_kernel(const __grid_constant__ EncodeParams p) {
  extern __shared__ __align__(128) unsigned char raw[];
  float* tile = reinterpret_cast<float*>(raw);
  const int S = p.S, B = p.B, C = p.C, D = 5 * B + C;
  const int img = S * S * D;
  uint64_t pol = 0;
  if (threadIdx.x == 0) pol = policy_evict_first();
  for (int64_t g0 = (int64_t)blockIdx.x * p.G; g0 < p.N; g0 += (int64_t)gridDim.x * p.G) {
    const int n_img = (int)((p.N - g0 < p.G) ? p.N - g0 : p.G);
    const int total = n_img * img;
    for (int t = threadIdx.x; t < (total >> 2); t += blockDim.x)
      reinterpret_cast<float4*>(tile)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = (total & ~3) + threadIdx.x; t < total; t += blockDim.x) tile[t] = 0.f;
    __syncthreads();
    // Scatter: one WARP per image, lanes <-> channels of the object's cell.  (Round 1: one thread per image wrote the
    // 30 values of every object one after the other -- ~100 dependent instructions per object on a single lane while
    // the rest of the CTA waited at the barrier: 6.2 TB/s with no objects, 5.9 with three per image, 3.9 with six;
    // now 6.2 / 6.1 / 4.5.)  Objects of an image are still taken in input order, so the last one in a cell wins (:220).
    {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
      for (int i = warp; i < n_img; i += nwarps) {
        float* timg = tile + i * img;
        const int64_t k0 = p.offsets[g0 + i], k1 = p.offsets[g0 + i + 1];
        for (int64_t k = k0; k < k1; ++k) {
          // every lane reads the same record (one broadcast transaction each)
          const float cx = p.boxes[4 * k], cy = p.boxes[4 * k + 1], w = p.boxes[4 * k + 2], h = p.boxes[4 * k + 3];
          float fi, fj, dx, dy;
          encode_axis(cx, p.cs, fi, dx);  // :218-219, :223-224
          encode_axis(cy, p.cs, fj, dy);
          int col = (int)fi, row = (int)fj, lab = p.labels[k];
          if (col < -S || col >= S || row < -S || row >= S || lab < -C || lab >= C) {  // reference: IndexError
            if (lane == 0) atomicExch(p.status, 1);
            continue;
          }
          if (col < 0) col += S;  // Python indexing: -1 is the last row / column
          if (row < 0) row += S;
          if (lab < 0) lab += C;
          float* t = timg + (row * S + col) * D;
          for (int c = lane; c < D; c += 32) {
            // :220 reset, :221 confidences, :225-227 the same (dx, dy, w, h) in every box slot, :222 class one-hot
            float v = 0.f;
            if (c < B) {
              v = 1.f;
            } else if (c < 5 * B) {
              const int q = (c - B) & 3;
              v = q == 0 ? dx : (q == 1 ? dy : (q == 2 ? w : h));
            } else if (c - 5 * B == lab) {
              v = 1.f;
            }
            t[c] = v;
          }
          __syncwarp();   // the next object of this image may hit the same cell: keep the order
        }
      }
    }
    float* dst = p.target + g0 * img;
    const bool bulk = (((size_t)total * 4) % 16 == 0) && ((uintptr_t)dst % 16 == 0);
    if (bulk) {
      fence_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        bulk_s2g(dst, tile, (uint32_t)total * 4u, pol);
        bulk_commit();
        bulk_wait_read<0>();  // the tile is zeroed again right after
      }
      __syncthreads();
    } else {
      __syncthreads();
      for (int t = threadIdx.x; t < total; t += blockDim.x) dst[t] = tile[t];
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) bulk_wait_all<0>();
}

}  // namespace
}  // namespace yolo1

extern "C" YOLO1_API int yolo1_encode_targets(const float* boxes, const int32_t* labels, const int64_t* offsets, int64_t N, int S,
                                    int B, int C, float* target, int32_t* status, void* stream) {
  using namespace yolo1;
  if (N < 0 || S <= 0 || B <= 0 || C <= 0) return YOLO1_ERR_ARG;
  if (N == 0) return 0;
  if (!offsets || !target || !status) return YOLO1_ERR_ARG;  // boxes / labels may be null when there is no object
  if (5 * B + C > 128 || (int64_t)S * S * (5 * B + C) * 4 > 96 * 1024) return YOLO1_ERR_UNSUPPORTED;
  if ((uintptr_t)boxes % 4 || (uintptr_t)labels % 4 || (uintptr_t)offsets % 8 || (uintptr_t)target % 4 ||
      (uintptr_t)status % 4)
    return YOLO1_ERR_ALIGN;
  EncodeParams p;
  p.boxes = boxes, p.labels = labels, p.offsets = offsets, p.target = target, p.status = status;
  p.N = N, p.S = S, p.B = B, p.C = C;
  p.cs = (float)(1.0 / (double)S);
  const size_t img_bytes = (size_t)S * S * (5 * B + C) * 4;
  // ~24 KB tiles, 8 CTAs per SM: more, smaller tiles in flight hide the scatter's serial latency (one thread per
  // image, dependent loads) better (measured, GB/s written: 48 KB x 4 CTAs 4163 (S=14) / 3668 (S=7); 24 KB x 8:
  // 5156 / 4164; 96 KB x 2: 3287; a pure fill reaches 7290).
  int G = (int)((24 * 1024) / img_bytes);
  if (G < 1) G = 1;
  if (G > 256) G = 256;
  if (G > 1 && (G & 1)) G -= 1;  // an even image count keeps every tile a multiple of 16 bytes
  p.G = G;
  const size_t smem = (size_t)G * img_bytes;
  static yolo1::KernelPrep prep;
  int sms = kNumSMs, per_sm = 1;
  if (int rc = yolo1::prepare_kernel(prep, encode_kernel, 256, smem, true, &sms, &per_sm)) return rc;
  if (per_sm > 8) per_sm = 8;
  int64_t grid = (N + G - 1) / G;
  if (grid > (int64_t)sms * per_sm) grid = (int64_t)sms * per_sm;   // every CTA resident: the loop is grid-strided
  YOLO1_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), (cudaStream_t)stream));
  encode_kernel<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}
