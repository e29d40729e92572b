// host_ctx.cu -- host-buffer entry points: the same kernels for callers that hold HOST memory.
//
// The reference's call sites hand the loss / decoder CPU tensors when run with device='cpu'
// (v1Loss.py:10 `_device`, utils/utils.py:94 `device='cpu'`).  A context owns three sets of device staging
// buffers and three streams (H2D, compute, D2H); a batch is cut into chunks of `chunk_images` images and
// chunk c+1's upload, chunk c's kernel and chunk c-1's download overlap (PCIe is full duplex, the copy
// engines are independent of the SMs).  The calls block until the results are in the caller's buffers.
// The loss call's "first two objects of the call" rule (v1Loss.py:101) spans chunks through the carry
// counter in the loss workspace (loss.cu), so a chunked call equals one un-chunked call.
#include <new>

#include "common.cuh"

namespace yolo1 {
int loss_launch_chunk(const void* pred, const int64_t ps[4], int pred_dtype, const float* target,
                      const int64_t ts[4], void* grad, const int64_t gs[4], float* terms, int64_t N, int S,
                      int B, int C, float lambda_coord, float lambda_noobj, float inv_batch_size, int coord_mode,
                      void* workspace, size_t workspace_bytes, int chunk_flags, int variant, cudaStream_t stream);
}

// Inside the chunk loops an error must not return straight away: copies and kernels of earlier chunks are still
// in flight on the caller's buffers.  Record the code, leave the loop, and let the common exit drain the streams.
#define HOST_TRY_BREAK(expr)          \
  {                                   \
    cudaError_t _e = (expr);          \
    if (_e != cudaSuccess) {          \
      rc = (int)_e;                   \
      break;                          \
    }                                 \
  }

namespace {
constexpr int kBuf = 3;
constexpr int kVariantHostMapped = 100;

// Device-visible alias of a pinned (cudaHostAlloc / cudaHostRegister) host pointer, or nullptr for pageable
// memory.  No allocation, no copy: a query of the driver's address map.
template <typename T>
T* mapped_alias(T* host) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return nullptr;
  return reinterpret_cast<T*>(a.devicePointer);
}
}

struct yolo1_host_ctx {
  int device, S, B, C, D, max_n;
  int64_t chunk;  // images per chunk
  cudaStream_t s_in, s_k, s_out;
  cudaEvent_t in_done[kBuf], k_done[kBuf], out_done[kBuf];
  float *d_pred[kBuf], *d_tgt[kBuf], *d_grad[kBuf];
  float *d_boxes[kBuf], *d_scores[kBuf];
  int32_t *d_cls[kBuf], *d_counts[kBuf];
  void* d_ws;
  size_t ws_bytes;
  float* d_terms;
  float* h_terms;  // pinned
  int zero_copy;   // 0 staged pipeline; 1 target by copy engine + pred/grad in place; 2 everything in place
};

namespace {

void free_ctx(yolo1_host_ctx* c) {
  if (!c) return;
  for (int b = 0; b < kBuf; ++b) {
    cudaFree(c->d_pred[b]), cudaFree(c->d_tgt[b]), cudaFree(c->d_grad[b]);
    cudaFree(c->d_boxes[b]), cudaFree(c->d_scores[b]), cudaFree(c->d_cls[b]), cudaFree(c->d_counts[b]);
    if (c->in_done[b]) cudaEventDestroy(c->in_done[b]);
    if (c->k_done[b]) cudaEventDestroy(c->k_done[b]);
    if (c->out_done[b]) cudaEventDestroy(c->out_done[b]);
  }
  cudaFree(c->d_ws), cudaFree(c->d_terms);
  if (c->h_terms) cudaFreeHost(c->h_terms);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_k) cudaStreamDestroy(c->s_k);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  delete c;
}

// decode outputs are allocated on first use (a loss-only caller never pays for them)
int ensure_decode_buffers(yolo1_host_ctx* c) {
  if (c->d_boxes[0]) return 0;
  for (int b = 0; b < kBuf; ++b) {
    YOLO1_CUDA_TRY(cudaMalloc(&c->d_boxes[b], (size_t)c->chunk * c->max_n * 16));
    YOLO1_CUDA_TRY(cudaMalloc(&c->d_scores[b], (size_t)c->chunk * c->max_n * 4));
    YOLO1_CUDA_TRY(cudaMalloc(&c->d_cls[b], (size_t)c->chunk * c->max_n * 4));
    YOLO1_CUDA_TRY(cudaMalloc(&c->d_counts[b], (size_t)c->chunk * 4));
  }
  return 0;
}
// Images per chunk for a call of N images: the context's chunk size is the maximum (its staging buffers), but a
// batch that would fit one or two such chunks is cut finer (>= 8 chunks of >= ~2 MB) so that upload, kernel and
// download still overlap -- a single chunk would run H2D, kernel and D2H back to back.
int64_t call_chunk(const yolo1_host_ctx* c, int64_t N) {
  const int64_t img_bytes = (int64_t)c->S * c->S * c->D * 4;
  int64_t eff = (N + 7) / 8;
  const int64_t floor_imgs = (2 << 20) / img_bytes > 0 ? (2 << 20) / img_bytes : 1;
  if (eff < floor_imgs) eff = floor_imgs;
  if ((img_bytes % 16) != 0 && (eff & 1)) eff += 1;   // keep chunk boundaries 16-byte aligned
  if (eff > c->chunk) eff = c->chunk;
  return eff;
}

int ensure_loss_buffers(yolo1_host_ctx* c) {
  if (c->d_tgt[0]) return 0;
  const size_t bytes = (size_t)c->chunk * c->S * c->S * c->D * 4;
  for (int b = 0; b < kBuf; ++b) {
    YOLO1_CUDA_TRY(cudaMalloc(&c->d_tgt[b], bytes));
    YOLO1_CUDA_TRY(cudaMalloc(&c->d_grad[b], bytes));
  }
  return 0;
}

}  // namespace

extern "C" {

int yolo1_host_ctx_create(yolo1_host_ctx** out, int device, int S, int B, int C, int64_t chunk_images) {
  if (!out || S <= 0 || B <= 0 || C <= 0 || chunk_images < 0) return YOLO1_ERR_ARG;
  if (B > 8 || 5 * B + C > 128 || (int64_t)S * S * B > 1024) return YOLO1_ERR_UNSUPPORTED;
  *out = nullptr;
  YOLO1_CUDA_TRY(cudaSetDevice(device));
  yolo1_host_ctx* c = new (std::nothrow) yolo1_host_ctx();
  if (!c) return (int)cudaErrorMemoryAllocation;
  c->zero_copy = 2;
  c->device = device, c->S = S, c->B = B, c->C = C, c->D = 5 * B + C, c->max_n = S * S * B;
  const size_t img_bytes = (size_t)S * S * c->D * 4;
  // default: ~24 MB per tensor per chunk -- long enough for PCIe to reach its plateau, short enough that the
  // pipeline fill/drain (one chunk each) is a small part of a large batch
  c->chunk = chunk_images > 0 ? chunk_images : (int64_t)((24u << 20) / img_bytes > 0 ? (24u << 20) / img_bytes : 1);
  // chunk boundaries stay 16-byte aligned in the caller's buffers (bulk copies): an image of S*S*D floats is a
  // multiple of 8 bytes, so an even image count per chunk is enough
  if ((img_bytes % 16) != 0 && (c->chunk & 1)) c->chunk += 1;
  int rc = 0;
  auto fail = [&](cudaError_t e) {
    if (e != cudaSuccess && rc == 0) rc = (int)e;
  };
  fail(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
  fail(cudaStreamCreateWithFlags(&c->s_k, cudaStreamNonBlocking));
  fail(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
  for (int b = 0; b < kBuf && rc == 0; ++b) {
    fail(cudaEventCreateWithFlags(&c->in_done[b], cudaEventDisableTiming));
    fail(cudaEventCreateWithFlags(&c->k_done[b], cudaEventDisableTiming));
    fail(cudaEventCreateWithFlags(&c->out_done[b], cudaEventDisableTiming));
    fail(cudaMalloc(&c->d_pred[b], (size_t)c->chunk * img_bytes));
  }
  c->ws_bytes = yolo1_loss_workspace_bytes(0, S, B, C);
  fail(cudaMalloc(&c->d_ws, c->ws_bytes));
  fail(cudaMalloc(&c->d_terms, 5 * sizeof(float)));
  fail(cudaMallocHost(&c->h_terms, 5 * sizeof(float)));
  if (rc) {
    free_ctx(c);
    return rc;
  }
  *out = c;
  return 0;
}

void yolo1_host_ctx_destroy(yolo1_host_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  free_ctx(c);
}

int yolo1_host_ctx_set_zero_copy(yolo1_host_ctx* c, int enable) {
  if (!c) return YOLO1_ERR_ARG;
  c->zero_copy = enable < 0 ? 0 : (enable > 4 ? 4 : enable);
  return 0;
}

int yolo1_host_pin(void* ptr, size_t bytes) {
  if (!ptr || bytes == 0) return YOLO1_ERR_ARG;
  return (int)cudaHostRegister(ptr, bytes, cudaHostRegisterDefault);
}
int yolo1_host_unpin(void* ptr) {
  if (!ptr) return YOLO1_ERR_ARG;
  return (int)cudaHostUnregister(ptr);
}

int yolo1_loss_fwd_bwd_host(yolo1_host_ctx* c, const float* pred, const float* target, float* grad,
                            float terms[5], int64_t N, float lambda_coord, float lambda_noobj,
                            float inv_batch_size, int coord_mode) {
  if (!c || !pred || !target || !terms || N < 0) return YOLO1_ERR_ARG;
  if (coord_mode != YOLO1_COORD_REFERENCE && coord_mode != YOLO1_COORD_PAPER) return YOLO1_ERR_ARG;
  YOLO1_CUDA_TRY(cudaSetDevice(c->device));
  const int S = c->S, D = c->D;
  const int64_t img = (int64_t)S * S * D;
  const int64_t st[4] = {img, (int64_t)S * D, D, 1};
  const int64_t chunk = call_chunk(c, N);
  // Pinned + mapped buffers: one kernel reads the bytes it needs straight from host memory and stores the
  // gradient straight back (loss_hostmapped_kernel) -- no staging, far fewer bytes over PCIe.
  if ((c->zero_copy == 1 || c->zero_copy == 2) && c->B == 2 && c->C == 20 && N > 0 &&
      N * (int64_t)S * S < 0xFFFFFFFFll) {
    const float* dp = mapped_alias(pred);
    const float* dt = mapped_alias(target);
    float* dg = grad ? mapped_alias(grad) : nullptr;
    if (dp && dt && (!grad || dg) && (uintptr_t)dp % 16 == 0 && (uintptr_t)dt % 16 == 0 && (uintptr_t)dg % 16 == 0) {
      int rc = 0;
      if (c->zero_copy == 1) {
        // balanced use of the link: the copy engine streams the (dense) target chunk by chunk into HBM with
        // large read requests, while the SMs pull only pred's confidences (one sector per cell) from host
        // memory and bulk-store the gradient into the host buffer.  Chunks keep the `[:2]` carry (loss.cu).
        rc = ensure_loss_buffers(c);
        if (rc) return rc;
        const int64_t nchunks = (N + chunk - 1) / chunk;
        for (int64_t k = 0; k < nchunks && rc == 0; ++k) {
          const int b = (int)(k % kBuf);
          const int64_t n0 = k * chunk, n = (N - n0 < chunk) ? N - n0 : chunk;
          HOST_TRY_BREAK(cudaStreamWaitEvent(c->s_in, c->k_done[b], 0));
          HOST_TRY_BREAK(cudaMemcpyAsync(c->d_tgt[b], target + n0 * img, (size_t)n * img * 4, cudaMemcpyHostToDevice,
                                         c->s_in));
          HOST_TRY_BREAK(cudaEventRecord(c->in_done[b], c->s_in));
          HOST_TRY_BREAK(cudaStreamWaitEvent(c->s_k, c->in_done[b], 0));
          const int flags = (k == 0 ? 1 : 0) | (k == nchunks - 1 ? 2 : 0);
          rc = yolo1::loss_launch_chunk(dp + n0 * img, st, YOLO1_DTYPE_F32, c->d_tgt[b], st, dg ? dg + n0 * img : nullptr,
                                        st, c->d_terms, n, S, c->B, c->C, lambda_coord, lambda_noobj, inv_batch_size,
                                        coord_mode, c->d_ws, c->ws_bytes, flags, kVariantHostMapped, c->s_k);
          if (rc == 0) HOST_TRY_BREAK(cudaEventRecord(c->k_done[b], c->s_k));
        }
      } else {  // zero_copy == 2: everything through the SMs, one launch
        rc = yolo1::loss_launch_chunk(dp, st, YOLO1_DTYPE_F32, dt, st, dg, st, c->d_terms, N, S, c->B, c->C,
                                      lambda_coord, lambda_noobj, inv_batch_size, coord_mode, c->d_ws, c->ws_bytes,
                                      3, kVariantHostMapped, c->s_k);
      }
      if (rc == 0) {
        cudaError_t e = cudaMemcpyAsync(c->h_terms, c->d_terms, 5 * sizeof(float), cudaMemcpyDeviceToHost, c->s_k);
        if (e != cudaSuccess) rc = (int)e;
      }
      cudaError_t e1 = cudaStreamSynchronize(c->s_k), e2 = cudaStreamSynchronize(c->s_in);
      if (rc) return rc;
      if (e1 != cudaSuccess) return (int)e1;
      if (e2 != cudaSuccess) return (int)e2;
      for (int t = 0; t < 5; ++t) terms[t] = c->h_terms[t];
      return 0;
    }
  }
  int rc = ensure_loss_buffers(c);
  if (rc) return rc;
  // zero_copy == 3: only pred's confidences are pulled by the SMs from the caller's (pinned, mapped) buffer; the
  // dense target goes up and the gradient comes down through the copy engines as in the staged pipeline
  // zero_copy == 4: pred AND target pulled sector-wise by the SMs, only the gradient goes through a copy engine
  const float* pred_alias = nullptr;
  const float* tgt_alias = nullptr;
  if ((c->zero_copy == 3 || c->zero_copy == 4) && c->B == 2 && c->C == 20 && N > 0) {
    pred_alias = mapped_alias(pred);
    if ((uintptr_t)pred_alias % 16) pred_alias = nullptr;
    if (pred_alias && c->zero_copy == 4) {
      tgt_alias = mapped_alias(target);
      if ((uintptr_t)tgt_alias % 16) tgt_alias = nullptr;
    }
  }
  const int64_t nchunks = N == 0 ? 1 : (N + chunk - 1) / chunk;
  for (int64_t k = 0; k < nchunks; ++k) {
    const int b = (int)(k % kBuf);
    const int64_t n0 = k * chunk, n = (N - n0 < chunk) ? N - n0 : chunk;
    const size_t bytes = (size_t)n * img * 4;
    // upload: the kernel that last read these staging buffers must be done
    HOST_TRY_BREAK(cudaStreamWaitEvent(c->s_in, c->k_done[b], 0));
    if (bytes) {
      if (!pred_alias)
        HOST_TRY_BREAK(cudaMemcpyAsync(c->d_pred[b], pred + n0 * img, bytes, cudaMemcpyHostToDevice, c->s_in));
      if (!tgt_alias)
        HOST_TRY_BREAK(cudaMemcpyAsync(c->d_tgt[b], target + n0 * img, bytes, cudaMemcpyHostToDevice, c->s_in));
    }
    HOST_TRY_BREAK(cudaEventRecord(c->in_done[b], c->s_in));
    // compute: inputs uploaded, previous download of this gradient buffer done
    HOST_TRY_BREAK(cudaStreamWaitEvent(c->s_k, c->in_done[b], 0));
    HOST_TRY_BREAK(cudaStreamWaitEvent(c->s_k, c->out_done[b], 0));
    const int flags = (k == 0 ? 1 : 0) | (k == nchunks - 1 ? 2 : 0);
    rc = yolo1::loss_launch_chunk(pred_alias ? pred_alias + n0 * img : c->d_pred[b], st, YOLO1_DTYPE_F32,
                                  tgt_alias ? tgt_alias + n0 * img : c->d_tgt[b], st, grad ? c->d_grad[b] : nullptr, st, c->d_terms, n, S, c->B, c->C, lambda_coord,
                                  lambda_noobj, inv_batch_size, coord_mode, c->d_ws, c->ws_bytes, flags,
                                  pred_alias ? kVariantHostMapped : 0, c->s_k);
    if (rc) break;
    HOST_TRY_BREAK(cudaEventRecord(c->k_done[b], c->s_k));
    // download
    if (grad) {
      HOST_TRY_BREAK(cudaStreamWaitEvent(c->s_out, c->k_done[b], 0));
      if (bytes)
        HOST_TRY_BREAK(cudaMemcpyAsync(grad + n0 * img, c->d_grad[b], bytes, cudaMemcpyDeviceToHost, c->s_out));
      HOST_TRY_BREAK(cudaEventRecord(c->out_done[b], c->s_out));
    }
  }
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(c->h_terms, c->d_terms, 5 * sizeof(float), cudaMemcpyDeviceToHost, c->s_k);
    if (e != cudaSuccess) rc = (int)e;
  }
  cudaError_t e1 = cudaStreamSynchronize(c->s_k), e2 = cudaStreamSynchronize(c->s_out),
              e3 = cudaStreamSynchronize(c->s_in);
  if (rc) return rc;
  if (e1 != cudaSuccess) return (int)e1;
  if (e2 != cudaSuccess) return (int)e2;
  if (e3 != cudaSuccess) return (int)e3;
  for (int t = 0; t < 5; ++t) terms[t] = c->h_terms[t];
  return 0;
}

int yolo1_decode_nms_host(yolo1_host_ctx* c, const float* pred, int64_t N, double thresh, float iou_thr,
                          int per_class, float* out_boxes, float* out_scores, int32_t* out_cls,
                          int32_t* out_counts) {
  if (!c || !pred || !out_boxes || !out_scores || !out_cls || !out_counts || N < 0) return YOLO1_ERR_ARG;
  YOLO1_CUDA_TRY(cudaSetDevice(c->device));
  int rc = ensure_decode_buffers(c);
  if (rc) return rc;
  const int S = c->S, D = c->D, M = c->max_n;
  const int64_t img = (int64_t)S * S * D;
  const int64_t st[4] = {img, (int64_t)S * D, D, 1};
  const int64_t chunk = call_chunk(c, N);
  const int64_t nchunks = (N + chunk - 1) / chunk;
  for (int64_t k = 0; k < nchunks; ++k) {
    const int b = (int)(k % kBuf);
    const int64_t n0 = k * chunk, n = (N - n0 < chunk) ? N - n0 : chunk;
    HOST_TRY_BREAK(cudaStreamWaitEvent(c->s_in, c->k_done[b], 0));
    HOST_TRY_BREAK(cudaMemcpyAsync(c->d_pred[b], pred + n0 * img, (size_t)n * img * 4, cudaMemcpyHostToDevice,
                                   c->s_in));
    HOST_TRY_BREAK(cudaEventRecord(c->in_done[b], c->s_in));
    HOST_TRY_BREAK(cudaStreamWaitEvent(c->s_k, c->in_done[b], 0));
    HOST_TRY_BREAK(cudaStreamWaitEvent(c->s_k, c->out_done[b], 0));
    rc = yolo1_decode_nms(c->d_pred[b], st, YOLO1_DTYPE_F32, n, S, c->B, c->C, thresh, iou_thr, per_class,
                          c->d_boxes[b], c->d_scores[b], c->d_cls[b], c->d_counts[b], nullptr, nullptr, c->s_k);
    if (rc) break;
    HOST_TRY_BREAK(cudaEventRecord(c->k_done[b], c->s_k));
    HOST_TRY_BREAK(cudaStreamWaitEvent(c->s_out, c->k_done[b], 0));
    HOST_TRY_BREAK(cudaMemcpyAsync(out_boxes + n0 * M * 4, c->d_boxes[b], (size_t)n * M * 16,
                                   cudaMemcpyDeviceToHost, c->s_out));
    HOST_TRY_BREAK(cudaMemcpyAsync(out_scores + n0 * M, c->d_scores[b], (size_t)n * M * 4, cudaMemcpyDeviceToHost,
                                   c->s_out));
    HOST_TRY_BREAK(cudaMemcpyAsync(out_cls + n0 * M, c->d_cls[b], (size_t)n * M * 4, cudaMemcpyDeviceToHost,
                                   c->s_out));
    HOST_TRY_BREAK(cudaMemcpyAsync(out_counts + n0, c->d_counts[b], (size_t)n * 4, cudaMemcpyDeviceToHost,
                                   c->s_out));
    HOST_TRY_BREAK(cudaEventRecord(c->out_done[b], c->s_out));
  }
  cudaError_t e1 = cudaStreamSynchronize(c->s_k), e2 = cudaStreamSynchronize(c->s_out),
              e3 = cudaStreamSynchronize(c->s_in);
  if (rc) return rc;
  if (e1 != cudaSuccess) return (int)e1;
  if (e2 != cudaSuccess) return (int)e2;
  if (e3 != cudaSuccess) return (int)e3;
  return 0;
}

int yolo1_abi_version(void) { return YOLO1_ABI_VERSION; }

const char* yolo1_error_string(int rc) {
  if (rc == 0) return "success";
  if (rc == YOLO1_ERR_ARG) return "yolo1: invalid argument (null pointer, negative size or bad enum)";
  if (rc == YOLO1_ERR_UNSUPPORTED) return "yolo1: shape outside what the kernels are built for";
  if (rc == YOLO1_ERR_ALIGN) return "yolo1: pointer not aligned to its element size";
  if (rc < 0) return "yolo1: unknown library error";
  return cudaGetErrorString((cudaError_t)rc);
}

}  // extern "C"
