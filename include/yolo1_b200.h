/*
 * yolo1_b200.h -- C ABI of the B200-native YOLO v1 hot path (libyolo1_b200.so).
 *
 * The reference (haoran1062/YOLO_V1, pure Python on ATen ops) has no FFI layer of its own; its
 * boundary for this path is three Python call signatures.  Each entry point below names the reference
 * interface it replaces (file:line relative to /root/reference); the Python host side in
 * yolo_v1_b200/ re-creates those signatures on top of this ABI through ctypes, and INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; no torch / pybind types.  `stream` is a cudaStream_t passed as
 *     void* (NULL = the legacy default stream).
 *   - return 0 on success, < 0 for an invalid argument (no CUDA call was made), > 0 = the cudaError_t
 *     of the failing CUDA call.  No exceptions, no exit() (unlike utils/utils.py:30-32).
 *   - device entry points never allocate, never synchronise the host and keep no global state: the
 *     caller owns every buffer; calls are stream ordered and re-entrant.
 *   - strides are in ELEMENTS over (image n, row i, column j, channel c).  The backbone hands the loss
 *     a permuted NCHW view (backbones/OriginResNet.py:189): strides (D*S*S, S, 1, S*S).
 *   - dtype: 0 = float32, 1 = bfloat16 (pred / grad only; targets and outputs are float32).
 *   - channel layout of pred and target (v1Loss.py:24-25, utils/YOLODataLoader.py:220-227):
 *       [conf_0..conf_{B-1}, (x,y,w,h)_0..(x,y,w,h)_{B-1}, cls_0..cls_{C-1}],  D = 5B + C.
 */
#ifndef YOLO1_B200_H_
#define YOLO1_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define YOLO1_API __attribute__((visibility("default")))
#else
#define YOLO1_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define YOLO1_ABI_VERSION 1

#define YOLO1_DTYPE_F32 0
#define YOLO1_DTYPE_BF16 1

#define YOLO1_COORD_REFERENCE 0 /* row-slice behaviour of v1Loss.py:101 (first two objects plain, rest sqrt) */
#define YOLO1_COORD_PAPER 1     /* xy plain, wh sqrt for every object */

#define YOLO1_ERR_ARG (-1)         /* null pointer / negative size / bad enum */
#define YOLO1_ERR_UNSUPPORTED (-2) /* shape outside what the kernels are built for (see each call) */
#define YOLO1_ERR_ALIGN (-3)       /* a pointer is not aligned to its element size */

YOLO1_API int yolo1_abi_version(void);

/* Human readable text for a return code of this library (static storage). */
YOLO1_API const char* yolo1_error_string(int rc);

/* ------------------------------------------------------------------------------------------------
 * Loss: replaces YOLOLossV1.forward (v1Loss.py:22-118) AND its autograd backward (train.py:171) in one
 * fused pass.  B <= 8, 5B+C <= 128.
 * ---------------------------------------------------------------------------------------------- */

/* Bytes of device scratch `yolo1_loss_fwd_bwd` needs (independent of N; 64 KB + a header; 8-byte aligned). */
YOLO1_API size_t yolo1_loss_workspace_bytes(int64_t N, int S, int B, int C);

/*
 * pred [N,S,S,D] (pred_dtype, pred_strides), target [N,S,S,D] float32 (target_strides).
 * grad  : d total / d pred, same dtype as pred, layout grad_strides (NULL => forward only).
 * terms : device float[5] = {location, contain, not_contain, classify} each divided by batch_size (the
 *         four numbers v1Loss.py:108 logs) and the total loss (v1Loss.py:104-105).
 * inv_batch_size : 1 / _batch_size of the constructor (v1Loss.py:18,105) -- NOT 1/N.
 * coord_mode : YOLO1_COORD_REFERENCE reproduces v1Loss.py:101 exactly.
 * The gradient includes the path through the IoU target (v1Loss.py:78 keeps the graph).
 * grad must not overlap pred or target (the row-slice fix-up re-reads two cells of pred after the streaming pass).
 * Fast paths: contiguous [N,S,S,30] tensors or the permuted NCHW view, B = 2, C = 20, 16-byte aligned bases; any
 * other strides / alignment / (B, C) take the strided one-thread-per-cell kernel (same results, ~3x slower).
 * Small calls (train.py:38-41 trains with 12 x 14 x 14 = 2 352 cells; up to 7 360 fp32 cells of the two fast layouts,
 * 2 048 cells of any other) run as ONE 8-CTA cluster launch that resolves the `[:2]` rule before evaluating any cell
 * and does not touch the workspace (loss_small.cu).
 * N * S * S must be below 2^32 - 1 per call.
 */
YOLO1_API int yolo1_loss_fwd_bwd(const void* pred, const int64_t pred_strides[4], int pred_dtype,
                       const float* target, const int64_t target_strides[4],
                       void* grad, const int64_t grad_strides[4], float* terms,
                       int64_t N, int S, int B, int C,
                       float lambda_coord, float lambda_noobj, float inv_batch_size, int coord_mode,
                       void* workspace, size_t workspace_bytes, void* stream);

/*
 * Tuning / diagnostic twin of yolo1_loss_fwd_bwd: `variant` picks the launch shape of the streaming kernel
 * (0 = the default yolo1_loss_fwd_bwd uses; 1, 2, 3, 5, 8, 13 = other tile-cells x input-stages x output-buffers
 * shapes, see loss_nhwc.cu launch_tma_variant; 20 = force the warp-specialised kernel for a channel-planar view;
 * < 0 = force the strided one-thread-per-cell kernel that also serves non-contiguous views; 30 = force the
 * small-call cluster kernels (YOLO1_ERR_UNSUPPORTED when the call does not fit them: 7 360 fp32 NHWC cells resident,
 * 2 048 cells otherwise); 31 = the default streaming shape even for a small call; 40 / 41 / 42 = the sector-read
 * ("confidence first") kernel for fp32 NHWC tensors with 128 / 64 / 256 cells per tile -- an experiment that moves the
 * same DRAM bytes as the dense stream, see profiles/ncu_sparse_r2.md; 50 = force the confidence-first kernel for the
 * permuted NCHW view, 51 = force the dense planar kernels).  Same results for every variant up to summation order.
 */
YOLO1_API int yolo1_loss_fwd_bwd_ex(const void* pred, const int64_t pred_strides[4], int pred_dtype,
                          const float* target, const int64_t target_strides[4],
                          void* grad, const int64_t grad_strides[4], float* terms,
                          int64_t N, int S, int B, int C,
                          float lambda_coord, float lambda_noobj, float inv_batch_size, int coord_mode,
                          void* workspace, size_t workspace_bytes, int variant, void* stream);

/*
 * Loss with the head's epilogue fused in (backbones/OriginResNet.py:186-188, OriginDenseNet.py:124-128: the
 * network ends in `torch.sigmoid` followed by `permute(0,2,3,1)`): `logits` are the PRE-sigmoid outputs seen
 * through the same [N,S,S,D] strides (the permuted NCHW view is read in place); the kernel applies the sigmoid
 * on load and returns d total / d logit = d total / d p * p (1 - p) -- one full read + write of the S x S x D
 * tensor less in the forward and in the backward pass, and no sigmoid kernels.  Same terms as
 * yolo1_loss_fwd_bwd(sigmoid(logits), ...).
 */
YOLO1_API int yolo1_loss_fwd_bwd_logits(const void* logits, const int64_t logit_strides[4], int dtype,
                                        const float* target, const int64_t target_strides[4],
                                        void* grad, const int64_t grad_strides[4], float* terms,
                                        int64_t N, int S, int B, int C,
                                        float lambda_coord, float lambda_noobj, float inv_batch_size, int coord_mode,
                                        void* workspace, size_t workspace_bytes, void* stream);

/*
 * Loss straight from object lists -- the dense target tensor is never materialised.
 * boxes [n_obj,4] (cx,cy,w,h normalised to the image; 16-byte aligned), labels [n_obj], offsets [N+1] (CSR) are
 * what yoloDataset.encoder (utils/YOLODataLoader.py:200-230) would be fed per image; the call computes exactly
 * yolo1_loss_fwd_bwd(pred, encoder(boxes, labels), ...) (or ..._logits when from_logits != 0) but reads
 * 4 bytes per cell (the index of the owning object) instead of a 120-byte target row: 248 instead of 360
 * algorithmic bytes per cell.  Contiguous fp32 calls beyond the small-call sizes read not even that: a helper warp
 * of the streaming kernel finds the owners of the tiles ahead from the lists themselves (no pre-pass kernel, no
 * map: 240 bytes per cell).  workspace: yolo1_loss_objects_workspace_bytes(N,S,B,C) bytes, 16-byte aligned
 * (header + per-CTA partials + the N*S*S int32 cell map).  status: device int32, 1 if a centre / label was
 * outside the grid / class range (the reference raises IndexError; such objects are skipped), else 0.
 */
YOLO1_API size_t yolo1_loss_objects_workspace_bytes(int64_t N, int S, int B, int C);
YOLO1_API int yolo1_loss_fwd_bwd_objects(const void* pred, const int64_t pred_strides[4], int pred_dtype,
                                         int from_logits, const float* boxes, const int32_t* labels,
                                         const int64_t* offsets, void* grad, const int64_t grad_strides[4],
                                         float* terms, int64_t N, int S, int B, int C,
                                         float lambda_coord, float lambda_noobj, float inv_batch_size,
                                         int coord_mode, void* workspace, size_t workspace_bytes, int32_t* status,
                                         void* stream);

/* Tuning twin of yolo1_loss_fwd_bwd_objects: `variant` as in yolo1_loss_fwd_bwd_ex; in addition 60 = force the
 * form that finds the owners inside the streaming kernel (YOLO1_ERR_UNSUPPORTED where it does not apply), 61 = force
 * the pre-pass + ownership map.  Bit-identical results either way. */
YOLO1_API int yolo1_loss_fwd_bwd_objects_ex(const void* pred, const int64_t pred_strides[4], int pred_dtype,
                                            int from_logits, const float* boxes, const int32_t* labels,
                                            const int64_t* offsets, void* grad, const int64_t grad_strides[4],
                                            float* terms, int64_t N, int S, int B, int C,
                                            float lambda_coord, float lambda_noobj, float inv_batch_size,
                                            int coord_mode, void* workspace, size_t workspace_bytes,
                                            int32_t* status, int variant, void* stream);

/*
 * grad *= *grad_out_dev, in place, for autograd's backward(grad_output) (train.py:171 calls
 * loss.backward() with grad_output = 1; AMP loss scaling makes it != 1).  The kernel reads the scalar
 * on the device and returns without touching memory when it is exactly 1.0f, so the usual training
 * step pays one empty launch and no traffic.  `storage_numel` elements starting at grad (the dense
 * storage of the gradient, any stride order).
 */
YOLO1_API int yolo1_scale_grad(void* grad, int dtype, int64_t storage_numel, const float* grad_out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decode + NMS: replaces decoder (utils/utils.py:94-147) and nms (utils/utils.py:150-184), batched
 * over images.  B is the number of boxes per cell, max_n = S*S*B <= 1024 candidates per image.
 * Arithmetic is bit-exact with the reference's fp32 ATen CPU ops (no FMA contraction).
 * ---------------------------------------------------------------------------------------------- */

/*
 * decoder's candidate loop, utils/utils.py:108-132.  Per image n writes the boxes that pass
 * `conf*maxprob > thresh` in row-major (i,j,b) emission order:
 *   boxes [N,max_n,4] xyxy (image-normalised), scores [N,max_n], cls [N,max_n], counts [N].
 * `thresh` is a double because the reference compares in Python double (utils/utils.py:129).
 */
YOLO1_API int yolo1_decode(const void* pred, const int64_t pred_strides[4], int pred_dtype,
                 int64_t N, int S, int B, int C, double thresh,
                 float* boxes, float* scores, int32_t* cls, int32_t* counts, void* stream);

/*
 * nms, utils/utils.py:150-184, for N independent box sets of counts[n] <= max_n boxes each
 * (boxes [N,max_n,4], scores [N,max_n]).  keep [N,max_n] receives indices into the input set in
 * descending score order (ties: lower index first), keep_counts [N] their number.
 * per_class = 0 is the reference behaviour (class-agnostic; cls may be NULL); per_class = 1 only lets
 * boxes of equal cls suppress each other.  A box survives iff ovr <= (float)iou_thr.
 */
YOLO1_API int yolo1_nms(const float* boxes, const float* scores, const int32_t* cls, const int32_t* counts,
              int64_t N, int max_n, float iou_thr, int per_class,
              int32_t* keep, int32_t* keep_counts, void* stream);

/*
 * decoder end to end (utils/utils.py:94-147) in one kernel per image; candidates never leave shared
 * memory.  Outputs, per image, the kept detections in descending score order:
 *   out_boxes [N,max_n,4], out_scores [N,max_n], out_cls [N,max_n], out_counts [N]
 *   (out_counts[n] == 0 <=> the reference returns its all-zero sentinel, utils/utils.py:134-137).
 * Optional (may be NULL): keep_idx [N,max_n] = kept candidate indices in emission order (what nms()
 * returns), cand_counts [N] = candidates before NMS.
 */
YOLO1_API int yolo1_decode_nms(const void* pred, const int64_t pred_strides[4], int pred_dtype,
                     int64_t N, int S, int B, int C, double thresh, float iou_thr, int per_class,
                     float* out_boxes, float* out_scores, int32_t* out_cls, int32_t* out_counts,
                     int32_t* keep_idx, int32_t* cand_counts, void* stream);

/*
 * Post-processing of run_test_mAP (utils/utils.py:406-407 `bboxes.clamp(0,1)` and bbox_un_norm :347-354):
 * pixels[k] = (int)(clamp(boxes[k], 0, 1) * {img_w, img_h, img_w, img_h}) -- fp32 multiply, truncation --
 * for n_boxes xyxy boxes (boxes and pixels 16-byte aligned; int32 [n_boxes,4]).
 */
YOLO1_API int yolo1_boxes_to_pixels(const float* boxes, int64_t n_boxes, float img_w, float img_h, int32_t* pixels,
                                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Target encoder: replaces yoloDataset.encoder (utils/YOLODataLoader.py:200-230), batched.
 * boxes [n_obj,4] = (cx, cy, w, h) normalised to the image, labels [n_obj], offsets [N+1] (CSR: image n owns
 * objects offsets[n] .. offsets[n+1]-1, in the order the reference would iterate them) -- all device memory.
 * target [N,S,S,5B+C] float32 contiguous is fully written (zeros where no object): cell = ceil(c / fl32(1/S)) - 1,
 * the last object that falls into a cell wins, every confidence slot 1, class one-hot, the same
 * (dx, dy, w, h) in every box slot.  Bit-exact with the reference's fp32 arithmetic.
 * status: device int32, set to 1 if a centre or label was outside the grid / class range (the reference
 * raises IndexError there; such objects are skipped), else 0.  One image must fit 96 KB.
 * ---------------------------------------------------------------------------------------------- */
YOLO1_API int yolo1_encode_targets(const float* boxes, const int32_t* labels, const int64_t* offsets, int64_t N,
                                   int S, int B, int C, float* target, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer entry points: the same path for callers that hold HOST memory (the reference's callers
 * hold CPU tensors when run as shipped with device='cpu').  A context owns the device staging
 * buffers, the pinned bounce buffers and the copy/compute streams; batches are cut into chunks and
 * H2D copy, kernel and D2H copy of consecutive chunks overlap.  These calls block until the results
 * are in the host buffers.  One context per host thread.
 * ---------------------------------------------------------------------------------------------- */
typedef struct yolo1_host_ctx yolo1_host_ctx;

/* chunk_images: images per pipeline chunk (0 = pick automatically); max S*S*D per image fixed here. */
YOLO1_API int yolo1_host_ctx_create(yolo1_host_ctx** ctx, int device, int S, int B, int C, int64_t chunk_images);
YOLO1_API void yolo1_host_ctx_destroy(yolo1_host_ctx* ctx);

/*
 * Loss calls whose pred/target/grad buffers are pinned AND device-mapped (cudaHostAlloc, torch pin_memory(),
 * yolo1_host_pin) can be served without staging copies: the kernel reads the bytes it needs (one 32-byte
 * sector of target and of pred per cell; the full 240 B only for object cells) straight from host memory and
 * bulk-stores the gradient straight into the caller's buffer.  mode 2 (default): everything in place, one
 * launch.  mode 1: target streamed by the copy engine chunk by chunk, pred / grad in place.  mode 3: target up
 * and gradient down by the copy engines, only pred in place.  mode 0: the staged H2D / kernel / D2H pipeline
 * (always used for pageable memory).  All modes give the same results; measured rates are in DESIGN.md.
 */
YOLO1_API int yolo1_host_ctx_set_zero_copy(yolo1_host_ctx* ctx, int enable);

/* Register / unregister caller memory as pinned (cudaHostRegister) so the copies run at full PCIe rate. */
YOLO1_API int yolo1_host_pin(void* ptr, size_t bytes);
YOLO1_API int yolo1_host_unpin(void* ptr);

/* pred/target/grad: contiguous float32 [N,S,S,D] in host memory; terms: host float[5]. grad may be NULL. */
YOLO1_API int yolo1_loss_fwd_bwd_host(yolo1_host_ctx* ctx, const float* pred, const float* target, float* grad,
                            float terms[5], int64_t N, float lambda_coord, float lambda_noobj,
                            float inv_batch_size, int coord_mode);

/* pred: contiguous float32 [N,S,S,D] in host memory; outputs as yolo1_decode_nms, in host memory. */
YOLO1_API int yolo1_decode_nms_host(yolo1_host_ctx* ctx, const float* pred, int64_t N, double thresh, float iou_thr,
                          int per_class, float* out_boxes, float* out_scores, int32_t* out_cls,
                          int32_t* out_counts);

#ifdef __cplusplus
}
#endif
#endif /* YOLO1_B200_H_ */
