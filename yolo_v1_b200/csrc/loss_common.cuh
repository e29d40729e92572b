// loss_common.cuh -- shared device code of K1 (loss.cu, loss_nhwc.cu, loss_planar.cu): YOLO v1 loss forward + backward in one fused pass (sm_100a).
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
#pragma once

#include "common.cuh"

namespace yolo1 {


constexpr int kMaxB = 8;
constexpr int kMaxGrid = 2048;
constexpr int kGenericThreads = 256;
constexpr int kVariantHostMapped = 100;  // loss_launch_chunk: the tensors are pinned, mapped host memory

struct LossWs {
  unsigned long long pair;  // (~idx of 1st object cell) << 32 | (~idx of 2nd) of the current chunk; 0 = none
  unsigned int ticket;      // CTAs finished in the current launch
  unsigned int carry;       // object cells seen in earlier chunks of this call, saturating at 2
  double acc[4];            // raw sums (loc, hit, miss, cls) over the chunks so far
  double partial[kMaxGrid][4];
};

struct LossParams {
  const void* pred;
  const float* target;
  void* grad;
  float* terms;
  LossWs* ws;
  int64_t ps[4], ts[4], gs[4];  // element strides over (n, i, j, channel)
  int64_t cells;                // N*S*S of this launch
  int S, B, C;
  float Sf, lc, ln, inv_bs;
  float k2ln;  // 2 * lambda_noobj / batch_size
  float k2;    // 2 / batch_size
  int coord_mode;
  int last_chunk;
  int logits;  // 1: `pred` holds the head's pre-sigmoid outputs; the kernel applies sigmoid and returns d loss / d logit
  // object-list targets (yolo1_loss_fwd_bwd_objects): the dense target tensor is never materialised.
  // cellobj[q] = index of the object that owns cell q (the last one that falls into it, as the reference encoder
  // resolves collisions) or -1; the target values of an object cell are recomputed from boxes / labels.
  // list_mode 2 (contiguous streaming kernel): no map at all -- the kernel finds the owners itself from `offsets`
  // (objects of image n are [offsets[n], offsets[n+1])) and reports objects outside the grid through `status`.
  int list_mode;
  const int32_t* cellobj;
  const float* boxes;
  const int32_t* labels;
  const int64_t* offsets;
  int32_t* status;
  float cs;  // fl32(1/S)
};

struct CellSums {
  float loc, hit, miss, cls;
};


// host-side launchers of the two streaming kernels (loss_nhwc.cu, loss_planar.cu); dispatch lives in loss.cu
int launch_loss_nhwc(const LossParams& p, bool bf16, bool has_grad, int variant, cudaStream_t stream);
int launch_loss_nhwc_any(const LossParams& p, bool bf16, bool has_grad, cudaStream_t stream);   // any (B, C)
int launch_loss_planar(const LossParams& p, bool bf16, bool has_grad, int tile_imgs, cudaStream_t stream);
int planar_tile_imgs(int S, size_t esz, int target_cells, bool list_mode);
// warp-specialised form (loss_ws.cu): tile_cells cells per tile, stages 2 or 3, gradient tile in place
int launch_loss_ws(const LossParams& p, bool bf16, bool has_grad, bool is_planar, int tile_cells, int stages,
                   cudaStream_t stream);
// small calls (loss_small.cu): one cluster, no workspace; any layout / (B, C) / dtype
int64_t loss_small_max_cells();
int launch_loss_small(const LossParams& p, bool bf16, bool has_grad, int layout, cudaStream_t stream);
constexpr int64_t kSmallTryCells = 16384;   // calls up to this size ask the small path first
// confidence-first form (loss_sparse.cu): sector reads of target[0] / pred[0:2], full rows for object cells only
int launch_loss_sparse(const LossParams& p, bool has_grad, int shape, cudaStream_t stream);
constexpr int kVariantSparse = 40;  // 40, 41, 42: 128 / 64 / 256 cells per tile
// the same idea where it pays: the channel-planar view, whose confidence planes are contiguous (loss_planar_sparse.cu)
int launch_loss_planar_sparse(const LossParams& p, bool bf16, bool has_grad, int tile_imgs, cudaStream_t stream);
constexpr int kVariantPlanarSparse = 50;   // force it; 51 = force the dense planar kernels
constexpr int kVariantSmall = 30;   // yolo1_loss_fwd_bwd_ex: force the small-call kernel (error when too large)
constexpr int kVariantNoSmall = 31; // ... or keep a small call on the streaming kernels (A/B measurements)

namespace {

// ---- box math: utils/utils.py:59-75 and :10-57.  The forward IoU is evaluated with explicitly rounded
// operations (no FMA contraction) so that the arg-max over the B predictors takes the same decision as
// the reference's ATen ops even when two IoUs are one ulp apart. ----
__device__ __forceinline__ void to_xyxy(const float b[4], float S, float o[4]) {
  const float cx = __fdiv_rn(b[0], S), cy = __fdiv_rn(b[1], S);
  const float hw = 0.5f * b[2], hh = 0.5f * b[3];
  o[0] = __fsub_rn(cx, hw);
  o[1] = __fsub_rn(cy, hh);
  o[2] = __fadd_rn(cx, hw);
  o[3] = __fadd_rn(cy, hh);
}

__device__ __forceinline__ float iou_xyxy(const float p[4], const float g[4]) {
  const float lx = p[0] > g[0] ? p[0] : g[0], ly = p[1] > g[1] ? p[1] : g[1];
  const float rx = p[2] < g[2] ? p[2] : g[2], ry = p[3] < g[3] ? p[3] : g[3];
  float iw = __fsub_rn(rx, lx), ih = __fsub_rn(ry, ly);
  if (iw < 0.f) iw = 0.f;
  if (ih < 0.f) ih = 0.f;
  const float inter = __fmul_rn(iw, ih);
  const float ap = __fmul_rn(__fsub_rn(p[2], p[0]), __fsub_rn(p[3], p[1]));
  const float ag = __fmul_rn(__fsub_rn(g[2], g[0]), __fsub_rn(g[3], g[1]));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(ap, ag), inter));
}

// d IoU / d (x, y, w, h) of the predicted box: the sub-gradient autograd takes through
// utils/utils.py:38-55 and :72-73; min/max ties split 0.5/0.5, a clipped extent kills the gradient.
__device__ __forceinline__ void iou_grad(const float p[4], const float g[4], float S, float dg[4]) {
  const float lx = p[0] > g[0] ? p[0] : g[0], ly = p[1] > g[1] ? p[1] : g[1];
  const float rx = p[2] < g[2] ? p[2] : g[2], ry = p[3] < g[3] ? p[3] : g[3];
  const float iw = rx - lx, ih = ry - ly;
  dg[0] = dg[1] = dg[2] = dg[3] = 0.f;
  if (iw < 0.f || ih < 0.f) return;
  const float pw = p[2] - p[0], ph = p[3] - p[1];
  const float ap = pw * ph, ag = (g[2] - g[0]) * (g[3] - g[1]);
  const float I = iw * ih, U = ap + ag - I;
  const float inv = 1.0f / U;
  const float a = (ap + ag) * inv * inv, c = I * inv * inv;
  const float m2x = p[2] < g[2] ? 1.f : (p[2] == g[2] ? 0.5f : 0.f);
  const float m1x = p[0] > g[0] ? 1.f : (p[0] == g[0] ? 0.5f : 0.f);
  const float m2y = p[3] < g[3] ? 1.f : (p[3] == g[3] ? 0.5f : 0.f);
  const float m1y = p[1] > g[1] ? 1.f : (p[1] == g[1] ? 0.5f : 0.f);
  dg[0] = a * ih * (m2x - m1x) / S;
  dg[1] = a * iw * (m2y - m1y) / S;
  dg[2] = a * ih * (m2x + m1x) * 0.5f - c * ph;
  dg[3] = a * iw * (m2y + m1y) * 0.5f - c * pw;
}

// one coordinate of the location term (v1Loss.py:101): returns d loc / d p (without lambda / batch_size)
__device__ __forceinline__ float coord_term(float p, float g, bool plain, float& loc) {
  if (plain) {
    const float e = p - g;
    loc += e * e;
    return 2.0f * e;
  }
  const float sp = sqrtf(p), sg = sqrtf(g);
  const float e = sp - sg;
  loc += e * e;
  return e / sp;
}

// ---- accessors: a cell seen as D consecutive channels --------------------------------------------
struct SmemInF32 {
  const float* p;
  __device__ __forceinline__ float2 ld2(int c) const { return *reinterpret_cast<const float2*>(p + c); }
};
struct SmemInBF16 {
  const __nv_bfloat16* p;
  __device__ __forceinline__ float2 ld2(int c) const {
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p + c));
  }
};
// writers: st2 stores a gradient pair; st2p additionally receives the two probabilities the caller already holds --
// ignored here, used by the fused-head wrapper SigOut (d loss / d z = d loss / d p * p (1 - p)) so that no sigmoid
// is evaluated twice
struct SmemOutF32 {
  float* p;
  __device__ __forceinline__ void st2(int c, float x, float y) const {
    *reinterpret_cast<float2*>(p + c) = make_float2(x, y);
  }
  __device__ __forceinline__ void st2p(int c, float x, float y, float, float) const { st2(c, x, y); }
};
struct SmemOutBF16 {
  __nv_bfloat16* p;
  __device__ __forceinline__ void st2(int c, float x, float y) const {
    *reinterpret_cast<__nv_bfloat162*>(p + c) = __floats2bfloat162_rn(x, y);
  }
  __device__ __forceinline__ void st2p(int c, float x, float y, float, float) const { st2(c, x, y); }
};
template <typename E>
struct SmemIn;
template <>
struct SmemIn<float> {
  using type = SmemInF32;
};
template <>
struct SmemIn<__nv_bfloat16> {
  using type = SmemInBF16;
};
template <typename E>
struct SmemOut;
template <>
struct SmemOut<float> {
  using type = SmemOutF32;
};
template <>
struct SmemOut<__nv_bfloat16> {
  using type = SmemOutBF16;
};

// head epilogue fusion (backbones/OriginResNet.py:186-188: ... bn_end -> torch.sigmoid -> permute): p = sigmoid(z),
// d loss / d z = d loss / d p * p (1 - p)
// Branch-free: one MUFU.EX2, one MUFU.RCP and one Newton step on the reciprocal (two FMAs; the result is then
// within 1 ulp of 1/d, so the only approximation left is ex2's 2 ulp).  The IEEE forms (expf, 1/x, __frcp_rn) carry
// a slow-path branch each; 30 of them in a row on the one or two lanes of a warp that hold an object serialise into
// a dependent chain the rest of the CTA waits for at the tile barrier (measured: 1.22 ms vs 0.72 ms without the head).
// The exponent is clamped to 126 so that d = 1 + 2^t stays finite (d = inf would turn the Newton step into
// inf * 0 = NaN); sigmoid(z) for z < -87 is then 2^-126 instead of a smaller number -- both are 0 at 1e-5.
// Error budget: ex2.approx 2^-22 relative, the rounded product z * -log2(e) another |z| 2^-24, the reciprocal
// 2^-23 -- under 1e-6 for |z| < 10, against the 1e-5 gate of the fused-head tests.
__device__ __forceinline__ float sigmoid_(float z) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(z * -1.4426950408889634f, 126.0f)));
  const float d = 1.0f + e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return fmaf(r, fmaf(-d, r, 1.0f), r);
}
__device__ __forceinline__ float dsigmoid_(float z) {
  const float pz = sigmoid_(z);
  return pz * (1.0f - pz);
}

template <typename E>
struct GlobIn {
  const E* p;
  int64_t cs;
  bool sig;  // values are logits: apply sigmoid on load
  __device__ __forceinline__ float ld(int c) const {
    const float v = ld_elem(p + c * cs);
    return sig ? sigmoid_(v) : v;
  }
};
template <typename E>
struct GlobOut {
  E* p;
  int64_t cs;
  const E* z;  // logits of the same cell (sig only)
  int64_t zs;
  bool sig;
  __device__ __forceinline__ void st(int c, float v) const {
    if (sig && v != 0.f) v *= dsigmoid_(ld_elem(z + c * zs));
    st_elem(p + c * cs, v);
  }
};
// scalar view of a cell inside a shared-memory tile (runtime B, C; optional sigmoid head)
template <typename E>
struct SmemInS {
  const E* p;
  bool sig;
  __device__ __forceinline__ float ld(int c) const {
    const float v = ld_elem(p + c);
    return sig ? sigmoid_(v) : v;
  }
};
template <typename E>
struct SmemOutS {
  E* p;
  const E* z;  // logits of the same cell (sig only)
  bool sig;
  __device__ __forceinline__ void st(int c, float v) const {
    if (sig && v != 0.f) v *= dsigmoid_(ld_elem(z + c));
    st_elem(p + c, v);
  }
};
// wrappers that put the sigmoid head in front of any pair accessor (shared-memory tiles)
template <typename In>
struct SigIn {
  In in;
  __device__ __forceinline__ float2 ld2(int c) const {
    const float2 v = in.ld2(c);
    return make_float2(sigmoid_(v.x), sigmoid_(v.y));
  }
};
template <typename Out>
struct SigOut {
  Out out;
  __device__ __forceinline__ void st2(int c, float x, float y) const { out.st2(c, x, y); }   // zeros only
  __device__ __forceinline__ void st2p(int c, float x, float y, float px, float py) const {
    out.st2p(c, x * (px * (1.0f - px)), y * (py * (1.0f - py)), px, py);   // (st2p: a writer may drop plain st2 zeros)
  }
};

// channel-planar tile in shared memory: channel c of a cell lives `plane` elements apart (lanes <-> consecutive
// cells, so 32-bit accesses are conflict-free)
template <typename E>
struct PlanarIn {
  const E* p;
  int plane;
  __device__ __forceinline__ float2 ld2(int c) const {
    return make_float2(ld_elem(p + c * plane), ld_elem(p + (c + 1) * plane));
  }
};
template <typename E>
struct PlanarOut {
  E* p;
  int plane;
  __device__ __forceinline__ void st2(int c, float x, float y) const {
    st_elem(p + c * plane, x);
    st_elem(p + (c + 1) * plane, y);
  }
  __device__ __forceinline__ void st2p(int c, float x, float y, float, float) const { st2(c, x, y); }
};

// ---- targets from object lists: what the reference encoder would have written into the cell ---------------
// (utils/YOLODataLoader.py:220-227: every confidence slot 1, the same (dx, dy, w, h) in every box slot, class
// one-hot).  k < 0: a cell without object (all zero).
struct ListTarget2 {  // channel-pair view for cell_b2c20 (B = 2, C = 20)
  float dx, dy, w, h;
  int label;
  bool obj;
  __device__ __forceinline__ float2 ld2(int c) const {
    if (c == 0) return obj ? make_float2(1.f, 1.f) : make_float2(0.f, 0.f);
    if (c < 10) return ((c - 2) & 2) ? make_float2(w, h) : make_float2(dx, dy);
    return make_float2(c - 10 == label ? 1.f : 0.f, c - 9 == label ? 1.f : 0.f);
  }
};
struct ListTargetS {  // scalar view for cell_generic (any B, C)
  float v[4];
  int label, B;
  bool obj;
  __device__ __forceinline__ float ld(int c) const {
    if (!obj) return 0.f;
    if (c < B) return 1.f;
    if (c < 5 * B) return v[(c - B) & 3];
    return c - 5 * B == label ? 1.f : 0.f;
  }
};
// the raw object record of a cell, fetched one tile ahead of its use (the gather is a dependent global load:
// issued at the end of the previous tile it completes behind the barrier and the next mbarrier wait)
struct ObjFetch {
  float4 box;
  int label;
  int k;
};
__device__ __forceinline__ ObjFetch fetch_object(const LossParams& p, int k) {
  ObjFetch o;
  o.k = k, o.box = make_float4(0.f, 0.f, 0.f, 0.f), o.label = 0;
  if (k >= 0) {
    o.box = __ldg(reinterpret_cast<const float4*>(p.boxes + 4 * (int64_t)k));
    o.label = __ldg(p.labels + k);
  }
  return o;
}
__device__ __forceinline__ ListTarget2 list_target2(const LossParams& p, const ObjFetch& o) {
  ListTarget2 t = {0.f, 0.f, 0.f, 0.f, -1, false};
  if (o.k >= 0) {
    float f;
    encode_axis(o.box.x, p.cs, f, t.dx);
    encode_axis(o.box.y, p.cs, f, t.dy);
    t.w = o.box.z, t.h = o.box.w, t.obj = true;
    t.label = o.label < 0 ? o.label + p.C : o.label;
  }
  return t;
}
__device__ __forceinline__ ListTarget2 list_target2(const LossParams& p, int k) {
  return list_target2(p, fetch_object(p, k));
}
// The cell an object falls into, as the reference encoder computes it (utils/YOLODataLoader.py:218-222, Python
// indexing for negative indices); false = outside the grid or label out of range (the reference raises IndexError).
__device__ __forceinline__ bool object_cell(const LossParams& p, const float4& box, int label, int& cell) {
  float fi, fj, d;
  encode_axis(box.x, p.cs, fi, d);
  encode_axis(box.y, p.cs, fj, d);
  int col = (int)fi, row = (int)fj;
  const int S = p.S;
  if (col < -S || col >= S || row < -S || row >= S || label < -p.C || label >= p.C) return false;
  if (col < 0) col += S;
  if (row < 0) row += S;
  cell = row * S + col;
  return true;
}
// list_mode 2, one cell at a time (ragged tail, the last CTA's fix-up): the last object of the cell's image that
// falls into it, or -1
__device__ __forceinline__ int owner_direct(const LossParams& p, int64_t q) {
  const int SS = p.S * p.S;
  const int64_t n = q / SS;
  const int c = (int)(q - n * SS);
  int owner = -1;
  for (int64_t k = __ldg(p.offsets + n), hi = __ldg(p.offsets + n + 1); k < hi; ++k) {
    int cell;
    if (object_cell(p, __ldg(reinterpret_cast<const float4*>(p.boxes + 4 * k)), __ldg(p.labels + k), cell)) {
      if (cell == c) owner = (int)k;
    } else {
      atomicExch(p.status, 1);
    }
  }
  return owner;
}
__device__ __forceinline__ ListTargetS list_targetS(const LossParams& p, int64_t q) {
  const ListTarget2 a = list_target2(p, p.list_mode == 2 ? owner_direct(p, q) : p.cellobj[q]);
  ListTargetS t;
  t.v[0] = a.dx, t.v[1] = a.dy, t.v[2] = a.w, t.v[3] = a.h, t.label = a.label, t.B = p.B, t.obj = a.obj;
  return t;
}

// ---- fast cell: B = 2, C = 20, channel pairs (conflict-free 64-bit shared accesses) -----------------
// Returns true when the cell holds an object (target channel 0 == 1, v1Loss.py:28).
// plain = true (small-call kernel): this is one of the call's first two object cells -- plain form right away.
template <bool HAS_GRAD, typename PA, typename TA, typename GA>
__device__ __forceinline__ bool cell_b2c20(const PA& P, const TA& T, const GA& G, const LossParams& k,
                                           CellSums& s, bool plain = false) {
  const float2 t01 = T.ld2(0);
  const float2 c01 = P.ld2(0);
  if (t01.x != 1.0f) {
    // v1Loss.py:91 -- both slots of a cell without object: conf^2 against the untouched 0 target
    s.miss += c01.x * c01.x + c01.y * c01.y;
    if (HAS_GRAD) {
      G.st2p(0, k.k2ln * c01.x, k.k2ln * c01.y, c01.x, c01.y);
#pragma unroll
      for (int c = 2; c < 30; c += 2) G.st2(c, 0.f, 0.f);
    }
    return false;
  }
  // v1Loss.py:66-74 -- IoU of both predictors against GT slot 0, first arg-max wins
  float2 a = T.ld2(2), b = T.ld2(4);
  const float g0[4] = {a.x, a.y, b.x, b.y};
  a = P.ld2(2), b = P.ld2(4);
  const float p0[4] = {a.x, a.y, b.x, b.y};
  a = P.ld2(6), b = P.ld2(8);
  const float p1[4] = {a.x, a.y, b.x, b.y};
  float gx[4], px0[4], px1[4];
  to_xyxy(g0, k.Sf, gx);
  to_xyxy(p0, k.Sf, px0);
  to_xyxy(p1, k.Sf, px1);
  const float iou0 = iou_xyxy(px0, gx), iou1 = iou_xyxy(px1, gx);
  const bool r = iou1 > iou0;
  const float best = r ? iou1 : iou0;
  // class term, v1Loss.py:33-41
  float cls = 0.f;
#pragma unroll
  for (int c = 10; c < 30; c += 2) {
    const float2 pv = P.ld2(c), tv = T.ld2(c);
    const float dx = pv.x - tv.x, dy = pv.y - tv.y;
    cls += dx * dx + dy * dy;
    if (HAS_GRAD) G.st2p(c, k.k2 * dx, k.k2 * dy, pv.x, pv.y);
  }
  s.cls += cls;
  // confidences, v1Loss.py:90-91 (the IoU target is not detached: see the -2 dconf dIoU term below)
  const float conf_r = r ? c01.y : c01.x, conf_o = r ? c01.x : c01.y;
  const float dconf = conf_r - best;
  s.hit += dconf * dconf;
  s.miss += conf_o * conf_o;
  // coordinates, v1Loss.py:94-101: GT slot r.  Square-root form here; the call's first two objects are
  // re-evaluated in plain form by finalize_fixup (reference mode) -- paper mode: xy plain, wh sqrt.
  float gr[4] = {g0[0], g0[1], g0[2], g0[3]};
  if (r) {
    a = T.ld2(6), b = T.ld2(8);
    gr[0] = a.x, gr[1] = a.y, gr[2] = b.x, gr[3] = b.y;
  }
  const float pr[4] = {r ? p1[0] : p0[0], r ? p1[1] : p0[1], r ? p1[2] : p0[2], r ? p1[3] : p0[3]};
  const bool paper = k.coord_mode == YOLO1_COORD_PAPER;
  float loc = 0.f, gl[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) gl[d] = coord_term(pr[d], gr[d], plain || (paper && d < 2), loc);
  s.loc += loc;
  if (HAS_GRAD) {
    float dI[4];
    iou_grad(r ? px1 : px0, gx, k.Sf, dI);
    float gv[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) gv[d] = (k.lc * gl[d] - 2.0f * dconf * dI[d]) * k.inv_bs;
    const float g_r = k.k2 * dconf, g_o = k.k2ln * conf_o;
    G.st2p(0, r ? g_o : g_r, r ? g_r : g_o, c01.x, c01.y);
    G.st2p(2, r ? 0.f : gv[0], r ? 0.f : gv[1], p0[0], p0[1]);
    G.st2p(4, r ? 0.f : gv[2], r ? 0.f : gv[3], p0[2], p0[3]);
    G.st2p(6, r ? gv[0] : 0.f, r ? gv[1] : 0.f, p1[0], p1[1]);
    G.st2p(8, r ? gv[2] : 0.f, r ? gv[3] : 0.f, p1[2], p1[3]);
  }
  return true;
}

// ---- generic cell: any B <= 8, any C, any strides -------------------------------------------------
// FIX = false: streaming pass (square-root / paper form).  FIX = true: finalize pass for one of the
// call's first two object cells: writes only the 4 coordinate gradients of the responsible box in plain
// form and returns (plain - sqrt) of the location sum in s.loc.
// ZEROED = true: the gradient row was already cleared (cooperative vector stores), skip the zero stores.
// plain = true (small-call kernel, loss_small.cu): the caller already knows that this is one of the call's first two
// object cells, so the cell is evaluated in plain form right away and no fix-up follows.
template <bool HAS_GRAD, bool FIX, bool ZEROED = false, typename PA, typename TA, typename GA>
__device__ __forceinline__ bool cell_generic(const PA& P, const TA& T, const GA& G, const LossParams& k,
                                             CellSums& s, bool plain = false) {
  const int B = k.B, C = k.C, D = 5 * B + C;
  if (T.ld(0) != 1.0f) {
    if (!FIX) {
      for (int b = 0; b < B; ++b) {
        const float cf = P.ld(b);
        s.miss += cf * cf;
        if (HAS_GRAD) G.st(b, k.k2ln * cf);
      }
      if (HAS_GRAD && !ZEROED)
        for (int c = B; c < D; ++c) G.st(c, 0.f);
    }
    return false;
  }
  float g0[4], gx[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) g0[d] = T.ld(B + d);
  to_xyxy(g0, k.Sf, gx);
  int r = 0;
  float best = 0.f, pr[4] = {0.f, 0.f, 0.f, 0.f}, pxr[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b = 0; b < B; ++b) {
    float pb[4], px[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) pb[d] = P.ld(B + 4 * b + d);
    to_xyxy(pb, k.Sf, px);
    const float v = iou_xyxy(px, gx);
    if (b == 0 || v > best) {
      best = v;
      r = b;
#pragma unroll
      for (int d = 0; d < 4; ++d) pr[d] = pb[d], pxr[d] = px[d];
    }
  }
  if (!FIX) {
    float cls = 0.f;
    for (int c = 0; c < C; ++c) {
      const float d = P.ld(5 * B + c) - T.ld(5 * B + c);
      cls += d * d;
      if (HAS_GRAD) G.st(5 * B + c, k.k2 * d);
    }
    s.cls += cls;
  }
  float dconf = 0.f;
  for (int b = 0; b < B; ++b) {
    const float cf = P.ld(b);
    if (b == r) {
      dconf = cf - best;
      if (!FIX) {
        s.hit += dconf * dconf;
        if (HAS_GRAD) G.st(b, k.k2 * dconf);
      }
    } else if (!FIX) {
      s.miss += cf * cf;
      if (HAS_GRAD) G.st(b, k.k2ln * cf);
    }
  }
  float dI[4] = {0.f, 0.f, 0.f, 0.f};
  if (HAS_GRAD) iou_grad(pxr, gx, k.Sf, dI);
  const bool paper = k.coord_mode == YOLO1_COORD_PAPER;
  float loc = 0.f, loc_sqrt = 0.f;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const float g = T.ld(B + 4 * r + d);
    const float gl = coord_term(pr[d], g, FIX || plain || (paper && d < 2), loc);
    if (FIX) (void)coord_term(pr[d], g, false, loc_sqrt);
    if (HAS_GRAD) G.st(B + 4 * r + d, (k.lc * gl - 2.0f * dconf * dI[d]) * k.inv_bs);
  }
  s.loc += loc - loc_sqrt;
  if (!FIX && HAS_GRAD && !ZEROED)
    for (int b = 0; b < B; ++b)
      if (b != r)
        for (int d = 0; d < 4; ++d) G.st(B + 4 * b + d, 0.f);
  return true;
}

__device__ __forceinline__ int64_t cell_offset(const int64_t st[4], int64_t q, int S) {
  const int64_t n = q / (S * S);
  const int rem = (int)(q - n * (S * S));
  const int i = rem / S, j = rem - i * S;
  return n * st[0] + i * st[1] + j * st[2];
}

// ---- block epilogue: partial sums, first-two-objects pair, last-CTA finalize -----------------------
__device__ __forceinline__ void merge_pair(uint32_t& a1, uint32_t& a2, uint32_t b1, uint32_t b2) {
  // values are inverted cell indices (larger = earlier cell, 0 = none); keep the two largest
  const uint32_t hi = max(a1, b1), lo = min(a1, b1);
  a2 = max(lo, max(a2, b2));
  a1 = hi;
}
__device__ __forceinline__ void note_object(uint32_t& m1, uint32_t& m2, int64_t q) {
  const uint32_t v = 0xFFFFFFFFu - (uint32_t)q;
  merge_pair(m1, m2, v, 0u);
}

template <typename E, bool HAS_GRAD, bool BULK>
__device__ __noinline__ void block_epilogue(CellSums s, uint32_t m1, uint32_t m2, const LossParams& p) {
  __shared__ double red[32][4];
  __shared__ uint32_t redm[32][2];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  double v[4] = {(double)s.loc, (double)s.hit, (double)s.miss, (double)s.cls};
#pragma unroll
  for (int t = 0; t < 4; ++t) v[t] = warp_sum(v[t]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const uint32_t b1 = __shfl_xor_sync(0xffffffffu, m1, o), b2 = __shfl_xor_sync(0xffffffffu, m2, o);
    merge_pair(m1, m2, b1, b2);
  }
  if (lane == 0) {
#pragma unroll
    for (int t = 0; t < 4; ++t) red[warp][t] = v[t];
    redm[warp][0] = m1;
    redm[warp][1] = m2;
  }
  __syncthreads();
  LossWs* ws = p.ws;
  if (threadIdx.x == 0) {
    double t4[4] = {0, 0, 0, 0};
    uint32_t a1 = 0, a2 = 0;
    for (int w = 0; w < nwarps; ++w) {
#pragma unroll
      for (int t = 0; t < 4; ++t) t4[t] += red[w][t];
      merge_pair(a1, a2, redm[w][0], redm[w][1]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) __stcg(&ws->partial[blockIdx.x][t], t4[t]);
    if (a1 != 0u) {
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&ws->pair);
      while (true) {
        uint32_t c1 = (uint32_t)(cur >> 32), c2 = (uint32_t)cur;
        if (a1 <= c2) break;  // cannot improve on the two earliest cells already recorded
        merge_pair(c1, c2, a1, a2);
        const unsigned long long want = ((unsigned long long)c1 << 32) | c2;
        const unsigned long long old = atomicCAS(&ws->pair, cur, want);
        if (old == cur) break;
        cur = old;
      }
    }
    if (HAS_GRAD && BULK) {
      bulk_wait_all<0>();  // this thread's bulk gradient stores have landed
      fence_async_all();
    }
    __threadfence();
    const unsigned int old = atomicAdd(&ws->ticket, 1u);
    s_last = (old == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  // ---- last CTA of the launch ----
  // Three jobs run side by side: warp 0 adds up the per-CTA partial sums; lane 0 of two other warps re-evaluates one
  // of the call's first two object cells each.  (One lane doing both after the reduction was a chain of dependent
  // global loads, ~7.5 us per launch whatever the call size: tools/loss_size_sweep.py, reference vs paper mode.)
  __shared__ double fixloc[2];
  __threadfence();
  const unsigned long long pr = *reinterpret_cast<volatile unsigned long long*>(&ws->pair);
  const uint32_t c[2] = {(uint32_t)(pr >> 32), (uint32_t)pr};   // c[1] != 0 implies c[0] != 0
  const unsigned int carry = *reinterpret_cast<volatile unsigned int*>(&ws->carry);
  if (threadIdx.x < 2) fixloc[threadIdx.x] = 0.0;
  __syncthreads();
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    // v1Loss.py:101 `[:2]`: object t of this chunk is one of the first two of the CALL -> plain form
    if (warp == (t + 1) % nwarps && lane == 0 && c[t] != 0u && carry + t < 2 && p.coord_mode == YOLO1_COORD_REFERENCE) {
      const int64_t q = (int64_t)(0xFFFFFFFFu - c[t]);
      const E* zq = reinterpret_cast<const E*>(p.pred) + cell_offset(p.ps, q, p.S);
      GlobIn<E> P{zq, p.ps[3], p.logits != 0};
      GlobOut<E> G{HAS_GRAD ? reinterpret_cast<E*>(p.grad) + cell_offset(p.gs, q, p.S) : nullptr, p.gs[3], zq,
                   p.ps[3], p.logits != 0};
      CellSums d = {0.f, 0.f, 0.f, 0.f};
      if (p.list_mode) {
        cell_generic<HAS_GRAD, true>(P, list_targetS(p, q), G, p, d);
      } else {
        GlobIn<float> T{p.target + cell_offset(p.ts, q, p.S), p.ts[3], false};
        cell_generic<HAS_GRAD, true>(P, T, G, p, d);
      }
      fixloc[t] = (double)d.loc;
    }
  }
  double t4[4] = {0, 0, 0, 0};
  if (warp == 0) {
    for (unsigned int b = lane; b < gridDim.x; b += 32) {
#pragma unroll
      for (int t = 0; t < 4; ++t) t4[t] += __ldcg(&ws->partial[b][t]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) t4[t] = warp_sum(t4[t]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    t4[0] += fixloc[0] + fixloc[1];
    const unsigned int seen = carry + (c[0] != 0u ? 1u : 0u) + (c[1] != 0u ? 1u : 0u);
    double acc[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      acc[t] = ws->acc[t] + t4[t];
      ws->acc[t] = acc[t];
    }
    ws->carry = seen > 2 ? 2 : seen;
    ws->pair = 0ull;
    ws->ticket = 0u;
    if (p.last_chunk) {
      // v1Loss.py:104-108: the four logged components and the total, each / batch_size
      const double ib = (double)p.inv_bs;
      p.terms[0] = (float)(acc[0] * ib);
      p.terms[1] = (float)(acc[1] * ib);
      p.terms[2] = (float)(acc[2] * ib);
      p.terms[3] = (float)(acc[3] * ib);
      p.terms[4] = (float)(((double)p.lc * acc[0] + acc[1] + (double)p.ln * acc[2] + acc[3]) * ib);
    }
  }
}


}  // namespace
}  // namespace yolo1
