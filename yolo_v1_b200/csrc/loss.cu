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
#include "common.cuh"

namespace yolo1 {
namespace {

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
};

struct CellSums {
  float loc, hit, miss, cls;
};

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
struct SmemOutF32 {
  float* p;
  __device__ __forceinline__ void st2(int c, float x, float y) const {
    *reinterpret_cast<float2*>(p + c) = make_float2(x, y);
  }
};
struct SmemOutBF16 {
  __nv_bfloat16* p;
  __device__ __forceinline__ void st2(int c, float x, float y) const {
    *reinterpret_cast<__nv_bfloat162*>(p + c) = __floats2bfloat162_rn(x, y);
  }
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
__device__ __forceinline__ float sigmoid_(float z) { return 1.0f / (1.0f + expf(-z)); }
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
// wrappers that put the sigmoid head in front of any pair accessor (shared-memory tiles)
template <typename In>
struct SigIn {
  In in;
  __device__ __forceinline__ float2 ld2(int c) const {
    const float2 v = in.ld2(c);
    return make_float2(sigmoid_(v.x), sigmoid_(v.y));
  }
};
template <typename Out, typename In>
struct SigOut {
  Out out;
  In z;  // the logits of the same cell; read before the (possibly aliasing) store
  __device__ __forceinline__ void st2(int c, float x, float y) const {
    if (x != 0.f || y != 0.f) {
      const float2 v = z.ld2(c);
      x *= dsigmoid_(v.x), y *= dsigmoid_(v.y);
    }
    out.st2(c, x, y);
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
};

// ---- fast cell: B = 2, C = 20, channel pairs (conflict-free 64-bit shared accesses) -----------------
// Returns true when the cell holds an object (target channel 0 == 1, v1Loss.py:28).
template <bool HAS_GRAD, typename PA, typename TA, typename GA>
__device__ __forceinline__ bool cell_b2c20(const PA& P, const TA& T, const GA& G, const LossParams& k,
                                           CellSums& s) {
  const float2 t01 = T.ld2(0);
  const float2 c01 = P.ld2(0);
  if (t01.x != 1.0f) {
    // v1Loss.py:91 -- both slots of a cell without object: conf^2 against the untouched 0 target
    s.miss += c01.x * c01.x + c01.y * c01.y;
    if (HAS_GRAD) {
      G.st2(0, k.k2ln * c01.x, k.k2ln * c01.y);
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
    if (HAS_GRAD) G.st2(c, k.k2 * dx, k.k2 * dy);
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
  for (int d = 0; d < 4; ++d) gl[d] = coord_term(pr[d], gr[d], paper && d < 2, loc);
  s.loc += loc;
  if (HAS_GRAD) {
    float dI[4];
    iou_grad(r ? px1 : px0, gx, k.Sf, dI);
    float gv[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) gv[d] = (k.lc * gl[d] - 2.0f * dconf * dI[d]) * k.inv_bs;
    const float g_r = k.k2 * dconf, g_o = k.k2ln * conf_o;
    G.st2(0, r ? g_o : g_r, r ? g_r : g_o);
    G.st2(2, r ? 0.f : gv[0], r ? 0.f : gv[1]);
    G.st2(4, r ? 0.f : gv[2], r ? 0.f : gv[3]);
    G.st2(6, r ? gv[0] : 0.f, r ? gv[1] : 0.f);
    G.st2(8, r ? gv[2] : 0.f, r ? gv[3] : 0.f);
  }
  return true;
}

// ---- generic cell: any B <= 8, any C, any strides -------------------------------------------------
// FIX = false: streaming pass (square-root / paper form).  FIX = true: finalize pass for one of the
// call's first two object cells: writes only the 4 coordinate gradients of the responsible box in plain
// form and returns (plain - sqrt) of the location sum in s.loc.
template <bool HAS_GRAD, bool FIX, typename PA, typename TA, typename GA>
__device__ __forceinline__ bool cell_generic(const PA& P, const TA& T, const GA& G, const LossParams& k,
                                             CellSums& s) {
  const int B = k.B, C = k.C, D = 5 * B + C;
  if (T.ld(0) != 1.0f) {
    if (!FIX) {
      for (int b = 0; b < B; ++b) {
        const float cf = P.ld(b);
        s.miss += cf * cf;
        if (HAS_GRAD) G.st(b, k.k2ln * cf);
      }
      if (HAS_GRAD)
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
    const float gl = coord_term(pr[d], g, FIX || (paper && d < 2), loc);
    if (FIX) (void)coord_term(pr[d], g, false, loc_sqrt);
    if (HAS_GRAD) G.st(B + 4 * r + d, (k.lc * gl - 2.0f * dconf * dI[d]) * k.inv_bs);
  }
  s.loc += loc - loc_sqrt;
  if (!FIX && HAS_GRAD)
    for (int b = 0; b < B; ++b)
      if (b != r)
        for (int d = 0; d < 4; ++d) G.st(B + 4 * b + d, 0.f);
  return true;
}

template <typename E>
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
  __threadfence();
  if (warp == 0) {
    double t4[4] = {0, 0, 0, 0};
    for (unsigned int b = lane; b < gridDim.x; b += 32) {
#pragma unroll
      for (int t = 0; t < 4; ++t) t4[t] += __ldcg(&ws->partial[b][t]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) t4[t] = warp_sum(t4[t]);
    if (lane == 0) {
      const unsigned long long pr = *reinterpret_cast<volatile unsigned long long*>(&ws->pair);
      const uint32_t c[2] = {(uint32_t)(pr >> 32), (uint32_t)pr};
      const unsigned int carry = ws->carry;
      unsigned int seen = carry;
      for (int t = 0; t < 2; ++t) {
        if (c[t] == 0u) break;
        if (seen < 2 && p.coord_mode == YOLO1_COORD_REFERENCE) {
          // v1Loss.py:101 `[:2]`: this object is one of the first two of the call -> plain form
          const int64_t q = (int64_t)(0xFFFFFFFFu - c[t]);
          const E* zq = reinterpret_cast<const E*>(p.pred) + cell_offset<E>(p.ps, q, p.S);
          GlobIn<E> P{zq, p.ps[3], p.logits != 0};
          GlobIn<float> T{p.target + cell_offset<float>(p.ts, q, p.S), p.ts[3], false};
          GlobOut<E> G{HAS_GRAD ? reinterpret_cast<E*>(p.grad) + cell_offset<E>(p.gs, q, p.S) : nullptr, p.gs[3], zq,
                       p.ps[3], p.logits != 0};
          CellSums d = {0.f, 0.f, 0.f, 0.f};
          cell_generic<HAS_GRAD, true>(P, T, G, p, d);
          t4[0] += (double)d.loc;
        }
        ++seen;
      }
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
}

// ---- K1 fast kernel: contiguous layout, TMA in / TMA out ----------------------------------------------
template <typename E, bool HAS_GRAD, int TILE, int STAGES, int NOUT, bool SIG = false>
__global__ void __launch_bounds__(TILE) loss_tma_kernel(const __grid_constant__ LossParams p) {
  constexpr int D = 30;
  constexpr uint32_t PB = TILE * D * sizeof(E), TB = TILE * D * sizeof(float), GB = PB;
  static_assert(PB % 16 == 0 && TB % 16 == 0, "bulk copies move multiples of 16 bytes");
  static_assert(NOUT == 0 || NOUT >= 2, "NOUT = 0: gradient tile overwrites the pred stage in place; else >= 2 buffers");
  constexpr bool INPLACE = NOUT == 0;
  extern __shared__ __align__(128) unsigned char smem[];
  E* sp = reinterpret_cast<E*>(smem);
  float* st = reinterpret_cast<float*>(smem + STAGES * PB);
  E* so = reinterpret_cast<E*>(smem + STAGES * (PB + TB));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (PB + TB) + NOUT * GB);

  const int tid = threadIdx.x;
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
    mbar_arrive_expect_tx(&bars[s], PB + TB);
    bulk_g2s(sp + s * (TILE * D), gp + off, PB, &bars[s], pol);
    bulk_g2s(st + s * (TILE * D), p.target + off, TB, &bars[s], pol);
  };
  if (tid == 0)
    for (int64_t k = 0; k < my_n && k < STAGES; ++k) issue(k);

  CellSums sums = {0.f, 0.f, 0.f, 0.f};
  uint32_t m1 = 0, m2 = 0;
  using PIn = typename SmemIn<E>::type;
  using GOut = typename SmemOut<E>::type;
  for (int64_t k = 0; k < my_n; ++k) {
    const int s = (int)(k % STAGES), o = INPLACE ? 0 : (int)(k % (NOUT > 0 ? NOUT : 1));
    mbar_wait(&bars[s], (uint32_t)((k / STAGES) & 1));
    const PIn P{sp + s * (TILE * D) + tid * D};
    const SmemInF32 T{st + s * (TILE * D) + tid * D};
    // in-place: every thread reads its own cell's 30 values before it overwrites them with the gradient
    E* gtile = INPLACE ? sp + s * (TILE * D) : so + o * (TILE * D);
    const GOut G{gtile + tid * D};
    bool obj;
    if (SIG)
      obj = cell_b2c20<HAS_GRAD>(SigIn<PIn>{P}, T, SigOut<GOut, PIn>{G, P}, p, sums);
    else
      obj = cell_b2c20<HAS_GRAD>(P, T, G, p, sums);
    if (obj) note_object(m1, m2, ((int64_t)blockIdx.x + k * gridDim.x) * TILE + tid);
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
    const GlobIn<float> T{p.target + q * D, 1, false};
    const GlobOut<E> G{HAS_GRAD ? gg + q * D : nullptr, 1, gp + q * D, 1, SIG};
    if (cell_generic<HAS_GRAD, false>(P, T, G, p, sums)) note_object(m1, m2, q);
  }
  block_epilogue<E, HAS_GRAD, true>(sums, m1, m2, p);
}

// ---- K1 fast kernel, channel-planar pred/grad: the backbone's permuted NCHW view (OriginResNet.py:189) ------
// pred / grad are [N][30][S*S] in memory (element strides (30 S^2, S, 1, S^2)), target is contiguous NHWC.
// An image's 30 planes are one contiguous block, so a tile of `tile_imgs` whole images still moves with one
// bulk copy per tensor; one thread per cell reads its channels S*S elements apart (conflict-free) and writes
// the gradient tile in the same planar layout, so `permute`'s backward stays a free view.
template <typename E, bool HAS_GRAD, int STAGES, int NOUT, bool SIG = false>
__global__ void __launch_bounds__(256) loss_tma_planar_kernel(const __grid_constant__ LossParams p, int tile_imgs) {
  constexpr int D = 30;
  const int SS = p.S * p.S, tile_cells = tile_imgs * SS, tile_elems = tile_cells * D;
  const uint32_t PB = tile_elems * sizeof(E), TB = tile_elems * sizeof(float), GB = PB;
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
    mbar_arrive_expect_tx(&bars[s], PB + TB);
    bulk_g2s(sp + s * tile_elems, gp + off, PB, &bars[s], pol);
    bulk_g2s(st + s * tile_elems, p.target + off, TB, &bars[s], pol);
  };
  if (tid == 0)
    for (int64_t k = 0; k < my_n && k < STAGES; ++k) issue(k);

  CellSums sums = {0.f, 0.f, 0.f, 0.f};
  uint32_t m1 = 0, m2 = 0;
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
      if (SIG)
        obj = cell_b2c20<HAS_GRAD>(SigIn<PlanarIn<E>>{P}, T, SigOut<PlanarOut<E>, PlanarIn<E>>{G, P}, p, sums);
      else
        obj = cell_b2c20<HAS_GRAD>(P, T, G, p, sums);
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
    const E* zq = gp + cell_offset<E>(p.ps, q, p.S);
    const GlobIn<E> P{zq, p.ps[3], SIG};
    const GlobIn<float> T{p.target + cell_offset<float>(p.ts, q, p.S), p.ts[3], false};
    const GlobOut<E> G{HAS_GRAD ? gg + cell_offset<E>(p.gs, q, p.S) : nullptr, p.gs[3], zq, p.ps[3], SIG};
    if (cell_generic<HAS_GRAD, false>(P, T, G, p, sums)) note_object(m1, m2, q);
  }
  block_epilogue<E, HAS_GRAD, true>(sums, m1, m2, p);
}

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
    const E* zq = reinterpret_cast<const E*>(p.pred) + cell_offset<E>(p.ps, q, p.S);
    const GlobIn<E> P{zq, p.ps[3], p.logits != 0};
    const GlobIn<float> T{p.target + cell_offset<float>(p.ts, q, p.S), p.ts[3], false};
    const GlobOut<E> G{HAS_GRAD ? reinterpret_cast<E*>(p.grad) + cell_offset<E>(p.gs, q, p.S) : nullptr, p.gs[3], zq,
                       p.ps[3], p.logits != 0};
    if (cell_generic<HAS_GRAD, false>(P, T, G, p, sums)) note_object(m1, m2, q);
  }
  block_epilogue<E, HAS_GRAD, false>(sums, m1, m2, p);  // no bulk stores in flight here
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

template <typename E, bool HAS_GRAD, int TILE, int STAGES, int NOUT, bool SIG = false>
int launch_tma(const LossParams& p, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * TILE * 30 * (sizeof(E) + 4) + (size_t)NOUT * TILE * 30 * sizeof(E) +
                          STAGES * sizeof(uint64_t);
  auto kern = loss_tma_kernel<E, HAS_GRAD, TILE, STAGES, NOUT, SIG>;
  YOLO1_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = kNumSMs, per_sm = 1;
  YOLO1_CUDA_TRY(cudaGetDevice(&dev));
  YOLO1_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  YOLO1_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TILE, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t tiles = p.cells / TILE;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > tiles) grid = tiles;
  if (grid > kMaxGrid) grid = kMaxGrid;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, TILE, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

template <typename E, bool HAS_GRAD>
int launch_tma_variant(const LossParams& p, int variant, cudaStream_t stream) {
  if (p.logits) return launch_tma<E, HAS_GRAD, 128, 2, 2, true>(p, stream);  // one launch shape with the fused head
  switch (variant) {
    case 0:
    case 1: return launch_tma<E, HAS_GRAD, 128, 2, 2>(p, stream);
    case 2: return launch_tma<E, HAS_GRAD, 128, 3, 2>(p, stream);
    case 3: return launch_tma<E, HAS_GRAD, 64, 3, 2>(p, stream);
    case 4: return launch_tma<E, HAS_GRAD, 64, 4, 3>(p, stream);
    case 5: return launch_tma<E, HAS_GRAD, 256, 2, 2>(p, stream);
    case 6: return launch_tma<E, HAS_GRAD, 32, 4, 2>(p, stream);
    case 7: return launch_tma<E, HAS_GRAD, 32, 6, 3>(p, stream);
    case 8: return launch_tma<E, HAS_GRAD, 128, 2, 0>(p, stream);   // in-place gradient tile: 61 KB, 3 CTAs/SM
    case 9: return launch_tma<E, HAS_GRAD, 128, 3, 0>(p, stream);   // 92 KB, 2 CTAs/SM
    case 10: return launch_tma<E, HAS_GRAD, 192, 2, 0>(p, stream);  // 92 KB, 2 CTAs/SM
    case 11: return launch_tma<E, HAS_GRAD, 96, 2, 0>(p, stream);   // 46 KB, 4 CTAs/SM
    case 12: return launch_tma<E, HAS_GRAD, 64, 2, 0>(p, stream);   // 31 KB, 7 CTAs/SM
    case 13: return launch_tma<E, HAS_GRAD, 256, 2, 0>(p, stream);  // 123 KB, 1 CTA/SM
    default: return YOLO1_ERR_ARG;
  }
}

bool planar(const int64_t st[4], int S, int D) {
  return st[3] == (int64_t)S * S && st[2] == 1 && st[1] == S && st[0] == (int64_t)S * S * D;
}

// whole images per tile so that both tiles are multiples of 16 bytes and hold 128..256 cells; 0 = no fit
int planar_tile_imgs(int S, size_t esz, int target_cells) {
  const int SS = S * S;
  int m = 0;
  for (int k = 1; k <= 16; ++k)
    if (((size_t)k * SS * 30 * esz) % 16 == 0 && ((size_t)k * SS * 120) % 16 == 0) {
      m = k;
      break;
    }
  if (m == 0 || m * SS > 256) return 0;
  int t = m;
  while ((t + m) * SS <= target_cells) t += m;
  return t;
}

template <typename E, bool HAS_GRAD, int NOUT, bool SIG>
int launch_planar_n(const LossParams& p, int tile_imgs, cudaStream_t stream) {
  constexpr int STAGES = 2;
  const int tile_cells = tile_imgs * p.S * p.S;
  const size_t smem = (size_t)STAGES * tile_cells * 30 * (sizeof(E) + 4) + (size_t)NOUT * tile_cells * 30 * sizeof(E) +
                      STAGES * sizeof(uint64_t);
  const int threads = (tile_cells + 31) / 32 * 32;
  auto kern = loss_tma_planar_kernel<E, HAS_GRAD, STAGES, NOUT, SIG>;
  YOLO1_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = kNumSMs, per_sm = 1;
  YOLO1_CUDA_TRY(cudaGetDevice(&dev));
  YOLO1_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  YOLO1_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t tiles = p.cells / tile_cells;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > tiles) grid = tiles;
  if (grid > kMaxGrid) grid = kMaxGrid;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, threads, smem, stream>>>(p, tile_imgs);
  return (int)cudaGetLastError();
}

// Two CTAs per SM are what keeps the copy engine busy (tools/tune_loss.py): separate output buffers while the
// tile is small enough for that (<= 110 KB per CTA), gradient written in place over the pred stage otherwise.
template <typename E, bool HAS_GRAD>
int launch_planar(const LossParams& p, int tile_imgs, cudaStream_t stream) {
  const size_t tile_cells = (size_t)tile_imgs * p.S * p.S;
  const size_t separate = 2 * tile_cells * 30 * (sizeof(E) + 4) + 2 * tile_cells * 30 * sizeof(E);
  if (p.logits) {
    if (separate <= 110 * 1024) return launch_planar_n<E, HAS_GRAD, 2, true>(p, tile_imgs, stream);
    return launch_planar_n<E, HAS_GRAD, 0, true>(p, tile_imgs, stream);
  }
  if (separate <= 110 * 1024) return launch_planar_n<E, HAS_GRAD, 2, false>(p, tile_imgs, stream);
  return launch_planar_n<E, HAS_GRAD, 0, false>(p, tile_imgs, stream);
}

template <bool HAS_GRAD>
int launch_hostmapped(const LossParams& p, cudaStream_t stream) {
  constexpr int TILE = 128;
  constexpr size_t smem = 2 * TILE * 30 * sizeof(float);
  auto kern = loss_hostmapped_kernel<HAS_GRAD, TILE>;
  int dev = 0, sms = kNumSMs;
  YOLO1_CUDA_TRY(cudaGetDevice(&dev));
  YOLO1_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
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
  int dev = 0, sms = kNumSMs;
  YOLO1_CUDA_TRY(cudaGetDevice(&dev));
  YOLO1_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
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
int loss_launch_chunk(const void* pred, const int64_t ps[4], int pred_dtype, const float* target,
                      const int64_t ts[4], void* grad, const int64_t gs[4], float* terms, int64_t N, int S,
                      int B, int C, float lambda_coord, float lambda_noobj, float inv_batch_size, int coord_mode,
                      void* workspace, size_t workspace_bytes, int chunk_flags, int variant, cudaStream_t stream) {
  if (!terms || !workspace || !ps || !ts || N < 0) return YOLO1_ERR_ARG;
  if (N > 0 && (!pred || !target)) return YOLO1_ERR_ARG;  // an empty batch may come with null data pointers
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

  LossParams p;
  p.pred = pred, p.target = target, p.grad = grad, p.terms = terms, p.ws = reinterpret_cast<LossWs*>(workspace);
  for (int d = 0; d < 4; ++d) p.ps[d] = ps[d], p.ts[d] = ts[d], p.gs[d] = grad ? gs[d] : 0;
  p.cells = cells, p.S = S, p.B = B, p.C = C;
  p.Sf = (float)S, p.lc = lambda_coord, p.ln = lambda_noobj, p.inv_bs = inv_batch_size;
  p.k2ln = 2.0f * lambda_noobj * inv_batch_size, p.k2 = 2.0f * inv_batch_size;
  p.coord_mode = coord_mode, p.last_chunk = (chunk_flags & 2) ? 1 : 0, p.logits = (chunk_flags & 4) ? 1 : 0;

  if (chunk_flags & 1) YOLO1_CUDA_TRY(cudaMemsetAsync(workspace, 0, offsetof(LossWs, partial), stream));

  const int D = 5 * B + C;
  const bool fast = variant >= 0 && B == 2 && C == 20 && contiguous(ps, S, D) && contiguous(ts, S, D) &&
                    (!grad || contiguous(gs, S, D)) && (uintptr_t)pred % 16 == 0 && (uintptr_t)target % 16 == 0 &&
                    (!grad || (uintptr_t)grad % 16 == 0);
  const bool bf = pred_dtype == YOLO1_DTYPE_BF16;
  if (variant == kVariantHostMapped) {  // pointers are device-visible HOST memory (host_ctx.cu)
    if (!fast || bf || p.logits) return YOLO1_ERR_UNSUPPORTED;
    return grad ? launch_hostmapped<true>(p, stream) : launch_hostmapped<false>(p, stream);
  }
  if (fast) {
    if (bf) return grad ? launch_tma_variant<__nv_bfloat16, true>(p, variant, stream)
                        : launch_tma_variant<__nv_bfloat16, false>(p, variant, stream);
    return grad ? launch_tma_variant<float, true>(p, variant, stream)
                : launch_tma_variant<float, false>(p, variant, stream);
  }
  const int tile_imgs = planar_tile_imgs(S, esz, variant == 1 ? 224 : 128);
  const bool fast_planar = variant >= 0 && B == 2 && C == 20 && tile_imgs > 0 && planar(ps, S, D) &&
                           contiguous(ts, S, D) && (!grad || planar(gs, S, D)) && (uintptr_t)pred % 16 == 0 &&
                           (uintptr_t)target % 16 == 0 && (!grad || (uintptr_t)grad % 16 == 0);
  if (fast_planar) {
    if (bf) return grad ? launch_planar<__nv_bfloat16, true>(p, tile_imgs, stream)
                        : launch_planar<__nv_bfloat16, false>(p, tile_imgs, stream);
    return grad ? launch_planar<float, true>(p, tile_imgs, stream) : launch_planar<float, false>(p, tile_imgs, stream);
  }
  if (bf) return grad ? launch_generic<__nv_bfloat16, true>(p, stream) : launch_generic<__nv_bfloat16, false>(p, stream);
  return grad ? launch_generic<float, true>(p, stream) : launch_generic<float, false>(p, stream);
}

}  // namespace yolo1

extern "C" {

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
