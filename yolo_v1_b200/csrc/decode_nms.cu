// decode_nms.cu -- K2 (decode) and K3 (bitmask NMS), and their per-image fusion (sm_100a).
//
// Replaces decoder (reference utils/utils.py:94-147) and nms (utils/utils.py:150-184).
// Specification: SURVEY.md Appendix B.  Built with -fmad=false: every fp32 operation rounds on its own,
// exactly like the reference's ATen CPU ops, so candidates, scores and keep lists are bit-exact.
//
// One CTA per image; an image's candidates never leave shared memory in the fused kernel:
//   load   : the image's S*S*D values -> shared (dense fp32: one bulk (TMA) copy per image; else coalesced loads)
//   decode : per (cell, slot) candidate test, class arg-max (the two slots of a cell share the scan), score,
//            `double(score) > thresh`, order-preserving compaction (ballot + prefix) = emission order; the image's
//            maximum confidence rides on the same barrier (it only matters when no confidence exceeds 1e-4)
//   sort   : exact rank (score descending, emission index ascending).  Large grids: a counting sort into 256 score
//            buckets and a comparison inside the bucket only, O(n); 7x7 grids: rank by counting (no dependent chain)
//   mask   : suppression bit-matrix.  IoU <= min(area) / max(area), so a pair whose areas differ by more than
//            1 / thr cannot die: a second counting sort (area buckets, 8 per octave) puts every box next to the only
//            partners that can matter -- a third of all pairs at thr = 0.5 -- and those are dealt out to the threads
//            in equal shares; an fp32 pre-test settles the survivors, the few others take the exact test
//            (dies iff !(IoU <= thr), utils/utils.py:180; the image tile is reused for the matrix)
//   sweep  : the kept set as the fixed point of K(j) = no i < j with K(i) and M[i][j], a few passes of bit
//            operations.  A suppressed box never suppresses (iterated semantics of the reference).
//   store  : kept detections in descending score order (float4 boxes), counts.
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace yolo1 {
namespace {

// Unroll factors of the hot loops, from measurement (tools/tune_decode.py on B200, M images/s at S=7 N=65536 uniform /
// sigmoid inputs, S=14 N=16384): the compiler's own choice for the rank loop 133.7 / 166.7; rank x4 148.8 / 187.1;
// plus the deferred pair loop rolled (x1) 151.7 / 192.8 (x4: 148.8); the on-the-spot pair loop x4 11.73 at S=14
// (rolled 11.18).  The kernel is sensitive to its code size (ncu: no_instruction stalls), not only to its
// instruction count.
constexpr int kMaxCand = 1024;
#ifndef YOLO1_DECODE_PERSIST
#define YOLO1_DECODE_PERSIST 0
#endif
#ifndef YOLO1_DECODE_PERSIST_SPLIT
#define YOLO1_DECODE_PERSIST_SPLIT 1
#endif
#ifndef YOLO1_DECODE_L2PF
#define YOLO1_DECODE_L2PF 0
#endif
constexpr int kKeepA = 64, kKeepB = 128;            // slots of Smem::misc (sweep): the two copies of the kept-set words
constexpr int kNotes = 96, kKmin = 97, kKmax = 98, kRankSum = 99, kBar = 100;   // ... (nms_phase): unsettled-pair count, score-key
                                                                      // range, rank checksum
constexpr int kRankUnroll = 4, kPairUnroll = 1;      // measured in round 1 (tools/tune_decode.py): the small-grid kernel is
                                                     // sensitive to its code size, not only to its instruction count
constexpr int kNB = 256;                             // buckets of the two counting sorts (16-bit counters, two per word)
constexpr int kAreaKey0 = (127 - 24) << 3;           // area bucket 0 starts at 2^-24; 8 buckets per octave, 32 octaves

struct DecodeParams {
  const void* pred;
  int64_t st[4];
  int S, B, C;
  float cs;       // fl32(1/S), utils/utils.py:105
  double thresh;  // python double, utils/utils.py:129
  float iou_thr;
  int per_class;
  // division-free form of `fl32(inter / u) <= thr` (see set_threshold, iou_exceeds): thr_mid = midpoint between
  // thr and the next float above it (exact in double), nudged up by two double ulps when that midpoint itself
  // rounds down to thr (thr_tie_ok, round-to-nearest-even); thr_fast = 0 when thr is outside the proven range.
  double thr_mid;
  int thr_tie_ok;
  int thr_fast;
  // fp32 pre-test of the pair loop (pair_margin): thr_lo = thr (1 - 2^-18) rounded down, thr_k = 1 + thr_lo rounded up;
  // thr_lo = 0 (nothing is ever settled by the pre-test) unless 2^-20 <= thr <= 4
  float thr_lo, thr_k;
  // area pruning (nms_phase): boxes whose area buckets are at least `area_skip` apart have areas more than 1 / thr
  // apart, so their IoU cannot exceed thr (set_threshold)
  int area_skip;
  // bulk image load (decode_nms_image): dense fp32 [N,S,S,D] tensor on a 16-byte aligned base whose images are a
  // multiple of 8 bytes -- one cp.async.bulk per image instead of a load / store pair per 8 bytes
  int bulk_ok;
  int64_t n_images;
  int s_magic;    // cell / S == (cell * s_magic) >> 16 for every cell of the grid (0: use the division)
  // decode outputs / nms inputs
  float* boxes;
  float* scores;
  int32_t* cls;
  int32_t* counts;
  // nms / fused outputs
  float* out_boxes;
  float* out_scores;
  int32_t* out_cls;
  int32_t* keep;
  int32_t* out_counts;
  int32_t* cand_counts;
  int max_n;
};

// dynamic shared memory layout shared by the three kernels
struct Smem {
  float* img;       // [S*S*D]   (decode)   -- the region is reused after decode by the five arrays below
  uint32_t* mask;   // [n * W]
  int32_t* bidx;    // [max_n]  candidate indices in score-bucket order
  uint32_t* notes;  // [2 * max_n]  pairs the fp32 pre-test left unsettled: (area position a) << 16 | (area position b)
  uint32_t* tkey;   // [max_n]  per candidate: (score bucket << 16 | slot in it), later its rank
  uint32_t* akey;   // [max_n]  per candidate: (area bucket << 16 | slot in it), later its area position
  float4* box;      // [max_n]  candidates in emission order
  float* score;     // [max_n]
  int32_t* cls;     // [max_n]
  float4* abox;     // [max_n]  boxes in area-bucket order         } together exactly the bytes of sbox: the general
  float* ata;       // [max_n]  thr_lo * area, same order           } pair code (wild images) rebuilds sbox over them
  int32_t* arank;   // [max_n]  score rank of the box at an area position
  float4* sbox;     // [max_n + max_n/2]  sorted by score, sbox[n + r] = sbox[r] for r < n/2 (general pair code only)
  int32_t* lpre;    // [max_n + 1]  partners per area position, then their exclusive prefix (lpre[n] = total)
  int32_t* sidx;    // [max_n]  sorted position -> emission index
  int32_t* keep;    // [max_n]  kept sorted positions
  uint32_t* hist;   // [kNB]  kNB / 2 words of score-bucket counters, then kNB / 2 of area-bucket counters
  int32_t* misc;    // [160]  0..63 decode scratch / kept count; 64..95 and 128..159 kept-set words of the sweep
                    //        (128..159 decode scratch before); 96..98 note count and score-key range
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// split: the persistent small-grid kernel (decode_nms_persistent_kernel) loads image k+1 while image k is in its
// NMS phase, so the arrays of that phase must not lie over the image: mask and akey follow the image region, tkey
// takes the score-bucket half of `hist` (unused by the small-grid phases; 512 bytes >= 4 max_n for max_n <= 128), and
// bidx / notes (large-grid phases only) do not exist.
__host__ __device__ inline size_t smem_layout(unsigned char* base, int img_floats, int max_n, Smem* s,
                                              bool split = false) {
  const int W = (max_n + 31) / 32;
  const size_t mask_bytes = align16((size_t)max_n * W * 4), n4 = align16((size_t)max_n * 4);
  size_t region = (size_t)img_floats * 4 + (img_floats ? 16 : 0);   // + the bulk load's alignment slack
  const size_t after = mask_bytes + n4 /*bidx*/ + 2 * n4 /*notes*/ + n4 /*tkey*/ + n4 /*akey*/;
  if (!split && after > region) region = after;
  size_t off = align16(region);
  if (s) {
    s->img = reinterpret_cast<float*>(base);
    if (split) {
      s->mask = reinterpret_cast<uint32_t*>(base + off);
      s->akey = reinterpret_cast<uint32_t*>(base + off + mask_bytes);
      s->bidx = nullptr, s->notes = nullptr;
    } else {
      s->mask = reinterpret_cast<uint32_t*>(base);
      s->bidx = reinterpret_cast<int32_t*>(base + mask_bytes);
      s->notes = reinterpret_cast<uint32_t*>(base + mask_bytes + n4);
      s->tkey = reinterpret_cast<uint32_t*>(base + mask_bytes + 3 * n4);
      s->akey = reinterpret_cast<uint32_t*>(base + mask_bytes + 4 * n4);
    }
  }
  if (split) off += mask_bytes + n4;
  if (s) {
    s->sbox = reinterpret_cast<float4*>(base + off);
    s->abox = reinterpret_cast<float4*>(base + off);
    s->ata = reinterpret_cast<float*>(base + off + (size_t)max_n * 16);
    s->arank = reinterpret_cast<int32_t*>(base + off + (size_t)max_n * 20);
  }
  off += align16((size_t)(max_n + max_n / 2 + 1) * 16);   // >= 24 max_n: abox + ata + arank
  if (s) s->box = reinterpret_cast<float4*>(base + off);
  off += (size_t)max_n * 16;
  if (s) s->score = reinterpret_cast<float*>(base + off);
  off += n4;
  if (s) s->cls = reinterpret_cast<int32_t*>(base + off);
  off += n4;
  if (s) s->lpre = reinterpret_cast<int32_t*>(base + off);
  off += align16((size_t)(max_n + 1) * 4);
  if (s) s->sidx = reinterpret_cast<int32_t*>(base + off);
  off += n4;
  if (s) s->keep = reinterpret_cast<int32_t*>(base + off);
  off += n4;
  if (s) s->hist = reinterpret_cast<uint32_t*>(base + off);
  if (s && split) s->tkey = reinterpret_cast<uint32_t*>(base + off);
  off += kNB * 4;
  if (s) s->misc = reinterpret_cast<int32_t*>(base + off);
  off += 160 * 4;
  return off;
}

// ---- phase: load one image into shared memory as [cell][channel] floats ---------------------------------
template <typename E>
__device__ __forceinline__ void load_image(const DecodeParams& p, int64_t n, float* img) {
  const int S = p.S, D = 5 * p.B + p.C, total = S * S * D;
  const E* base = reinterpret_cast<const E*>(p.pred) + n * p.st[0];
  const bool dense = p.st[3] == 1 && p.st[2] == D && p.st[1] == (int64_t)S * D;
  if (dense && sizeof(E) == 4 && !(total & 1) && !((uintptr_t)base & 7)) {
    // fp32, 8-byte aligned image: 64-bit loads, all of a thread's loads in flight before the first store
    // (one trip to memory per CTA instead of one per unrolled group)
    const float2* b2 = reinterpret_cast<const float2*>(base);
    float2* i2 = reinterpret_cast<float2*>(img);
    const int n2 = total >> 1;
    for (int t0 = threadIdx.x; t0 < n2; t0 += 8 * blockDim.x) {
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = t0 + u * blockDim.x;
        if (t < n2) v[u] = __ldg(b2 + t);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = t0 + u * blockDim.x;
        if (t < n2) i2[t] = v[u];
      }
    }
  } else if (dense) {
    for (int t0 = threadIdx.x; t0 < total; t0 += 8 * blockDim.x) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = t0 + u * blockDim.x;
        if (t < total) v[u] = ld_elem(base + t);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int t = t0 + u * blockDim.x;
        if (t < total) img[t] = v[u];
      }
    }
  } else {
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
      const int cell = t / D, c = t - cell * D;
      const int i = cell / S, j = cell - i * S;
      img[t] = ld_elem(base + i * p.st[1] + j * p.st[2] + c * p.st[3]);
    }
  }
}

// ATen's max is NaN-propagating (`contain.max()`, utils/utils.py:113; torch.max(dim), :127): FMNMX.NAN, and the
// warp-wide form as one CREDUX.MAX.F32.NAN (sm_100a `redux.sync` on floats) instead of five shuffle rounds
__device__ __forceinline__ float fmax_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float warp_max_nan(float v) {
  float r;
  asm volatile("redux.sync.max.NaN.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// slot index t = cell * B + b, cell = i * S + j, without runtime divisions in the usual case (B = 2; s_magic)
__device__ __forceinline__ void split_slot(const DecodeParams& p, int t, int& cell, int& b) {
  cell = p.B == 2 ? (t >> 1) : t / p.B;
  b = t - cell * p.B;
}
__device__ __forceinline__ void split_cell(const DecodeParams& p, int cell, int& i, int& j) {
  i = p.s_magic ? (int)(((unsigned)cell * (unsigned)p.s_magic) >> 16) : cell / p.S;
  j = cell - i * p.S;
}

// ---- phase: decode (utils/utils.py:108-132).  Returns the candidate count (uniform over the CTA). -----
struct SlotEval {   // what a slot carries across the barrier of its pass; the box is computed after it (slot_box)
  float score;
  int cls;
  bool pass;
};

// One (cell, slot) of the image.  PAIRED (B == 2): the two slots of a cell sit in neighbouring lanes, so each lane
// scans half of the class scores and the halves are merged with one shuffle (every lane of the warp must call).
template <bool PAIRED>
__device__ __forceinline__ SlotEval eval_slot(const DecodeParams& p, const Smem& sm, int t, int slots, float mx,
                                             bool use_mx) {
  const int B = p.B, C = p.C, D = 5 * B + C;
  SlotEval r = {0.f, 0, false};
  const bool live = t < slots;
  int cell = 0, b = 0;
  if (live) split_slot(p, t, cell, b);
  const float* P = sm.img + cell * D;
  float best_p;
  int best_c;
  if (PAIRED) {
    // :127 first arg-max.  Lane b = 0 scans classes [0, C/2) starting from the first score like the reference's
    // scan; lane b = 1 scans [C/2, C) from -inf; the upper half wins only if it is strictly larger.  torch.max(dim)
    // propagates NaN: the value is carried by a NaN-propagating max (a NaN class score makes the slot's score NaN and
    // `> thresh` drops it, so the index of such a slot is never used); the index follows the `>` of finite scans.
    const int half = (C + 1) >> 1, c0 = b ? half : 1, c1 = b ? C : half;
    if ((C & 3) == 0 && !((5 * B) & 1)) {
      // two class scores per 64-bit shared load (both halves start at an even channel: 8-byte aligned).  Starting
      // from -inf with a strict `>` keeps the first arg-max: an all -inf half keeps its first index.
      const float2* P2 = reinterpret_cast<const float2*>(P + 5 * B + (b ? half : 0));
      best_p = -INFINITY;
      best_c = b ? half : 0;
      for (int j = 0; j < (half >> 1); ++j) {
        const float2 v = P2[j];
        const int c = (b ? half : 0) + 2 * j;
        if (v.x > best_p) best_c = c;
        best_p = fmax_nan(best_p, v.x);
        if (v.y > best_p) best_c = c + 1;
        best_p = fmax_nan(best_p, v.y);
      }
    } else {
      best_p = b ? -INFINITY : P[5 * B];
      best_c = b ? half : 0;
      for (int c = c0; c < c1; ++c) {
        const float v = P[5 * B + c];
        if (v > best_p) best_c = c;
        best_p = fmax_nan(best_p, v);
      }
    }
    const float op = __shfl_xor_sync(0xffffffffu, best_p, 1);
    const int oc = __shfl_xor_sync(0xffffffffu, best_c, 1);
    const float lo_p = b ? op : best_p, hi_p = b ? best_p : op;
    const int lo_c = b ? oc : best_c, hi_c = b ? best_c : oc;
    best_c = hi_p > lo_p ? hi_c : lo_c;
    best_p = fmax_nan(hi_p, lo_p);
  } else {
    best_p = P[5 * B], best_c = 0;
    for (int c = 1; c < C; ++c) {
      const float v = P[5 * B + c];
      if (v > best_p) best_c = c;
      best_p = fmax_nan(best_p, v);
    }
  }
  const float conf = P[b];
  if (live && (conf > 0.0001f || (use_mx && conf == mx))) {  // :108-114
    r.score = conf * best_p;                      // :129
    r.cls = best_c;
    r.pass = (double)r.score > p.thresh;
  }
  return r;
}

// :119-126 the box of slot t, cell-relative xywh -> image xyxy
__device__ __forceinline__ float4 slot_box(const DecodeParams& p, const Smem& sm, int t) {
  int cell, b, i, j;
  split_slot(p, t, cell, b);
  split_cell(p, cell, i, j);
  const float* Q = sm.img + cell * (5 * p.B + p.C) + p.B + 4 * b;
  const float x = Q[0], y = Q[1], w = Q[2], h = Q[3];
  const float cx = x * p.cs + (float)j * p.cs;  // :122-123 (no FMA: file is built with -fmad=false)
  const float cy = y * p.cs + (float)i * p.cs;
  const float hw = 0.5f * w, hh = 0.5f * h;
  return make_float4(cx - hw, cy - hh, cx + hw, cy + hh);
}

template <bool PAIRED, bool TWO>
__device__ __forceinline__ int decode_phase_impl(const DecodeParams& p, const Smem& sm) {
  const int S = p.S, B = p.B, D = 5 * B + p.C, slots = S * S * B;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float* red = reinterpret_cast<float*>(sm.misc);  // [32] floats
  int* wcount = sm.misc + 32;                      // [32] ints: first round of a pass; wcount2: second round
  int* wcount2 = sm.misc + 128;
  // :109-113 max over the confidences of the image.  It only matters when no confidence exceeds 0.0001 (a slot at
  // or below that is a candidate iff it equals the maximum), so the per-warp maxima travel with the first pass's
  // counts through the same barrier, the pass is evaluated without the rule, and it is repeated with the rule in
  // the rare image that needs it.
  float mx = -INFINITY;
  for (int t = threadIdx.x; t < slots; t += blockDim.x) {
    int cell, b;
    split_slot(p, t, cell, b);
    mx = fmax_nan(mx, sm.img[cell * D + b]);   // one NaN confidence: the maximum is NaN and `== max` selects nothing
  }
  mx = warp_max_nan(mx);
  if (lane == 0) red[warp] = mx;
  bool have_mx = false;

  // TWO: two rounds of slots per pass (slot t and slot t + blockDim).  With 98 slots on 96 threads the second round
  // is two slots in the first warp, and one pass means one barrier pair and one prefix.  (Not for the large-grid
  // kernel: carrying the second slot across the barrier costs it 4 registers and with them a resident CTA.)
  int base = 0;  // candidates emitted by earlier passes (uniform)
  const int stride = (TWO ? 2 : 1) * blockDim.x;
  for (int t0 = 0; t0 < slots;) {
    const int t = t0 + threadIdx.x, t2 = t + blockDim.x;
    const SlotEval e1 = eval_slot<PAIRED>(p, sm, t, slots, mx, have_mx);
    SlotEval e2 = {0.f, 0, false};
    if (TWO && t2 - lane < slots) e2 = eval_slot<PAIRED>(p, sm, t2, slots, mx, have_mx);   // warp-uniform condition
    const unsigned bal1 = __ballot_sync(0xffffffffu, e1.pass), bal2 = TWO ? __ballot_sync(0xffffffffu, e2.pass) : 0u;
    if (lane == 0) {
      wcount[warp] = __popc(bal1);
      if (TWO) wcount2[warp] = __popc(bal2);
    }
    __syncthreads();
    if (!have_mx) {
      mx = warp_max_nan(lane < nwarps ? red[lane] : -INFINITY);
      have_mx = true;
      if (!(mx > 0.0001f)) continue;   // uniform: this pass again, now with the `== max` rule
    }
    const int v1 = lane < nwarps ? wcount[lane] : 0, v2 = (TWO && lane < nwarps) ? wcount2[lane] : 0;
    const int total1 = __reduce_add_sync(0xffffffffu, v1), total2 = TWO ? __reduce_add_sync(0xffffffffu, v2) : 0;
    const int before1 = base + __reduce_add_sync(0xffffffffu, lane < warp ? v1 : 0);
    const unsigned lower = (1u << lane) - 1u;
    if (e1.pass) {
      const int k = before1 + __popc(bal1 & lower);
      sm.box[k] = slot_box(p, sm, t), sm.score[k] = e1.score, sm.cls[k] = e1.cls;
    }
    if (TWO) {
      const int before2 = base + total1 + __reduce_add_sync(0xffffffffu, lane < warp ? v2 : 0);
      if (e2.pass) {
        const int k = before2 + __popc(bal2 & lower);
        sm.box[k] = slot_box(p, sm, t2), sm.score[k] = e2.score, sm.cls[k] = e2.cls;
      }
    }
    base += total1 + total2;
    t0 += stride;
    __syncthreads();
  }
  return base;
}

template <bool TWO>
__device__ __forceinline__ int decode_phase(const DecodeParams& p, const Smem& sm) {
  // lanes 2m, 2m+1 hold the two slots of one cell when B == 2 (slot index and CTA size are even)
  if (p.B == 2 && !(blockDim.x & 1)) return decode_phase_impl<true, TWO>(p, sm);
  return decode_phase_impl<false, TWO>(p, sm);
}

// ---- phase: rank sort + suppression matrix + sweep (utils/utils.py:150-184).  Returns kept count. --------
// `!(fl32(inter / u) <= thr)` without the division.  fl32() is monotone, so fl32(x) <= thr  <=>  x < mid, or
// x == mid when mid rounds to thr, where mid is the midpoint between thr and its successor.  For u > 0 that is
// inter < mid * u (or <=); inter and u carry 24 significant bits and mid 25, so the product is exact in double.
// The `<=` case is folded into the constant: inter and mid * u are both multiples of 2^(e_mid + e_u - 47), so two
// distinct values differ by more than 2^-49 relative; thr_mid = mid + 2 ulp_double (set_threshold) gives
// fl64(thr_mid * u) in (mid u, mid u (1 + 2^-50)]: `<=` becomes `<` and no other case moves.
// The test holds for every u > 0 including +inf (inter / inf = 0 survives; inf / u dies; NaN inter dies);
// u <= 0 or NaN (0/0, negative areas) and thresholds outside the proven range take the real IEEE division.
__device__ __forceinline__ bool iou_exceeds(float inter, float u, const DecodeParams& p) {
  if (p.thr_fast && u > 0.f) return !((double)inter < p.thr_mid * (double)u);
  return !(inter / u <= p.iou_thr);
}

// ---- the pairs ---------------------------------------------------------------------------------------------
// intersection and union term of box A (area_a) with box Bx, by the reference's op sequence (:166-179).
// FINITE = true: every coordinate of the image is an ordinary number (|c| < 1e18) and every area lies in
// [1e-30, 1e30], so clamp(min=)/clamp(max=) are plain max/min (one FMNMX each), the intersection is symmetric in
// the two boxes, nothing overflows or underflows, and the fp32 pre-test (pair_margin) applies.
template <bool FINITE>
__device__ __forceinline__ void pair_terms(const float4& A, float area_a, const float4& Bx, bool a_first, float& inter,
                                           float& u) {
  const float area_b = (Bx.z - Bx.x) * (Bx.w - Bx.y);   // :159 area of box j
  float ww, hh;
  if (FINITE) {
    ww = fmaxf(fminf(Bx.z, A.z) - fmaxf(Bx.x, A.x), 0.f);
    hh = fmaxf(fminf(Bx.w, A.w) - fmaxf(Bx.y, A.y), 0.f);
  } else {
    const float4 L = a_first ? A : Bx, R = a_first ? Bx : A;   // L: the earlier box, the reference's box i
    const float xx1 = R.x < L.x ? L.x : R.x;  // clamp(min=x1[i])
    const float yy1 = R.y < L.y ? L.y : R.y;
    const float xx2 = R.z > L.z ? L.z : R.z;  // clamp(max=x2[i])
    const float yy2 = R.w > L.w ? L.w : R.w;
    ww = xx2 - xx1, hh = yy2 - yy1;
    if (ww < 0.f) ww = 0.f;
    if (hh < 0.f) hh = 0.f;
  }
  inter = ww * hh;
  u = (area_a + area_b) - inter;   // ovr = inter / (a_i + a_j - inter)
}

// the later box of a dying pair gets the earlier one's bit: column hi (sorted position), bit lo -- "kept lo kills hi",
// which is what the sweep reads
__device__ __forceinline__ void set_kill(const Smem& sm, int W, int ra, int rb, const DecodeParams& p) {
  const int lo = min(ra, rb), hi = max(ra, rb);
  if (p.per_class && sm.cls[sm.sidx[lo]] != sm.cls[sm.sidx[hi]]) return;
  atomicOr(&sm.mask[hi * W + (lo >> 5)], 1u << (lo & 31));
}

// The pre-test of a pair (tame images only: every |coordinate| < 1e18 and every area in [1e-30, 1e30]).
// With ta = fl32(thr_lo * area) per box and k = thr_k, the sign of
//     m = fma(inter, k, -(ta_a + ta_b))            (one fp32 add, one fused multiply-add: exact sign)
// settles the common case: m < 0  =>  inter (1 + thr_lo) < thr_lo (a_a + a_b) (1 + 3 * 2^-24)
//                                  =>  inter < thr_lo (1 + 7 * 2^-24) u      (u = fl(fl(a_a + a_b) - inter), inter <= u)
//                                  =>  inter / u < thr  =>  fl32(inter / u) <= thr: the pair survives (:180).
// (thr_lo carries a margin of 2^-18 = 64 * 2^-24; an inverted box has inter = 0 and u > 0 and survives as well.)
// 12 instructions instead of 20 for inter, the union and a compare; whatever the sign does not settle (m >= 0: the
// pair dies or is within 4e-6 of the threshold) is noted and takes the exact test afterwards.
__device__ __forceinline__ float pair_margin(const float4& A, float nta_a, const float4& Bx, float ta_b, float k) {
  const float ww = fmaxf(fminf(Bx.z, A.z) - fmaxf(Bx.x, A.x), 0.f);
  const float hh = fmaxf(fminf(Bx.w, A.w) - fmaxf(Bx.y, A.y), 0.f);
  return fmaf(ww * hh, k, nta_a - ta_b);
}

// General pair code (images with a NaN / infinite / huge coordinate or an area that is not an ordinary positive
// number, thresholds outside the pre-test's range, or more unsettled pairs than the note list holds): every
// unordered pair takes the exact test, the reference's op sequence with its clamp semantics.  Thread <-> row i of the
// score-sorted boxes against the columns (i + d) mod n for d = 1 .. n/2 (for even n the last offset meets each pair
// from both ends, so it runs over i < n/2 only); the sorted boxes are repeated behind their end, so the column walk
// needs no modulo.  When the CTA has room for several threads per row the offsets are split between them.
__device__ __forceinline__ void general_row(const Smem& sm, int n, int W, int i, int d0, int d1, const DecodeParams& p) {
  const float4 A = sm.sbox[i];
  const float area_a = (A.z - A.x) * (A.w - A.y);   // :159 area of box i
  for (int jj = i + d0; jj <= i + d1; ++jj) {
    float inter, u;
    pair_terms<false>(A, area_a, sm.sbox[jj], jj < n, inter, u);
    if (iou_exceeds(inter, u, p)) set_kill(sm, W, i, jj < n ? jj : jj - n, p);
  }
}
__device__ __forceinline__ void general_pairs(const Smem& sm, int n, int W, const DecodeParams& p) {
  const int H = n >> 1;
  if (2 * n > (int)blockDim.x) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int d1 = (!(n & 1) && i >= H) ? H - 1 : H;
      if (d1 >= 1) general_row(sm, n, W, i, 1, d1, p);
    }
    return;
  }
  const int G = (int)blockDim.x / n, K = (H + G - 1) / G;
  for (int t = threadIdx.x; t < n * G; t += blockDim.x) {
    const int g = t / n, i = t - g * n;
    const int d0 = g * K + 1;
    int d1 = min(d0 + K - 1, H);
    if (!(n & 1) && d1 == H && i >= H) --d1;
    if (d0 > d1) continue;
    general_row(sm, n, W, i, d0, d1, p);
  }
}

// ---- sweep (utils/utils.py:162-182): which boxes does the greedy loop keep? --------------------------------
// The matrix holds, for every box j (sorted position), the earlier boxes that kill it if they are kept: column j,
// bit i (i < j).  The greedy result is the one set K with  K(j) = no i < j with K(i) and M[i][j];  it is reached by
// re-evaluating that line for all j at once until nothing changes: an index whose predecessors are settled settles
// with the next pass, so the number of passes is the depth of the kill chains (a handful), not the number of kept
// boxes -- and a pass is a few bit operations per box, with no serial walk over shared memory.
__device__ __forceinline__ unsigned valid_word(int n, int W, int w) {
  return (w == W - 1 && (n & 31)) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
}

// kept-set words -> sm.keep (sorted positions of the kept boxes, ascending) and the count; warp 0
__device__ __forceinline__ void emit_keep(const Smem& sm, int W, int lane, const unsigned* K) {
  int kept = 0;
  for (int w = 0; w < W; ++w) {
    const unsigned a = K[w];
    if ((a >> lane) & 1u) sm.keep[kept + __popc(a & ((1u << lane) - 1u))] = (w << 5) + lane;
    kept += __popc(a);
  }
  if (lane == 0) sm.misc[0] = kept;
}

template <bool SMALL>
__device__ __forceinline__ int sweep_phase(const Smem& sm, int n, int W) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if constexpr (SMALL) {
    // up to 128 boxes: one warp, the columns of lane, lane + 32, ... and the kept set in registers; a pass updates
    // word after word, so later words already see the new earlier ones
    if (warp == 0) {
      unsigned col[4][4], K[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = lane + 32 * c;
#pragma unroll
        for (int w = 0; w < 4; ++w) col[c][w] = (w <= c && c < W && j < n) ? sm.mask[j * W + w] : 0u;
        K[c] = c < W ? valid_word(n, W, c) : 0u;
      }
      bool changed;
      do {
        changed = false;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < W) {
            unsigned hit = 0;
#pragma unroll
            for (int w = 0; w <= c; ++w) hit |= K[w] & col[c][w];
            const unsigned nk = __ballot_sync(0xffffffffu, hit == 0u) & valid_word(n, W, c);
            changed |= nk != K[c];
            K[c] = nk;
          }
        }
      } while (changed);
      int kept = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {   // (K stays in registers: no run-time index)
        if ((K[c] >> lane) & 1u) sm.keep[kept + __popc(K[c] & ((1u << lane) - 1u))] = (c << 5) + lane;
        kept += __popc(K[c]);
      }
      if (lane == 0) sm.misc[0] = kept;
    }
  } else {
    // up to 1024 boxes: every thread re-evaluates its columns against the kept set of the previous pass (two
    // copies of the set in shared memory, one barrier per pass)
    unsigned* cur = reinterpret_cast<unsigned*>(sm.misc) + kKeepA;
    unsigned* nxt = reinterpret_cast<unsigned*>(sm.misc) + kKeepB;
    if ((int)threadIdx.x < W) cur[threadIdx.x] = valid_word(n, W, threadIdx.x);
    __syncthreads();
    while (true) {
      bool changed = false;
      for (int j = threadIdx.x; j < (W << 5); j += blockDim.x) {   // a warp covers one word of columns per trip
        const int wj = j >> 5;
        unsigned hit = 0;
        if (j < n) {
          const unsigned* c = sm.mask + j * W;
          for (int w = 0; w <= wj; ++w) hit |= cur[w] & c[w];
        }
        const unsigned nk = __ballot_sync(0xffffffffu, j < n && hit == 0u);
        if (lane == 0) nxt[wj] = nk;
        changed |= nk != cur[wj];
      }
      const int any = __syncthreads_or(changed);
      unsigned* t = cur;
      cur = nxt, nxt = t;
      if (!any) break;
    }
    if (warp == 0) emit_keep(sm, W, lane, cur);
  }
  __syncthreads();
  return sm.misc[0];
}

// ---- counting sorts -----------------------------------------------------------------------------------------
// kNB buckets with 16-bit counters, two per word (n <= 1024).  add() returns the element's slot inside its bucket.
__device__ __forceinline__ int hist_add(uint32_t* h, int b) {
  const int sh = (b & 1) << 4;
  return (int)((atomicAdd(&h[b >> 1], 1u << sh) >> sh) & 0xffffu);
}
__device__ __forceinline__ int hist_get(const uint32_t* h, int b) { return (int)((h[b >> 1] >> ((b & 1) << 4)) & 0xffffu); }
// counts -> exclusive starts, in place; one warp, 8 buckets (4 words) per lane.  start(kNB) is n (hist_start).
__device__ __forceinline__ void hist_scan(uint32_t* h, int lane) {
  uint32_t w[4];
  int c[8], total = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    w[t] = h[lane * 4 + t];
    c[2 * t] = (int)(w[t] & 0xffffu), c[2 * t + 1] = (int)(w[t] >> 16);
    total += c[2 * t] + c[2 * t + 1];
  }
  int incl = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  int run = incl - total;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int s0 = run, s1 = run + c[2 * t];
    run = s1 + c[2 * t + 1];
    h[lane * 4 + t] = (uint32_t)s0 | ((uint32_t)s1 << 16);
  }
}
__device__ __forceinline__ int hist_start(const uint32_t* h, int b, int n) { return b >= kNB ? n : hist_get(h, b); }

// scores as unsigned keys in their float order (torch.sort(descending=True), :161): NaN above everything, -0 == +0
__device__ __forceinline__ uint32_t score_key(float s) {
  if (s != s) return 0xffffffffu;
  if (s == 0.f) return 0x80000000u;
  const uint32_t b = __float_as_uint(s);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// before the barrier that precedes nms_phase: clear what the phase accumulates into with atomics
// (SMALL: only the area buckets are used -- 128 words, one 16-byte store in each of 32 threads)
template <bool SMALL>
__device__ __forceinline__ void nms_prepare(const Smem& sm) {
  uint4* h4 = reinterpret_cast<uint4*>(SMALL ? sm.hist + kNB / 2 : sm.hist);
  constexpr int kWords4 = SMALL ? kNB / 8 : kNB / 4;   // <= 64: one predicated store, no loop (every CTA has >= 96 threads)
  if ((int)threadIdx.x < kWords4) h4[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) sm.misc[kNotes] = 0, sm.misc[kKmin] = -1, sm.misc[kKmax] = 0, sm.misc[kRankSum] = 0;
}
// the suppression matrix starts empty (16-byte stores; the region behind it is 16-byte padded)
__device__ __forceinline__ void clear_mask(const Smem& sm, int words) {
  // (a 7x7 grid's matrix is one store per thread; as a generic strided loop the two clears were 3 % of the kernel's
  //  instructions.  Making the CTA size a compile-time constant in the small-grid phases, on the other hand, lets the
  //  compiler restructure their loops and costs 13 %: measured, not done.)
  uint4* m4 = reinterpret_cast<uint4*>(sm.mask);
  const int q = (words + 3) >> 2;
  if ((int)threadIdx.x < q) m4[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
  for (int t = threadIdx.x + blockDim.x; t < q; t += blockDim.x) m4[t] = make_uint4(0u, 0u, 0u, 0u);
}

// ---- phase: sort + suppression matrix + sweep (utils/utils.py:150-184).  Returns the kept count. -------------
//  (1) score order.  Candidates are dealt into kNB buckets by the leading bits of their score key (the image's own
//      key range mapped onto the buckets); a prefix over the buckets places each bucket, and inside a bucket -- one
//      or two candidates -- the rank is settled by comparing scores, ties by emission index (canonical, SURVEY B.3;
//      equal scores always share a bucket).  O(n) where rank-by-counting compared all n^2 pairs.
//  (2) area order.  IoU = inter / (a + b - inter) <= min(a, b) / max(a, b) (inter <= min(a, b) holds in fp32 too:
//      min / max / subtract / multiply are monotone), so a pair with max > min / thr survives whatever its position:
//      fl32(inter / u) <= (min / max)(1 + 5 * 2^-24).  Buckets of the area's float bits >> 20 (8 per octave, exact
//      powers of two between octaves) put every box next to the partners that can matter: positions up to the end
//      of bucket b + area_skip - 1 (set_threshold derives area_skip with a 1e-5 margin).  At thr = 0.5 that is a
//      third of all pairs.  Clamping the bucket index only ever shrinks a distance, so it never skips a live pair.
//  (3) the partner lists of all boxes are laid end to end (prefix sum) and cut into equal shares, one per thread:
//      every lane runs the same number of pre-tests whatever its boxes' list lengths.  A pair the pre-test does not
//      settle is noted (shared-memory list) and takes the exact test after the loop, all threads together.
__device__ __forceinline__ int nms_phase_large(const Smem& sm, int n, const DecodeParams& p) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  const int W = (n + 31) >> 5;
  uint32_t* hs = sm.hist;            // score buckets
  uint32_t* ha = sm.hist + kNB / 2;  // area buckets
  clear_mask(sm, n * W);   // dead pairs are OR-ed in by the exact tests
  // ---- the image's score-key range ----
  uint32_t kmin = 0xffffffffu, kmax = 0u;
  for (int k = tid; k < n; k += nt) {
    const uint32_t key = score_key(sm.score[k]);
    kmin = min(kmin, key), kmax = max(kmax, key);
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin), kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (lane == 0) {
    atomicMin(reinterpret_cast<unsigned*>(&sm.misc[kKmin]), kmin);
    atomicMax(reinterpret_cast<unsigned*>(&sm.misc[kKmax]), kmax);
  }
  __syncthreads();
  kmax = (uint32_t)sm.misc[kKmax];
  const uint32_t range = kmax - (uint32_t)sm.misc[kKmin];
  const int shift = range ? max(0, 32 - __clz(range) - 8) : 0;   // (kmax - key) >> shift < kNB
  // ---- slots in the score buckets and in the area buckets; is the image tame? ----
  bool wild = p.thr_lo == 0.f;   // a threshold outside the pre-test's range: general code for every pair
  for (int k = tid; k < n; k += nt) {
    const int db = (int)((kmax - score_key(sm.score[k])) >> shift);   // bucket 0 holds the largest scores
    sm.tkey[k] = ((uint32_t)db << 16) | (uint32_t)hist_add(hs, db);
    const float4 b = sm.box[k];
    const float area = (b.z - b.x) * (b.w - b.y);  // :159
    wild |= !(fabsf(b.x) < 1.0e18f && fabsf(b.y) < 1.0e18f && fabsf(b.z) < 1.0e18f && fabsf(b.w) < 1.0e18f &&
              area >= 1.0e-30f && area <= 1.0e30f);
    const int ab = min(max((int)(__float_as_uint(area) >> 20) - kAreaKey0, 0), kNB - 1);
    sm.akey[k] = ((uint32_t)ab << 16) | (uint32_t)hist_add(ha, ab);
  }
  const bool tame = !__syncthreads_or(wild);
  if (warp == 0) hist_scan(hs, lane);
  if (warp == (nwarps > 1 ? 1 : 0)) hist_scan(ha, lane);
  __syncthreads();
  // ---- into bucket order ----
  for (int k = tid; k < n; k += nt) {
    const uint32_t t = sm.tkey[k];
    sm.bidx[hist_get(hs, (int)(t >> 16)) + (int)(t & 0xffffu)] = k;
    if (tame) {
      const uint32_t a = sm.akey[k];
      const int ab = (int)(a >> 16), apos = hist_get(ha, ab) + (int)(a & 0xffffu);
      const float4 b = sm.box[k];
      sm.abox[apos] = b, sm.ata[apos] = p.thr_lo * ((b.z - b.x) * (b.w - b.y));
      sm.akey[k] = (uint32_t)apos;
      // partners: the later positions up to the end of bucket ab + area_skip - 1
      sm.lpre[apos] = max(hist_start(ha, ab + p.area_skip, n) - 1 - apos, 0);
    }
  }
  __syncthreads();
  // ---- exact rank inside the score bucket; the partner counts become offsets (last warp, meanwhile) ----
  if (tame && warp == nwarps - 1) {
    int carry = 0;
    for (int base = 0; base < n; base += 32) {
      const int v = base + lane < n ? sm.lpre[base + lane] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      if (base + lane < n) sm.lpre[base + lane] = carry + incl - v;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) sm.lpre[n] = carry;
  }
  for (int k = tid; k < n; k += nt) {
    const int db = (int)(sm.tkey[k] >> 16), q0 = hist_get(hs, db), q1 = hist_start(hs, db + 1, n);
    const float s = sm.score[k];
    const bool s_nan = s != s;
    int rank = q0;
    for (int q = q0; q < q1; ++q) {
      const int m = sm.bidx[q];
      const float v = sm.score[m];
      const bool v_nan = v != v;   // NaN scores (stand-alone yolo1_nms only) order before every number
      rank += (v > s) || (v_nan && !s_nan) || ((v == s || (v_nan && s_nan)) && m < k);
    }
    sm.sidx[rank] = k;
    sm.tkey[k] = (uint32_t)rank;
    if (tame) sm.arank[sm.akey[k]] = rank;
  }
  __syncthreads();
  bool general = !tame;
  if (tame) {
    // ---- pre-tests: equal shares of the concatenated partner lists ----
    const int T = sm.lpre[n], share = (T + nt - 1) / nt, cap = 2 * p.max_n;
    int idx = tid * share;
    const int end = min(idx + share, T);
    if (idx < end) {
      int lo = 0, hi = n - 1;   // the row that holds pair idx: the last position whose offset is <= idx
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (sm.lpre[mid] <= idx) lo = mid; else hi = mid - 1;
      }
      int row = lo, row_end = sm.lpre[row + 1], col = row + 1 + (idx - sm.lpre[row]);
      float4 A = sm.abox[row];
      float nta = -sm.ata[row];
      const float k = p.thr_k;
      for (; idx < end; ++idx, ++col) {
        if (idx >= row_end) {   // next row with partners
          do {
            ++row;
            row_end = sm.lpre[row + 1];
          } while (idx >= row_end);
          col = row + 1 + (idx - sm.lpre[row]);
          A = sm.abox[row], nta = -sm.ata[row];
        }
        if (!(pair_margin(A, nta, sm.abox[col], sm.ata[col], k) < 0.f)) {
          const int slot = atomicAdd(&sm.misc[kNotes], 1);
          if (slot < cap) sm.notes[slot] = ((uint32_t)row << 16) | (uint32_t)col;
        }
      }
    }
    __syncthreads();
    // ---- exact tests of the noted pairs ----
    const int notes = sm.misc[kNotes];
    general = notes > cap;   // more than the list holds (e.g. a stack of identical boxes): general code instead
    if (!general) {
      for (int t = tid; t < notes; t += nt) {
        const uint32_t nn = sm.notes[t];
        const int a = (int)(nn >> 16), b = (int)(nn & 0xffffu);
        const float4 A = sm.abox[a];
        float inter, u;
        pair_terms<true>(A, (A.z - A.x) * (A.w - A.y), sm.abox[b], true, inter, u);
        if (iou_exceeds(inter, u, p)) set_kill(sm, W, sm.arank[a], sm.arank[b], p);
      }
    }
    __syncthreads();
  }
  if (general) {   // uniform
    for (int k = tid; k < n; k += nt) {
      const int r = (int)sm.tkey[k];
      const float4 b = sm.box[k];
      sm.sbox[r] = b;
      if (r < (n >> 1)) sm.sbox[n + r] = b;   // wrap copy: row i reads the columns i+1 .. i+n/2 without a modulo
    }
    __syncthreads();
    general_pairs(sm, n, W, p);
    __syncthreads();
  }
  return sweep_phase<false>(sm, n, W);
}

// Grids of up to 128 candidates (96-thread CTAs, 16 to an SM).  Measured (B200, S = 7): the bucketed score order and
// the equal-share pair loop above cut the instruction count but not the time -- their phases are chains of dependent
// shared-memory operations (atomics that return slots, a binary search, bucket walks) behind eight CTA barriers, and a
// 3-warp CTA cannot hide them (90 M images / s at 4096 images against 110 M before).  So here:
//   * the score rank stays a rank by counting (four scores per 128-bit load, independent fp32 counters: pure
//     issue throughput, no dependent chain; a checksum detects equal scores, which are ranked again with the tie rule),
//   * the area buckets ride along: a candidate's bucket slot is requested before its rank loop and read after it,
//   * thread <-> box in area order, against its partners i+1 .. i+L (the later positions up to the end of bucket
//     b + area_skip - 1); the loop only collects the sign bits of the pre-test (32 columns per round) and the lanes
//     work their unsettled columns off together after the round.  A warp runs as long as its longest partner list.
__device__ __forceinline__ int nms_phase_small(const Smem& sm, int n, const DecodeParams& p) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int W = (n + 31) >> 5;
  uint32_t* ha = sm.hist + kNB / 2;
  clear_mask(sm, n * W);
  bool wild = p.thr_lo == 0.f;
  int ranks = 0;
  for (int k = tid; k < n; k += nt) {
    const float s = sm.score[k];
    const float4 b = sm.box[k];
    const float area = (b.z - b.x) * (b.w - b.y);  // :159
    wild |= !(fabsf(b.x) < 1.0e18f && fabsf(b.y) < 1.0e18f && fabsf(b.z) < 1.0e18f && fabsf(b.w) < 1.0e18f &&
              area >= 1.0e-30f && area <= 1.0e30f);
    const int ab = min(max((int)(__float_as_uint(area) >> 20) - kAreaKey0, 0), kNB - 1);
    const int aslot = hist_add(ha, ab);
    float r0 = 0.f, r1 = 0.f, r2 = 0.f, r3 = 0.f;   // counts < 2^24 are exact
    const int n4 = n & ~3;
#pragma unroll kRankUnroll
    for (int m = 0; m < n4; m += 4) {
      const float4 v = *reinterpret_cast<const float4*>(sm.score + m);
      r0 += v.x > s ? 1.f : 0.f, r1 += v.y > s ? 1.f : 0.f, r2 += v.z > s ? 1.f : 0.f, r3 += v.w > s ? 1.f : 0.f;
    }
    for (int m = n4; m < n; ++m) r0 += sm.score[m] > s ? 1.f : 0.f;
    const int rank = (int)((r0 + r1) + (r2 + r3));
    ranks += rank;
    sm.sidx[rank] = k;
    sm.tkey[k] = (uint32_t)rank;
    sm.akey[k] = ((uint32_t)ab << 16) | (uint32_t)aslot;
  }
  ranks = __reduce_add_sync(0xffffffffu, ranks);
  if (lane == 0 && ranks) atomicAdd(&sm.misc[kRankSum], ranks);
  const bool tame = !__syncthreads_or(wild);
  if (sm.misc[kRankSum] != n * (n - 1) / 2) {   // uniform: equal (or NaN) scores somewhere -- again, with the tie rule
    for (int k = tid; k < n; k += nt) {
      const float s = sm.score[k];
      const bool s_nan = s != s;   // NaN scores (stand-alone yolo1_nms only) order before every number (:161)
      int rank = 0;
      for (int m = 0; m < n; ++m) {
        const float v = sm.score[m];
        const bool v_nan = v != v;
        rank += (v > s) || (v_nan && !s_nan) || ((v == s || (v_nan && s_nan)) && m < k);
      }
      sm.sidx[rank] = k;
      sm.tkey[k] = (uint32_t)rank;
    }
  }
  if (tame) {
    if (warp == 0) hist_scan(ha, lane);
    __syncthreads();
    for (int k = tid; k < n; k += nt) {
      const uint32_t a = sm.akey[k];
      const int ab = (int)(a >> 16), apos = hist_get(ha, ab) + (int)(a & 0xffffu);
      const float4 b = sm.box[k];
      sm.abox[apos] = b, sm.ata[apos] = p.thr_lo * ((b.z - b.x) * (b.w - b.y));
      sm.arank[apos] = (int)sm.tkey[k];
      sm.lpre[apos] = max(hist_start(ha, ab + p.area_skip, n) - 1 - apos, 0);
    }
    __syncthreads();
    const float k = p.thr_k;
    for (int i = tid; i < n; i += nt) {
      const int jend = i + sm.lpre[i];
      if (jend == i) continue;
      const float4 A = sm.abox[i];
      const float nta = -sm.ata[i];
      for (int base = i + 1; base <= jend; base += 32) {
        const int cend = min(base + 31, jend);
        unsigned sure = 0;   // bit (cend - c) set: column c survives
        int jj = base;
#pragma unroll kPairUnroll
        for (; jj < cend; jj += 2) {   // two columns per trip
          const float m0 = pair_margin(A, nta, sm.abox[jj], sm.ata[jj], k);
          const float m1 = pair_margin(A, nta, sm.abox[jj + 1], sm.ata[jj + 1], k);
          sure = __funnelshift_l(__float_as_uint(m0), sure, 1);
          sure = __funnelshift_l(__float_as_uint(m1), sure, 1);
        }
        if (jj == cend) sure = __funnelshift_l(__float_as_uint(pair_margin(A, nta, sm.abox[jj], sm.ata[jj], k)), sure, 1);
        unsigned todo = ~sure & (0xffffffffu >> (31 - (cend - base)));
        while (todo) {   // the exact test of a column the pre-test left open (rare)
          const int c = cend - (__ffs(todo) - 1);
          todo &= todo - 1;
          float inter, u;
          pair_terms<true>(A, (A.z - A.x) * (A.w - A.y), sm.abox[c], true, inter, u);
          if (iou_exceeds(inter, u, p)) set_kill(sm, W, sm.arank[i], sm.arank[c], p);
        }
      }
    }
    __syncthreads();
  } else {
    __syncthreads();   // the ranks of the tie pass
    for (int k = tid; k < n; k += nt) {
      const int r = (int)sm.tkey[k];
      const float4 b = sm.box[k];
      sm.sbox[r] = b;
      if (r < (n >> 1)) sm.sbox[n + r] = b;
    }
    __syncthreads();
    general_pairs(sm, n, W, p);
    __syncthreads();
  }
  return sweep_phase<true>(sm, n, W);
}

template <bool SMALL>
__device__ __forceinline__ int nms_phase(const Smem& sm, int n, const DecodeParams& p) {
  if constexpr (SMALL)
    return nms_phase_small(sm, n, p);
  else
    return nms_phase_large(sm, n, p);
}

// ---- kernels ---------------------------------------------------------------------------------------------
template <typename E, bool SMALL>
__device__ __forceinline__ void decode_nms_image(const DecodeParams& p) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem sm;
  smem_layout(raw, p.S * p.S * (5 * p.B + p.C), p.max_n, &sm);
  const int64_t n = blockIdx.x;
  // Image load.  Dense fp32 tensors: ONE bulk (TMA) copy issued by thread 0 (UBLKCP; completion on an mbarrier)
  // replaces eight 8-byte load / store pairs per thread -- 8 % of the kernel's instructions, and the bytes land
  // without passing through registers.  Bulk copies move multiples of 16 bytes between 16-byte aligned addresses,
  // while an image of S*S*D floats may start 8 bytes off (odd images of a 7x7x30 tensor): the copy then starts 8
  // bytes early (inside the previous image) and `img` points 8 bytes into the buffer; an image whose rounded-up copy
  // would end beyond the tensor (the last one) takes the per-thread loads.
  bool bulk = false;
  if (sizeof(E) == 4 && p.bulk_ok) {
    const uint32_t img_bytes = (uint32_t)(p.S * p.S * (5 * p.B + p.C)) * 4u;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(p.pred) + n * (int64_t)img_bytes;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint32_t bytes = (img_bytes + mis + 15u) & ~15u;
    bulk = !(n == p.n_images - 1 && bytes > img_bytes + mis);
    if (bulk) {
      uint64_t* bar = reinterpret_cast<uint64_t*>(sm.misc + kBar);
      if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(raw, src - mis, bytes, bar, policy_evict_first());
#if YOLO1_DECODE_L2PF
        // the image of the CTA that will take this CTA's place (one wave = 148 SMs x 16 CTAs later): into L2 now
        if (n + YOLO1_DECODE_L2PF < p.n_images - 1)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src - mis + (int64_t)YOLO1_DECODE_L2PF * img_bytes),
                       "r"(bytes - 16u)
                       : "memory");
#endif
      }
      sm.img = reinterpret_cast<float*>(raw + mis);
      nms_prepare<SMALL>(sm);
      __syncthreads();          // the barrier is initialised before anyone polls it
      mbar_wait(bar, 0);
    }
  }
  if (!bulk) {
    load_image<E>(p, n, sm.img);
    nms_prepare<SMALL>(sm);
    __syncthreads();
  }
  const int cand = decode_phase<SMALL>(p, sm);
  const int kept = cand > 0 ? nms_phase<SMALL>(sm, cand, p) : 0;
  // kept detections in descending score order, rows beyond the count zero; streaming stores (written once, read by
  // another kernel or the host: no reason to keep the lines in L2 ahead of the next kernel's traffic)
  for (int t = threadIdx.x; t < p.max_n; t += blockDim.x) {
    const int64_t dst = n * p.max_n + t;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    float sc = 0.f;
    int cl = 0, e = 0;
    if (t < kept) {
      e = sm.sidx[sm.keep[t]];
      bx = sm.box[e], sc = sm.score[e], cl = sm.cls[e];
    }
    __stcs(reinterpret_cast<float4*>(p.out_boxes) + dst, bx);
    __stcs(p.out_scores + dst, sc);
    __stcs(p.out_cls + dst, cl);
    if (p.keep) __stcs(p.keep + dst, e);
  }
  if (threadIdx.x == 0) {
    p.out_counts[n] = kept;
    if (p.cand_counts) p.cand_counts[n] = cand;
  }
}

#if YOLO1_DECODE_PERSIST
// ---- measured and NOT shipped (round 2; -DYOLO1_DECODE_PERSIST=1 builds it, tools/build_decode_variant.sh) ----------
// VERDICT r1 item 5 asked for a persistent grid that prefetches the next image.  On B200 (tools/compare_decode_libs.py,
// one box, M images/s at S=7 N=4096 / N=65536, profiles/decode_persist_r2.log): one CTA per image 133 / 190;
// persistent with the split layout (14 CTAs per SM) 105 / 147; persistent over the shared layout, next image issued
// behind the sweep (16 CTAs per SM) 105 / 150.  The image loop keeps the loop state and the layout's addresses alive
// across all phases: at the 40 registers that 16 (or 14) resident CTAs allow that is 150 bytes of spills inside a
// kernel that is bound by instruction issue, and it costs more than the hidden load latency returns.  An L2 prefetch
// of the image one wave ahead (YOLO1_DECODE_L2PF=2368, one instruction per CTA) also loses: 125 / 189.
// (Also measured: the 7x7x30 shape as compile-time constants -- 16 % fewer static instructions, 196 at N=65536 but
//  125-130 for the lone 4096-image launch, profiles/decode_persist_r2.log; not kept.)
// Persistent small-grid form (dense fp32 input, up to 128 candidates): the grid is what fits the machine at once and
// CTA b takes images b, b + grid, b + 2 grid, ...  The image region is not shared with the NMS arrays (split layout),
// so the bulk load of the CTA's NEXT image is issued the moment the decode phase has read the current one and lands
// while rank, pair loop, sweep and the output stores run: the load latency a fresh CTA would sit through at its
// start (and the block scheduler's relaunch) is paid once per CTA instead of once per image.  Cost: 1 968 bytes more
// shared memory per CTA, 14 resident CTAs per SM instead of 16.
// image n as a bulk copy: the source rounded out to 16-byte bounds (see decode_nms_image); false for the one image
// whose rounded copy would end beyond the tensor
__device__ __forceinline__ bool bulk_shape(const DecodeParams& p, uint32_t img_bytes, int n, uint32_t& mis,
                                           uint32_t& bytes) {
  const uintptr_t src = reinterpret_cast<uintptr_t>(p.pred) + (uint64_t)n * img_bytes;
  mis = (uint32_t)(src & 15u);
  bytes = (img_bytes + mis + 15u) & ~15u;
  return !(n == (int)p.n_images - 1 && bytes > img_bytes + mis);
}
__device__ __forceinline__ void bulk_issue(const DecodeParams& p, uint32_t img_bytes, int n, unsigned char* raw,
                                           uint64_t* bar) {   // one thread
  uint32_t mis, bytes;
  if (!bulk_shape(p, img_bytes, n, mis, bytes)) return;
  mbar_arrive_expect_tx(bar, bytes);
  bulk_g2s(raw, reinterpret_cast<const unsigned char*>(p.pred) + (uint64_t)n * img_bytes - mis, bytes, bar,
           policy_evict_first());
}

template <bool SPLIT>
__global__ void __launch_bounds__(96, SPLIT ? 14 : 16) decode_nms_persistent_kernel(const __grid_constant__ DecodeParams p) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem sm;
  smem_layout(raw, p.S * p.S * (5 * p.B + p.C), p.max_n, &sm, SPLIT);
  const uint32_t img_bytes = (uint32_t)(p.S * p.S * (5 * p.B + p.C)) * 4u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm.misc + kBar);
  const int n_images = (int)p.n_images;   // one CTA per image in the other form: the count fits an int
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    bulk_issue(p, img_bytes, (int)blockIdx.x, raw, bar);
  }
  uint32_t phase = 0;
  for (int n = blockIdx.x; n < n_images; n += gridDim.x) {
    nms_prepare<true>(sm);
    __syncthreads();            // first trip: the barrier is initialised before anyone polls it
    uint32_t mis, bytes;
    if (bulk_shape(p, img_bytes, n, mis, bytes)) {
      sm.img = reinterpret_cast<float*>(raw + mis);
      mbar_wait(bar, phase);
      phase ^= 1u;
    } else {
      sm.img = reinterpret_cast<float*>(raw);
      load_image<float>(p, n, sm.img);
      __syncthreads();
    }
    const int cand = decode_phase<true>(p, sm);   // ends behind a CTA barrier: nobody reads the image any more
    if (SPLIT && threadIdx.x == 0 && n + (int)gridDim.x < n_images) bulk_issue(p, img_bytes, n + gridDim.x, raw, bar);
    const int kept = cand > 0 ? nms_phase<true>(sm, cand, p) : 0;
    // !SPLIT: the image region holds the suppression matrix until the sweep is over; the next image can only travel
    // behind the output stores (generic-proxy writes to the region happened before: order them before the copy)
    if (!SPLIT && threadIdx.x == 0 && n + (int)gridDim.x < n_images) {
      fence_async_smem();
      bulk_issue(p, img_bytes, n + gridDim.x, raw, bar);
    }
    for (int t = threadIdx.x; t < p.max_n; t += blockDim.x) {
      const int64_t dst = (int64_t)n * p.max_n + t;
      float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
      float sc = 0.f;
      int cl = 0, e = 0;
      if (t < kept) {
        e = sm.sidx[sm.keep[t]];
        bx = sm.box[e], sc = sm.score[e], cl = sm.cls[e];
      }
      __stcs(reinterpret_cast<float4*>(p.out_boxes) + dst, bx);
      __stcs(p.out_scores + dst, sc);
      __stcs(p.out_cls + dst, cl);
      if (p.keep) __stcs(p.keep + dst, e);
    }
    if (threadIdx.x == 0) {
      p.out_counts[n] = kept;
      if (p.cand_counts) p.cand_counts[n] = cand;
    }
    __syncthreads();            // the candidate arrays are free for the next image
  }
}
#endif  // YOLO1_DECODE_PERSIST

template <typename E, bool SMALL>
__global__ void decode_nms_kernel(const __grid_constant__ DecodeParams p) {
  decode_nms_image<E, SMALL>(p);
}
// grids of up to 128 candidates run 96-thread CTAs, 16 to an SM: 40 registers (the pair loop would take 48 unbounded)
template <>
__global__ void __launch_bounds__(96, 16) decode_nms_kernel<float, true>(const __grid_constant__ DecodeParams p) {
  decode_nms_image<float, true>(p);
}
template <>
__global__ void __launch_bounds__(96, 16)
    decode_nms_kernel<__nv_bfloat16, true>(const __grid_constant__ DecodeParams p) {
  decode_nms_image<__nv_bfloat16, true>(p);
}

template <typename E>
__global__ void decode_kernel(const __grid_constant__ DecodeParams p) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem sm;
  smem_layout(raw, p.S * p.S * (5 * p.B + p.C), p.max_n, &sm);
  const int64_t n = blockIdx.x;
  load_image<E>(p, n, sm.img);
  __syncthreads();
  const int cand = decode_phase<false>(p, sm);
  for (int t = threadIdx.x; t < cand; t += blockDim.x) {
    const int64_t dst = n * p.max_n + t;
    reinterpret_cast<float4*>(p.boxes)[dst] = sm.box[t];
    p.scores[dst] = sm.score[t];
    p.cls[dst] = sm.cls[t];
  }
  for (int t = cand + threadIdx.x; t < p.max_n; t += blockDim.x) {
    const int64_t dst = n * p.max_n + t;
    reinterpret_cast<float4*>(p.boxes)[dst] = make_float4(0.f, 0.f, 0.f, 0.f);
    p.scores[dst] = 0.f;
    p.cls[dst] = 0;
  }
  if (threadIdx.x == 0) p.counts[n] = cand;
}

template <bool SMALL>
__device__ __forceinline__ void nms_image(const DecodeParams& p) {
  extern __shared__ __align__(16) unsigned char raw[];
  Smem sm;
  smem_layout(raw, 0, p.max_n, &sm);
  const int64_t n = blockIdx.x;
  int cnt = p.counts[n];
  cnt = cnt < 0 ? 0 : (cnt > p.max_n ? p.max_n : cnt);
  for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
    const int64_t src = n * p.max_n + t;
    const float* b = p.boxes + 4 * src;  // caller memory: only 4-byte alignment is required
    sm.box[t] = make_float4(b[0], b[1], b[2], b[3]);
    sm.score[t] = p.scores[src];
    sm.cls[t] = p.cls ? p.cls[src] : 0;
  }
  nms_prepare<SMALL>(sm);
  __syncthreads();
  const int kept = cnt > 0 ? nms_phase<SMALL>(sm, cnt, p) : 0;
  for (int t = threadIdx.x; t < p.max_n; t += blockDim.x)
    p.keep[n * p.max_n + t] = t < kept ? sm.sidx[sm.keep[t]] : 0;
  if (threadIdx.x == 0) p.out_counts[n] = kept;
}

template <bool SMALL>
__global__ void nms_kernel(const __grid_constant__ DecodeParams p) {
  nms_image<SMALL>(p);
}
template <>
__global__ void __launch_bounds__(96, 16) nms_kernel<true>(const __grid_constant__ DecodeParams p) {
  nms_image<true>(p);
}

// run_test_mAP post-processing (utils/utils.py:406-407, :347-354): clamp to [0,1], scale to pixels in fp32,
// truncate toward zero.  One thread per box (float4 in, int4 out).
__global__ void __launch_bounds__(256) boxes_to_pixels_kernel(const float4* __restrict__ boxes, int64_t n,
                                                              float w, float h, int4* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float4 b = boxes[i];
    b.x = fminf(fmaxf(b.x, 0.f), 1.f), b.y = fminf(fmaxf(b.y, 0.f), 1.f);
    b.z = fminf(fmaxf(b.z, 0.f), 1.f), b.w = fminf(fmaxf(b.w, 0.f), 1.f);
    out[i] = make_int4((int)(b.x * w), (int)(b.y * h), (int)(b.z * w), (int)(b.w * h));
  }
}

// measured (profiles/tune_decode_r1.log): 96 threads for a 7x7 grid (98 slots: one row per thread in the pair loop,
// three full warps; 128: -4 %, 160: -13 %, 256: -27 %), 384 for 14x14 (256: -3 %, 512: -3 %)
int block_threads(int max_n) { return max_n <= 128 ? 96 : (max_n <= 256 ? 256 : 384); }

// KERN is a template VALUE parameter, so every kernel gets its own static KernelPrep (the kernels share one type)
template <auto KERN>
int launch(const DecodeParams& p, int64_t N, int img_floats, cudaStream_t stream) {
  if (N == 0) return 0;
  const size_t smem = smem_layout(nullptr, img_floats, p.max_n, nullptr);
  if (smem > 227 * 1024) return YOLO1_ERR_UNSUPPORTED;
  static KernelPrep prep;
  if (int rc = prepare_kernel(prep, KERN, block_threads(p.max_n), smem, false, nullptr, nullptr)) return rc;
  KERN<<<(unsigned)N, block_threads(p.max_n), smem, stream>>>(p);
  return (int)cudaGetLastError();
}

#if YOLO1_DECODE_PERSIST
int launch_persistent(const DecodeParams& p, int64_t N, int img_floats, cudaStream_t stream) {
  constexpr bool kSplit = YOLO1_DECODE_PERSIST_SPLIT;
  const size_t smem = smem_layout(nullptr, img_floats, p.max_n, nullptr, kSplit);
  static KernelPrep prep;
  int sms = kNumSMs, per_sm = 1;
  if (int rc = prepare_kernel(prep, decode_nms_persistent_kernel<kSplit>, 96, smem, true, &sms, &per_sm)) return rc;
  int64_t grid = (int64_t)sms * per_sm;
  if (grid > N) grid = N;
  decode_nms_persistent_kernel<kSplit><<<(unsigned)grid, 96, smem, stream>>>(p);
  return (int)cudaGetLastError();
}
// YOLO1_DECODE_PERSIST = 1: every eligible call; = k > 1: calls of up to k * 2368 images
inline bool persist_wanted(int64_t N) {
  return YOLO1_DECODE_PERSIST == 1 || N <= (int64_t)YOLO1_DECODE_PERSIST * 2368;
}
#endif

int check_decode_args(const void* pred, const int64_t st[4], int dtype, int64_t N, int S, int B, int C) {
  if (!st || N < 0 || S <= 0 || B <= 0 || C <= 0) return YOLO1_ERR_ARG;
  if (N > 0 && !pred) return YOLO1_ERR_ARG;
  if (dtype != YOLO1_DTYPE_F32 && dtype != YOLO1_DTYPE_BF16) return YOLO1_ERR_ARG;
  if (N > 0x7fffffffll) return YOLO1_ERR_UNSUPPORTED;  // one CTA per image
  if ((int64_t)S * S * B > kMaxCand || 5 * B + C > 128) return YOLO1_ERR_UNSUPPORTED;
  if ((uintptr_t)pred % (dtype == YOLO1_DTYPE_F32 ? 4 : 2)) return YOLO1_ERR_ALIGN;
  return 0;
}

// Precompute the division-free threshold test (iou_exceeds).  Proven for finite 0 <= thr < 2^100.
void set_threshold(DecodeParams& p, float thr) {
  p.iou_thr = thr;
  p.thr_fast = (thr >= 0.f && thr < 1.0e30f) ? 1 : 0;
  p.thr_mid = 0.0, p.thr_tie_ok = 0, p.thr_lo = 0.f, p.thr_k = 1.f, p.area_skip = kNB + 1;
  if (p.thr_fast) {
    const float up = nextafterf(thr, INFINITY);
    p.thr_mid = 0.5 * ((double)thr + (double)up);
    uint32_t bits;
    memcpy(&bits, &thr, 4);
    p.thr_tie_ok = (bits & 1u) == 0u;   // ties round to the even mantissa
    // `x <= mid u` as `x < mid' u` with mid' two doubles above mid (see iou_exceeds)
    if (p.thr_tie_ok) p.thr_mid = nextafter(nextafter(p.thr_mid, INFINITY), INFINITY);
    if (thr >= 0x1p-20f && thr <= 4.f) {   // pair_margin: range in which its products stay ordinary numbers
      p.thr_lo = nextafterf((float)((double)thr * (1.0 - 0x1p-18)), 0.f);
      p.thr_k = nextafterf((float)(1.0 + (double)p.thr_lo), INFINITY);
      // Area pruning (nms_phase): area keys are float bits >> 20, i.e. 8 steps per octave with boundaries at
      // 2^e (1 + f / 8).  Two boxes whose keys differ by at least d have areas a < lower(k + 1) and
      // b >= lower(k + d): b / a > G(d) = min over f of lower(f + d - 1) / lower(f).  The pair cannot die once
      // G(d) >= (1 + 1e-5) / thr (IoU <= a / b, five fp32 roundings); area_skip is the smallest such d.
      const double need = (1.0 + 1e-5) / (double)thr;
      p.area_skip = kNB + 1;
      for (int d = 1; d <= kNB; ++d) {
        double g = 1e300;
        for (int f = 0; f < 8; ++f) {
          const int t = f + d - 1;
          const double r = ldexp(1.0 + (t % 8) / 8.0, t / 8) / (1.0 + f / 8.0);
          if (r < g) g = r;
        }
        if (g >= need) {
          p.area_skip = d;
          break;
        }
      }
    }
  }
}

void fill_decode(DecodeParams& p, const void* pred, const int64_t st[4], int S, int B, int C, double thresh) {
  p = DecodeParams{};
  p.pred = pred;
  for (int d = 0; d < 4; ++d) p.st[d] = st[d];
  p.S = S, p.B = B, p.C = C;
  p.cs = (float)(1.0 / (double)S);
  p.thresh = thresh;
  p.max_n = S * S * B;
  p.s_magic = (65536 + S - 1) / S;
  for (int cell = 0; cell < S * S; ++cell)
    if ((int)(((unsigned)cell * (unsigned)p.s_magic) >> 16) != cell / S) p.s_magic = 0;
}

}  // namespace
}  // namespace yolo1

extern "C" {

int yolo1_decode(const void* pred, const int64_t pred_strides[4], int pred_dtype, int64_t N, int S, int B, int C,
                 double thresh, float* boxes, float* scores, int32_t* cls, int32_t* counts, void* stream) {
  using namespace yolo1;
  int rc = check_decode_args(pred, pred_strides, pred_dtype, N, S, B, C);
  if (rc) return rc;
  if (N == 0) return 0;
  if (!boxes || !scores || !cls || !counts) return YOLO1_ERR_ARG;
  if ((uintptr_t)boxes % 16 || (uintptr_t)scores % 4 || (uintptr_t)cls % 4 || (uintptr_t)counts % 4)
    return YOLO1_ERR_ALIGN;
  DecodeParams p;
  fill_decode(p, pred, pred_strides, S, B, C, thresh);
  p.boxes = boxes, p.scores = scores, p.cls = cls, p.counts = counts;
  const int img = S * S * (5 * B + C);
  if (pred_dtype == YOLO1_DTYPE_BF16) return launch<decode_kernel<__nv_bfloat16>>(p, N, img, (cudaStream_t)stream);
  return launch<decode_kernel<float>>(p, N, img, (cudaStream_t)stream);
}

int yolo1_nms(const float* boxes, const float* scores, const int32_t* cls, const int32_t* counts, int64_t N,
              int max_n, float iou_thr, int per_class, int32_t* keep, int32_t* keep_counts, void* stream) {
  using namespace yolo1;
  if (N < 0 || max_n <= 0) return YOLO1_ERR_ARG;
  if (N == 0) return 0;
  if (!boxes || !scores || !counts || !keep || !keep_counts) return YOLO1_ERR_ARG;
  if (per_class && !cls) return YOLO1_ERR_ARG;
  if (max_n > kMaxCand || N > 0x7fffffffll) return YOLO1_ERR_UNSUPPORTED;
  if ((uintptr_t)boxes % 4 || (uintptr_t)scores % 4 || (cls && (uintptr_t)cls % 4) || (uintptr_t)counts % 4 ||
      (uintptr_t)keep % 4 || (uintptr_t)keep_counts % 4)
    return YOLO1_ERR_ALIGN;
  DecodeParams p = DecodeParams{};
  p.boxes = const_cast<float*>(boxes), p.scores = const_cast<float*>(scores), p.cls = const_cast<int32_t*>(cls);
  p.counts = const_cast<int32_t*>(counts);
  p.keep = keep, p.out_counts = keep_counts;
  p.max_n = max_n, p.per_class = per_class ? 1 : 0;
  set_threshold(p, iou_thr);
  if (max_n <= 128) return launch<nms_kernel<true>>(p, N, 0, (cudaStream_t)stream);
  return launch<nms_kernel<false>>(p, N, 0, (cudaStream_t)stream);
}

int yolo1_boxes_to_pixels(const float* boxes, int64_t n_boxes, float img_w, float img_h, int32_t* pixels,
                          void* stream) {
  if (n_boxes < 0) return YOLO1_ERR_ARG;
  if (n_boxes == 0) return 0;
  if (!boxes || !pixels) return YOLO1_ERR_ARG;
  if ((uintptr_t)boxes % 16 || (uintptr_t)pixels % 16) return YOLO1_ERR_ALIGN;
  int64_t grid = (n_boxes + 255) / 256;
  if (grid > yolo1::kNumSMs * 16) grid = yolo1::kNumSMs * 16;
  yolo1::boxes_to_pixels_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(boxes), n_boxes, img_w, img_h, reinterpret_cast<int4*>(pixels));
  return (int)cudaGetLastError();
}

int yolo1_decode_nms(const void* pred, const int64_t pred_strides[4], int pred_dtype, int64_t N, int S, int B,
                     int C, double thresh, float iou_thr, int per_class, float* out_boxes, float* out_scores,
                     int32_t* out_cls, int32_t* out_counts, int32_t* keep_idx, int32_t* cand_counts,
                     void* stream) {
  using namespace yolo1;
  int rc = check_decode_args(pred, pred_strides, pred_dtype, N, S, B, C);
  if (rc) return rc;
  if (N == 0) return 0;
  if (!out_boxes || !out_scores || !out_cls || !out_counts) return YOLO1_ERR_ARG;
  if ((uintptr_t)out_boxes % 16 || (uintptr_t)out_scores % 4 || (uintptr_t)out_cls % 4 ||
      (uintptr_t)out_counts % 4 || (keep_idx && (uintptr_t)keep_idx % 4) ||
      (cand_counts && (uintptr_t)cand_counts % 4))
    return YOLO1_ERR_ALIGN;
  DecodeParams p;
  fill_decode(p, pred, pred_strides, S, B, C, thresh);
  p.per_class = per_class ? 1 : 0;
  set_threshold(p, iou_thr);
  p.out_boxes = out_boxes, p.out_scores = out_scores, p.out_cls = out_cls, p.out_counts = out_counts;
  p.keep = keep_idx, p.cand_counts = cand_counts;
  const int Dch = 5 * B + C;
  p.n_images = N;
  p.bulk_ok = pred_dtype == YOLO1_DTYPE_F32 && pred_strides[3] == 1 && pred_strides[2] == Dch &&
              pred_strides[1] == (int64_t)S * Dch && pred_strides[0] == (int64_t)S * S * Dch &&
              (uintptr_t)pred % 16 == 0 && ((int64_t)S * S * Dch * 4) % 8 == 0;
  const int img = S * S * (5 * B + C);
  const bool defer = p.max_n <= 128;   // up to 128 candidates: 96-thread CTAs and the one-warp register sweep
#if YOLO1_DECODE_PERSIST
  if (defer && p.bulk_ok && persist_wanted(N)) return launch_persistent(p, N, img, (cudaStream_t)stream);
#endif
  if (pred_dtype == YOLO1_DTYPE_BF16)
    return defer ? launch<decode_nms_kernel<__nv_bfloat16, true>>(p, N, img, (cudaStream_t)stream)
                 : launch<decode_nms_kernel<__nv_bfloat16, false>>(p, N, img, (cudaStream_t)stream);
  return defer ? launch<decode_nms_kernel<float, true>>(p, N, img, (cudaStream_t)stream)
               : launch<decode_nms_kernel<float, false>>(p, N, img, (cudaStream_t)stream);
}

}  // extern "C"
