/*
 * yolo1_oracle.c -- CPU restatement of the haoran1062/YOLO_V1 hot path.
 *
 * TEST INFRASTRUCTURE.  This file is the parity oracle: a plain scalar C restatement of the
 * reference's algorithm.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load it, and only as the checker or as the timed CPU baseline.
 * The product (yolo_v1_b200/) never links, imports or falls back to anything in oracle/.
 *
 * Parity status: PINNED.  oracle/make_golden.py runs the reference's own Python code
 * (/root/reference, torch CPU) on seeded synthetic inputs in the build container and commits the
 * input/output vectors to tests/golden/; tests/test_oracle_golden.py checks this file against them
 * (loss/grad <= 2e-6 relative, decode/NMS bit-exact) plus the reference's two print-only fixtures
 * (utils/utils.py:321-324 voc_eval, utils/utils.py:506-525 IoU matrix).
 *
 * All citations are relative to /root/reference.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: decode/NMS must round every fp32 operation exactly like the ATen CPU
 * kernels do (no fused multiply-add), otherwise borderline `ovr <= thr` decisions flip.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_B 8

/* ------------------------------------------------------------------------------------------ */
/* box helpers: utils/utils.py:59-75 (convert_CxCyWH_to_X1Y1X2Y2) and :10-57 (compute_iou_matrix) */
/* ------------------------------------------------------------------------------------------ */

/* utils/utils.py:72-73 : (x,y,w,h) -> (x/S - w/2, y/S - h/2, x/S + w/2, y/S + h/2); the cell offset is
 * not added (it cancels: pred and GT live in the same cell).  0.5*w is exact in fp32. */
static void cxcywh_to_xyxy(const float b[4], float S, float o[4]) {
  float cx = b[0] / S, cy = b[1] / S;
  float hw = 0.5f * b[2], hh = 0.5f * b[3];
  o[0] = cx - hw; o[1] = cy - hh; o[2] = cx + hw; o[3] = cy + hh;
}

/* utils/utils.py:34-55 for one (pred, gt) pair.  No +1, no epsilon; negative extents clipped to 0. */
static float iou_xyxy(const float p[4], const float g[4]) {
  float lx = p[0] > g[0] ? p[0] : g[0];
  float ly = p[1] > g[1] ? p[1] : g[1];
  float rx = p[2] < g[2] ? p[2] : g[2];
  float ry = p[3] < g[3] ? p[3] : g[3];
  float iw = rx - lx, ih = ry - ly;
  if (iw < 0) iw = 0;
  if (ih < 0) ih = 0;
  float inter = iw * ih;
  float ap = (p[2] - p[0]) * (p[3] - p[1]);
  float ag = (g[2] - g[0]) * (g[3] - g[1]);
  return inter / (ap + ag - inter);
}

/* exported for the helper-parity tests (utils/utils.py:10-57): out[n*M+m] */
void yolo1_oracle_iou_matrix(const float* b1, int N, const float* b2, int M, float* out) {
  for (int n = 0; n < N; ++n)
    for (int m = 0; m < M; ++m) out[n * M + m] = iou_xyxy(b1 + 4 * n, b2 + 4 * m);
}

void yolo1_oracle_cxcywh_to_xyxy(const float* in, int n, int S, float* out) {
  for (int i = 0; i < n; ++i) cxcywh_to_xyxy(in + 4 * i, (float)S, out + 4 * i);
}

/* d IoU / d (x,y,w,h) of the pred box, the sub-gradient autograd takes through
 * utils/utils.py:38-55 and :72-73 (min/max ties split 0.5/0.5 as torch >= 1.7 does;
 * `w_h[w_h < 0] = 0` kills the gradient of a clipped extent).  SURVEY.md App. A.4. */
static void iou_grad(const float pb[4], const float gb[4], float S, double dg[4]) {
  float p[4], g[4];
  cxcywh_to_xyxy(pb, S, p);
  cxcywh_to_xyxy(gb, S, g);
  double lx = p[0] > g[0] ? p[0] : g[0], ly = p[1] > g[1] ? p[1] : g[1];
  double rx = p[2] < g[2] ? p[2] : g[2], ry = p[3] < g[3] ? p[3] : g[3];
  double iw = rx - lx, ih = ry - ly;
  dg[0] = dg[1] = dg[2] = dg[3] = 0.0;
  if (iw < 0 || ih < 0) return;
  double pw = (double)p[2] - p[0], ph = (double)p[3] - p[1];
  double ap = pw * ph, ag = ((double)g[2] - g[0]) * ((double)g[3] - g[1]);
  double I = iw * ih, U = ap + ag - I;
  double a = (U + I) / (U * U), c = I / (U * U);
  double m2x = p[2] < g[2] ? 1.0 : (p[2] == g[2] ? 0.5 : 0.0);
  double m1x = p[0] > g[0] ? 1.0 : (p[0] == g[0] ? 0.5 : 0.0);
  double m2y = p[3] < g[3] ? 1.0 : (p[3] == g[3] ? 0.5 : 0.0);
  double m1y = p[1] > g[1] ? 1.0 : (p[1] == g[1] ? 0.5 : 0.0);
  dg[0] = a * ih * (m2x - m1x) / S;
  dg[1] = a * iw * (m2y - m1y) / S;
  dg[2] = a * ih * (m2x + m1x) * 0.5 - c * ph;
  dg[3] = a * iw * (m2y + m1y) * 0.5 - c * pw;
}

/* ------------------------------------------------------------------------------------------ */
/* loss forward + backward : v1Loss.py:22-118 and its autograd backward (train.py:171)         */
/* ------------------------------------------------------------------------------------------ */

static inline int64_t off4(const int64_t s[4], int64_t n, int i, int j, int c) {
  return n * s[0] + i * s[1] + j * s[2] + c * s[3];
}

/*
 * pred / grad: fp32 with element strides ps[4] over (n,i,j,channel); target: fp32 with strides ts[4].
 * Channel layout (v1Loss.py:24-25): [conf x B, (x,y,w,h) x B, cls x C].
 * terms[5] = {loc, contain, not_contain, cls}/batch_size (as logged at v1Loss.py:108) and the total
 * (v1Loss.py:104-105).  coord_mode 0 = the reference's row-slice behaviour at v1Loss.py:101
 * (`x[mask][:2]` slices ROWS: the first two responsible boxes of the call get plain squared error on
 * x,y,w,h, all later ones get squared error of square roots on x,y,w,h); 1 = the YOLO paper form.
 * grad may be NULL (forward only).  Returns 0, or -1 on bad arguments.
 */
int yolo1_oracle_loss(const float* pred, const int64_t ps[4], const float* target, const int64_t ts[4],
                      float* grad, float terms[5], int64_t N, int S, int B, int C, float lambda_coord,
                      float lambda_noobj, float batch_size, int coord_mode, int nthreads) {
  if (!pred || !target || !terms || N < 0 || S <= 0 || B <= 0 || B > ORACLE_MAX_B || C < 0) return -1;
  const int D = 5 * B + C;
  const int64_t cells = N * S * S;
  const double inv_bs = 1.0 / (double)batch_size;
  const double lc = lambda_coord, ln = lambda_noobj;

  /* rank boundary for the row-slice quirk: flat index of the 2nd object cell in row-major (n,i,j)
   * order (v1Loss.py:47 torch.nonzero order == boolean-mask gather order at :101). */
  int64_t second_obj = INT64_MAX;
  {
    int seen = 0;
    for (int64_t q = 0; q < cells && seen < 2; ++q) {
      int64_t n = q / (S * S);
      int r = (int)(q % (S * S));
      if (target[off4(ts, n, r / S, r % S, 0)] == 1.0f) {   /* v1Loss.py:28 : == 1 on channel 0 */
        if (++seen == 2) second_obj = q;
      }
    }
  }

  double s_loc = 0, s_hit = 0, s_miss = 0, s_cls = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static) reduction(+ : s_loc, s_hit, s_miss, s_cls)
#endif
  for (int64_t n = 0; n < N; ++n) {
    for (int i = 0; i < S; ++i)
      for (int j = 0; j < S; ++j) {
        const int64_t q = (n * S + i) * S + j;
        const float* P = pred + off4(ps, n, i, j, 0);
        const float* T = target + off4(ts, n, i, j, 0);
        float* G = grad ? grad + off4(ps, n, i, j, 0) : NULL;
        const int64_t pc = ps[3], tc = ts[3];
        const int obj = (T[0] == 1.0f);
        if (!obj) {
          /* v1Loss.py:91 : every slot of a cell without object contributes conf^2 (target 0). */
          for (int b = 0; b < B; ++b) {
            double cf = P[b * pc];
            s_miss += cf * cf;
            if (G) G[b * pc] = (float)(2.0 * ln * cf * inv_bs);
          }
          if (G)
            for (int c = B; c < D; ++c) G[c * pc] = 0.0f;
          continue;
        }
        /* v1Loss.py:66-74 : IoU of each predictor with GT slot 0, first arg-max wins */
        float gt0[4], gxy[4];
        for (int d = 0; d < 4; ++d) gt0[d] = T[(B + d) * tc];
        cxcywh_to_xyxy(gt0, (float)S, gxy);
        int r = 0;
        float best = 0;
        for (int b = 0; b < B; ++b) {
          float pb[4], pxy[4];
          for (int d = 0; d < 4; ++d) pb[d] = P[(B + 4 * b + d) * pc];
          cxcywh_to_xyxy(pb, (float)S, pxy);
          float v = iou_xyxy(pxy, gxy);
          if (b == 0 || v > best) { best = v; r = b; }
        }
        /* class term, v1Loss.py:33-41 */
        for (int c = 0; c < C; ++c) {
          double d = (double)P[(5 * B + c) * pc] - (double)T[(5 * B + c) * tc];
          s_cls += d * d;
          if (G) G[(5 * B + c) * pc] = (float)(2.0 * d * inv_bs);
        }
        /* confidences, v1Loss.py:90-91 */
        double dconf = 0;
        for (int b = 0; b < B; ++b) {
          double cf = P[b * pc];
          if (b == r) {
            dconf = cf - (double)best;
            s_hit += dconf * dconf;
            if (G) G[b * pc] = (float)(2.0 * dconf * inv_bs);
          } else {
            s_miss += cf * cf;
            if (G) G[b * pc] = (float)(2.0 * ln * cf * inv_bs);
          }
        }
        /* coordinates, v1Loss.py:94-101 : GT slot r; IoU path of v1Loss.py:78 stays in the graph */
        float pr[4];
        for (int d = 0; d < 4; ++d) pr[d] = P[(B + 4 * r + d) * pc];
        double dI[4];
        iou_grad(pr, gt0, (float)S, dI);
        for (int b = 0; b < B; ++b) {
          if (b == r) continue;
          if (G)
            for (int d = 0; d < 4; ++d) G[(B + 4 * b + d) * pc] = 0.0f;
        }
        for (int d = 0; d < 4; ++d) {
          double p = pr[d], g = T[(B + 4 * r + d) * tc];
          int plain = coord_mode == 0 ? (q <= second_obj) : (d < 2);
          double gl;
          if (plain) {
            double e = p - g;
            s_loc += e * e;
            gl = 2.0 * lc * e;
          } else {
            /* torch.sqrt in fp32 (v1Loss.py:101) */
            double sp = sqrtf((float)p), sg = sqrtf((float)g);
            double e = sp - sg;
            s_loc += e * e;
            gl = lc * e / sp;
          }
          if (G) G[(B + 4 * r + d) * pc] = (float)((gl - 2.0 * dconf * dI[d]) * inv_bs);
        }
      }
  }
  terms[0] = (float)(s_loc * inv_bs);
  terms[1] = (float)(s_hit * inv_bs);
  terms[2] = (float)(s_miss * inv_bs);
  terms[3] = (float)(s_cls * inv_bs);
  terms[4] = (float)((lc * s_loc + s_hit + ln * s_miss + s_cls) * inv_bs);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* decode : utils/utils.py:94-141.  Bit-exact fp32; every operation rounds separately.          */
/* ------------------------------------------------------------------------------------------ */

/*
 * One image. pred has element strides st[3] over (i,j,channel).  Emits candidates in row-major
 * (i,j,b) order: boxes[k*4..] xyxy, scores[k], cls[k]; returns the count (0 => the caller builds the
 * sentinel of utils/utils.py:134-137).
 *   :108-114 candidate iff conf > fl32(1e-4) or conf == max conf of the image.  `contain.max()` is ATen's
 *            NaN-propagating max: one NaN confidence makes it NaN and `contain == max` selects nothing
 *   :121-126 cx = x*cs + j*cs (cs = fl32(1/S)), xyxy = c -/+ 0.5*wh
 *   :127     first arg-max over the C class channels (torch.max(dim): a NaN class score wins and sticks, so
 *            the slot's score is NaN and `float(score) > thresh` drops it)
 *   :129     emit iff (double)(conf*maxprob) > thresh   (thresh is a Python double)
 */
int yolo1_oracle_decode_image(const float* pred, const int64_t st[3], int S, int B, int C, double thresh,
                              float* boxes, float* scores, int32_t* cls) {
  const float cs = (float)(1.0 / (double)S);
  const float eps = 0.0001f;
  float mx = -INFINITY;
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j)
      for (int b = 0; b < B; ++b) {
        float v = pred[i * st[0] + j * st[1] + b * st[2]];
        if (v > mx || v != v) mx = v; /* NaN sticks: nothing compares greater than it afterwards */
      }
  int k = 0;
  for (int i = 0; i < S; ++i)
    for (int j = 0; j < S; ++j) {
      const float* P = pred + i * st[0] + j * st[1];
      /* class arg-max is per cell, shared by the B slots */
      int best_c = 0;
      float best_p = C > 0 ? P[(5 * B) * st[2]] : 0.0f;
      for (int c = 1; c < C; ++c) {
        float v = P[(5 * B + c) * st[2]];
        if (v > best_p || (v != v && best_p == best_p)) { best_p = v; best_c = c; } /* first NaN wins */
      }
      for (int b = 0; b < B; ++b) {
        float conf = P[b * st[2]];
        if (!(conf > eps || conf == mx)) continue;
        const float* bx = P + (B + 4 * b) * st[2];
        float x = bx[0], y = bx[st[2]], w = bx[2 * st[2]], h = bx[3 * st[2]];
        float cx = x * cs + (float)j * cs;
        float cy = y * cs + (float)i * cs;
        float hw = 0.5f * w, hh = 0.5f * h;
        float score = conf * best_p;
        if ((double)score > thresh) {
          boxes[4 * k + 0] = cx - hw;
          boxes[4 * k + 1] = cy - hh;
          boxes[4 * k + 2] = cx + hw;
          boxes[4 * k + 3] = cy + hh;
          scores[k] = score;
          cls[k] = best_c;
          ++k;
        }
      }
    }
  return k;
}

/* ------------------------------------------------------------------------------------------ */
/* nms : utils/utils.py:150-184.  Class-agnostic greedy suppression, bit-exact fp32.            */
/* ------------------------------------------------------------------------------------------ */

typedef struct { float s; int32_t i; } sc_idx;
static int cmp_desc(const void* a, const void* b) {
  const sc_idx *x = (const sc_idx*)a, *y = (const sc_idx*)b;
  /* torch.sort(descending=True) orders NaN before every number */
  const int xn = x->s != x->s, yn = y->s != y->s;
  if (xn != yn) return xn ? -1 : 1;
  if (x->s > y->s) return -1;
  if (x->s < y->s) return 1;
  return (x->i > y->i) - (x->i < y->i); /* canonical tie order: lower input index first */
}

/*
 * keep[] receives indices into the input in descending score order; returns their count.
 *   :159 area = (x2-x1)*(y2-y1)      :161 sort descending (ties: lower index first -- the reference's
 *   own tie order is machine dependent)
 *   :163-183 take the head, IoU(head, every later live box); a box survives iff ovr <= fl32(thr)
 *   (so NaN dies).  cls may be NULL (class-agnostic, the reference behaviour); with cls != NULL and
 *   per_class != 0 only boxes of the same class suppress each other (the reference nms run per class
 *   subset and merged by score).
 */
int yolo1_oracle_nms(const float* boxes, const float* scores, const int32_t* cls, int n, float thr,
                     int per_class, int32_t* keep) {
  if (n <= 0) return 0;
  sc_idx* ord = (sc_idx*)malloc(sizeof(sc_idx) * (size_t)n);
  float* area = (float*)malloc(sizeof(float) * (size_t)n);
  unsigned char* dead = (unsigned char*)calloc((size_t)n, 1);
  for (int i = 0; i < n; ++i) {
    ord[i].s = scores[i];
    ord[i].i = i;
    area[i] = (boxes[4 * i + 2] - boxes[4 * i + 0]) * (boxes[4 * i + 3] - boxes[4 * i + 1]);
  }
  qsort(ord, (size_t)n, sizeof(sc_idx), cmp_desc);
  int k = 0;
  for (int a = 0; a < n; ++a) {
    if (dead[a]) continue;
    const int ia = ord[a].i;
    keep[k++] = ia;
    const float* A = boxes + 4 * ia;
    for (int b = a + 1; b < n; ++b) {
      if (dead[b]) continue;
      const int ib = ord[b].i;
      if (per_class && cls && cls[ia] != cls[ib]) continue;
      const float* Bx = boxes + 4 * ib;
      float xx1 = Bx[0] < A[0] ? A[0] : Bx[0]; /* clamp(min=x1[i]) */
      float yy1 = Bx[1] < A[1] ? A[1] : Bx[1];
      float xx2 = Bx[2] > A[2] ? A[2] : Bx[2]; /* clamp(max=x2[i]) */
      float yy2 = Bx[3] > A[3] ? A[3] : Bx[3];
      float w = xx2 - xx1, h = yy2 - yy1;
      if (w < 0) w = 0;
      if (h < 0) h = 0;
      float inter = w * h;
      float ovr = inter / ((area[ia] + area[ib]) - inter);
      if (!(ovr <= thr)) dead[b] = 1;
    }
  }
  free(ord); free(area); free(dead);
  return k;
}

/* ------------------------------------------------------------------------------------------ */
/* batched decode + nms (what run_test_mAP / eval.py do image by image, utils/utils.py:405, eval.py:94) */
/* ------------------------------------------------------------------------------------------ */

/*
 * pred [N,S,S,D] with element strides st[4].  Per image n writes up to max_n = S*S*B detections in
 * descending score order at out_*[n*max_n + k] and their number at counts[n] (0 = the reference would
 * return its all-zero sentinel).  cand_counts (optional) receives the pre-NMS candidate count;
 * keep_idx (optional) the kept candidate indices (row-major emission order, as nms() returns them).
 */
int yolo1_oracle_decode_nms(const float* pred, const int64_t st[4], int64_t N, int S, int B, int C,
                            double thresh, float nms_thr, int per_class, float* out_boxes,
                            float* out_scores, int32_t* out_cls, int32_t* counts, int32_t* cand_counts,
                            int32_t* keep_idx, int nthreads) {
  if (!pred || N < 0 || S <= 0 || B <= 0 || C <= 0) return -1;
  const int max_n = S * S * B;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
  {
    float* bx = (float*)malloc(sizeof(float) * 4 * (size_t)max_n);
    float* sc = (float*)malloc(sizeof(float) * (size_t)max_n);
    int32_t* cl = (int32_t*)malloc(sizeof(int32_t) * (size_t)max_n);
    int32_t* kp = (int32_t*)malloc(sizeof(int32_t) * (size_t)max_n);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t n = 0; n < N; ++n) {
      int m = yolo1_oracle_decode_image(pred + n * st[0], st + 1, S, B, C, thresh, bx, sc, cl);
      int k = yolo1_oracle_nms(bx, sc, cl, m, nms_thr, per_class, kp);
      if (cand_counts) cand_counts[n] = m;
      counts[n] = k;
      for (int t = 0; t < k; ++t) {
        const int src = kp[t];
        const int64_t dst = n * max_n + t;
        memcpy(out_boxes + 4 * dst, bx + 4 * src, 4 * sizeof(float));
        out_scores[dst] = sc[src];
        out_cls[dst] = cl[src];
        if (keep_idx) keep_idx[dst] = src;
      }
    }
    free(bx); free(sc); free(cl); free(kp);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* encoder : utils/YOLODataLoader.py:200-230.  Bit-exact fp32, every operation rounds separately. */
/* ------------------------------------------------------------------------------------------ */

/*
 * One image.  boxes [n,4] = (cx, cy, w, h) normalised to the image, labels [n]; target [S,S,5B+C]
 * contiguous, fully written (zero where no object).  In input order, for every object:
 *   :218-219 ij = ceil(cxcy / cell_size) - 1   (cell_size = fl32(1/S); fp32 division)
 *   :220     the cell is reset to zero, so the LAST object that falls into a cell wins
 *   :221-222 every confidence slot = 1, class one-hot
 *   :223-224 delta_xy = (cxcy - ij*cell_size) / cell_size
 *   :225-227 the same (delta_x, delta_y, w, h) in every box slot
 * Python indexing: ij = -1 (cx or cy == 0) addresses the last row / column.  Returns 0, or -2 when an
 * index falls outside [-S, S) (the reference raises IndexError there); the target is then unspecified.
 */
int yolo1_oracle_encode_image(const float* boxes, const int32_t* labels, int n, int S, int B, int C,
                              float* target) {
  const int D = 5 * B + C;
  const float cs = (float)(1.0 / (double)S);
  memset(target, 0, sizeof(float) * (size_t)S * S * D);
  for (int k = 0; k < n; ++k) {
    const float cx = boxes[4 * k], cy = boxes[4 * k + 1], w = boxes[4 * k + 2], h = boxes[4 * k + 3];
    const float fi = ceilf(cx / cs) - 1.0f, fj = ceilf(cy / cs) - 1.0f;
    int col = (int)fi, row = (int)fj;
    if (col < -S || col >= S || row < -S || row >= S) return -2;
    if (labels[k] < -C || labels[k] >= C) return -2;
    if (col < 0) col += S;
    if (row < 0) row += S;
    float* t = target + ((size_t)row * S + col) * D;
    for (int c = 0; c < D; ++c) t[c] = 0.0f;
    for (int b = 0; b < B; ++b) t[b] = 1.0f;
    t[5 * B + (labels[k] < 0 ? labels[k] + C : labels[k])] = 1.0f;
    const float dx = (cx - fi * cs) / cs, dy = (cy - fj * cs) / cs;
    for (int b = 0; b < B; ++b) {
      t[B + 4 * b] = dx;
      t[B + 4 * b + 1] = dy;
      t[B + 4 * b + 2] = w;
      t[B + 4 * b + 3] = h;
    }
  }
  return 0;
}

/* batched: offsets [N+1] into boxes/labels (CSR); target [N,S,S,5B+C] */
int yolo1_oracle_encode(const float* boxes, const int32_t* labels, const int64_t* offsets, int64_t N, int S,
                        int B, int C, float* target) {
  const size_t img = (size_t)S * S * (5 * B + C);
  for (int64_t n = 0; n < N; ++n) {
    int rc = yolo1_oracle_encode_image(boxes + 4 * offsets[n], labels + offsets[n],
                                       (int)(offsets[n + 1] - offsets[n]), S, B, C, target + n * img);
    if (rc) return rc;
  }
  return 0;
}

int yolo1_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
