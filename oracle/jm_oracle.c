/*
 * oracle/jm_oracle.c -- TEST INFRASTRUCTURE ONLY.  See jm_oracle.h.
 * Plain-C restatement of JM 19.0's ME + transform/quant leaf algorithms (file:line cited per
 * function, relative to /root/reference).  Written as clamped-index arithmetic rather than JM's
 * pointer walking; checked bit-for-bit against the real JM functions (oracle/_ref/libjmref.so).
 */
#include "jm_oracle.h"
#include <stdlib.h>
#include <string.h>

static inline int iclip(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int iabs_(int v) { return v < 0 ? -v : v; }

/* ------------------------------------------------------------------------------------------
 * Quarter-pel reference planes.  lencod/src/img_luma.c:611-680 (getSubImagesLuma) and helpers
 * :40-596.  All taps index the PADDED buffer with the index clamped to the padded extent
 * (img_luma.c:170-237 horizontal, :272-331 vertical); the centre plane [2][2] is filtered
 * vertically from the un-rounded horizontal intermediates (:347-423).
 * ---------------------------------------------------------------------------------------- */
jmo_ref *jmo_ref_create(const uint16_t *luma, int w, int h, int stride, int max_value)
{
  jmo_ref *r = (jmo_ref *)calloc(1, sizeof(*r));
  int W = w + 2 * JMO_PAD_X, H = h + 2 * JMO_PAD_Y;
  r->w = w; r->h = h; r->W = W; r->H = H;
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++)
      r->plane[a][b] = (uint16_t *)malloc((size_t)W * H * sizeof(uint16_t));
  int *tmp = (int *)malloc((size_t)W * H * sizeof(int));
  uint16_t *G = r->plane[0][0], *B = r->plane[0][2], *Hh = r->plane[2][0], *J = r->plane[2][2];

  /* integer plane: edge replication, img_luma.c:40-85 */
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      G[y * W + x] = luma[(size_t)iclip(0, h - 1, y - JMO_PAD_Y) * stride + iclip(0, w - 1, x - JMO_PAD_X)];

#define CX(x) iclip(0, W - 1, (x))
#define CY(y) iclip(0, H - 1, (y))
  /* horizontal six-tap (20,-5,1), img_luma.c:151-245; un-rounded copy kept in tmp (:168) */
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const uint16_t *g = G + y * W;
      int is = 20 * (g[x] + g[CX(x + 1)]) - 5 * (g[CX(x - 1)] + g[CX(x + 2)]) + (g[CX(x - 2)] + g[CX(x + 3)]);
      tmp[y * W + x] = is;
      B[y * W + x] = (uint16_t)iclip(0, max_value, (is + 16) >> 5);
    }
  /* vertical six-tap, img_luma.c:257-339, and vertical over tmp for [2][2], :347-431 */
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      int ya = y, yd = CY(y + 1), yb = CY(y - 1), ye = CY(y + 2), yc = CY(y - 2), yf = CY(y + 3);
      int is = 20 * (G[ya * W + x] + G[yd * W + x]) - 5 * (G[yb * W + x] + G[ye * W + x]) + (G[yc * W + x] + G[yf * W + x]);
      Hh[y * W + x] = (uint16_t)iclip(0, max_value, (is + 16) >> 5);
      int it = 20 * (tmp[ya * W + x] + tmp[yd * W + x]) - 5 * (tmp[yb * W + x] + tmp[ye * W + x]) + (tmp[yc * W + x] + tmp[yf * W + x]);
      J[y * W + x] = (uint16_t)iclip(0, max_value, (it + 512) >> 10);
    }
  /* bilinear quarter-pel planes, img_luma.c:440-596 and the source pairs at :647-678 */
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      int i = y * W + x, xr = y * W + CX(x + 1), yd = CY(y + 1) * W + x;
#define AVG(a, b) (uint16_t)(((a) + (b) + 1) >> 1)
      r->plane[0][1][i] = AVG(G[i], B[i]);
      r->plane[1][0][i] = AVG(G[i], Hh[i]);
      r->plane[1][1][i] = AVG(B[i], Hh[i]);
      r->plane[1][2][i] = AVG(B[i], J[i]);
      r->plane[2][1][i] = AVG(Hh[i], J[i]);
      r->plane[0][3][i] = AVG(B[i], G[xr]);
      r->plane[1][3][i] = AVG(B[i], Hh[xr]);
      r->plane[2][3][i] = AVG(J[i], Hh[xr]);
      r->plane[3][0][i] = AVG(Hh[i], G[yd]);
      r->plane[3][1][i] = AVG(Hh[i], B[yd]);
      r->plane[3][2][i] = AVG(J[i], B[yd]);
      r->plane[3][3][i] = AVG(B[yd], Hh[xr]);
#undef AVG
    }
#undef CX
#undef CY
  free(tmp);
  return r;
}

void jmo_ref_destroy(jmo_ref *r)
{
  if (!r) return;
  for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) free(r->plane[a][b]);
  free(r);
}

void jmo_ref_get_plane(const jmo_ref *r, int fy, int fx, uint16_t *out)
{
  memcpy(out, r->plane[fy][fx], (size_t)r->W * r->H * sizeof(uint16_t));
}

/* UMVLine4X, lencod/inc/refbuf.h:22-26: only the block ORIGIN is clamped, to
 * [-PAD, size_pad] with size_x_pad = w + PAD_X - 1 - 16 (mbuffer.c:564-565); rows then run linearly. */
static inline const uint16_t *umv_line(const jmo_ref *r, int qy, int qx)
{
  int iy = iclip(-JMO_PAD_Y, r->h + JMO_PAD_Y - 1 - 16, qy >> 2);
  int ix = iclip(-JMO_PAD_X, r->w + JMO_PAD_X - 1 - 16, qx >> 2);
  return r->plane[qy & 3][qx & 3] + (size_t)(iy + JMO_PAD_Y) * r->W + (ix + JMO_PAD_X);
}

/* square spiral, lencod/src/mv_search.c:406-442 */
void jmo_spiral(int R, int16_t *xy)
{
  int k = 1;
  xy[0] = xy[1] = 0;
  for (int l = 1; l <= (R > 1 ? R : 1); l++) {
    for (int i = -l + 1; i < l; i++) {
      xy[2 * k] = (int16_t)i; xy[2 * k + 1] = (int16_t)-l; k++;
      xy[2 * k] = (int16_t)i; xy[2 * k + 1] = (int16_t)l;  k++;
    }
    for (int i = -l; i <= l; i++) {
      xy[2 * k] = (int16_t)-l; xy[2 * k + 1] = (int16_t)i; k++;
      xy[2 * k] = (int16_t)l;  xy[2 * k + 1] = (int16_t)i; k++;
    }
  }
}

/* mvbits[], lencod/src/mv_search.c:366-374: 1 for 0, else 2*floor(log2|v|)+3 */
int jmo_mvbits(int v)
{
  int a = iabs_(v), n = 0;
  if (!a) return 1;
  while (a >> (n + 1)) n++;
  return 2 * n + 3;
}

/* HadamardSAD4x4, lencod/src/me_distortion.c:175-258: 2-D 4x4 Hadamard, (sum|.| + 1) >> 1 */
int jmo_hadamard_sad4x4(const int16_t *d)
{
  int m[16], t[16], s = 0;
  for (int c = 0; c < 4; c++) {           /* columns */
    int a0 = d[c] + d[12 + c], a1 = d[4 + c] + d[8 + c], a2 = d[4 + c] - d[8 + c], a3 = d[c] - d[12 + c];
    m[c] = a0 + a1; m[8 + c] = a0 - a1; m[4 + c] = a3 + a2; m[12 + c] = a3 - a2;
  }
  for (int r = 0; r < 4; r++) {           /* rows */
    const int *p = m + 4 * r;
    int a0 = p[0] + p[3], a1 = p[1] + p[2], a2 = p[1] - p[2], a3 = p[0] - p[3];
    t[4 * r] = a0 + a1; t[4 * r + 1] = a0 - a1; t[4 * r + 2] = a2 + a3; t[4 * r + 3] = a3 - a2;
  }
  for (int k = 0; k < 16; k++) s += iabs_(t[k]);
  return (s + 1) >> 1;
}

/* HadamardSAD8x8, lencod/src/me_distortion.c:266-341: (sum|.| + 2) >> 2 */
int jmo_hadamard_sad8x8(const int16_t *d)
{
  int a[64], s = 0;
  for (int i = 0; i < 64; i++) a[i] = d[i];
  for (int pass = 0; pass < 2; pass++) {
    int step = pass ? 8 : 1, line = pass ? 1 : 8;
    for (int l = 0; l < 8; l++) {
      int *p = a + l * line;
      for (int len = 4; len >= 1; len >>= 1)     /* butterflies of span 4, 2, 1 */
        for (int b = 0; b < 8; b += 2 * len)
          for (int k = 0; k < len; k++) {
            int u = p[(b + k) * step], v = p[(b + k + len) * step];
            p[(b + k) * step] = u + v; p[(b + k + len) * step] = u - v;
          }
    }
  }
  for (int i = 0; i < 64; i++) s += iabs_(a[i]);
  return (s + 2) >> 2;
}

/* computeSAD  lencod/src/me_distortion.c:349-426  (partition origin clamped once)
 * computeSSE  lencod/src/me_distortion.c:1190+     (same addressing, squared differences)
 * computeSATD lencod/src/me_distortion.c:745-825   (EVERY 4x4 / 8x8 sub-block origin clamped)
 * Early termination in JM only ever turns a losing candidate into "min_mcost" (mv_search.h:19),
 * which the strict '<' tests reject, so the complete sum decides identically. */
int jmo_dist(const jmo_ref *r, const uint16_t *src, int bsx, int bsy, int cx, int cy, int metric, int test8x8)
{
  int acc = 0;
  if (metric != JMO_SATD) {
    const uint16_t *ref = umv_line(r, cy, cx);
    for (int y = 0; y < bsy; y++)
      for (int x = 0; x < bsx; x++) {
        int d = src[y * bsx + x] - ref[(size_t)y * r->W + x];
        acc += (metric == JMO_SAD) ? iabs_(d) : d * d;
      }
    return acc;
  }
  int n = test8x8 ? 8 : 4;
  int16_t diff[64];
  for (int by = 0; by < bsy; by += n)
    for (int bx = 0; bx < bsx; bx += n) {
      const uint16_t *ref = umv_line(r, cy + (by << 2), cx + (bx << 2));
      for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
          diff[y * n + x] = (int16_t)(src[(by + y) * bsx + bx + x] - ref[(size_t)y * r->W + x]);
      acc += test8x8 ? jmo_hadamard_sad8x8(diff) : jmo_hadamard_sad4x4(diff);
    }
  return acc;
}

/* Prediction sample of the weighted / bi-predictive distortions:
 * computeSADWP me_distortion.c:434 (also SATDWP :833, SSEWP :1261):   clip(((w * ref + round) >> denom) + offset)
 * computeBiPredSAD1 :525 (SATD1 :943, SSE1 :1353):                    (ref1 + ref2 + 1) >> 1
 * computeBiPredSAD2 :624 (SATD2 :1038, SSE2 :1438):                   clip(((w1*ref1 + w2*ref2 + 2*round) >> (denom+1)) + offsetBi)
 * Clamps as in the plain functions, applied to each reference on its own. */
static inline int pred_sample(int form, int a, int b, const int *wp, int maxv)
{
  switch (form) {
  case 1:  return iclip(0, maxv, ((wp[0] * a + wp[4]) >> wp[3]) + wp[2]);
  case 2:  return (a + b + 1) >> 1;
  case 3:  return iclip(0, maxv, ((wp[0] * a + wp[1] * b + 2 * wp[4]) >> (wp[3] + 1)) + wp[2]);
  default: return a;
  }
}

int jmo_dist_ex(const jmo_ref *r1, const jmo_ref *r2, const uint16_t *src, int bsx, int bsy, int c1x, int c1y, int c2x, int c2y,
                int metric, int test8x8, int form, const int *wp, int maxv)
{
  int acc = 0, two = form >= 2;
  if (metric != JMO_SATD) {
    const uint16_t *a = umv_line(r1, c1y, c1x), *b = two ? umv_line(r2, c2y, c2x) : a;
    for (int y = 0; y < bsy; y++)
      for (int x = 0; x < bsx; x++) {
        int d = src[y * bsx + x] - pred_sample(form, a[(size_t)y * r1->W + x], b[(size_t)y * r1->W + x], wp, maxv);
        acc += (metric == JMO_SAD) ? iabs_(d) : d * d;
      }
    return acc;
  }
  int n = test8x8 ? 8 : 4;
  int16_t diff[64];
  for (int by = 0; by < bsy; by += n)
    for (int bx = 0; bx < bsx; bx += n) {
      const uint16_t *a = umv_line(r1, c1y + (by << 2), c1x + (bx << 2));
      const uint16_t *b = two ? umv_line(r2, c2y + (by << 2), c2x + (bx << 2)) : a;
      /* computeBiPredSATD2's 8x8 branch does not advance the source pointer after the eighth sample of a row
       * (me_distortion.c:1166 "*d++ = (short) ((*src_line) - weighted_pel);"), so row y of a sub-block reads the
       * block-compact source y samples early.  Reproduced: the encoder's decisions depend on it. */
      int slip = (form == 3 && test8x8) ? 1 : 0;
      for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
          diff[y * n + x] = (int16_t)(src[(by + y) * bsx + bx + x - slip * y] - pred_sample(form, a[(size_t)y * r1->W + x], b[(size_t)y * r1->W + x], wp, maxv));
      acc += test8x8 ? jmo_hadamard_sad8x8(diff) : jmo_hadamard_sad4x4(diff);
    }
  return acc;
}

static const int k_bs[8][2] = {{16,16},{16,16},{16,8},{8,16},{8,8},{8,4},{4,8},{4,4}}; /* macroblock.h:58-68 */

static void get_block(const uint16_t *cur, int stride, int px, int py, int bsx, int bsy, uint16_t *out)
{
  for (int y = 0; y < bsy; y++)
    memcpy(out + y * bsx, cur + (size_t)(py + y) * stride + px, bsx * sizeof(uint16_t));
}

static inline int64_t mvcost(int lambda, int cx, int cy, int px, int py)   /* mv_search.h:100-112 */
{
  return (int64_t)lambda * (jmo_mvbits(cx - px) + jmo_mvbits(cy - py));
}

/* full_search_motion_estimation, lencod/src/me_fullsearch.c:39-103 (rdopt on: no (0,0) bonus) */
int64_t jmo_full_search(const jmo_ref *r, const uint16_t *cur, int cur_stride, int blocktype,
                        int pos_x, int pos_y, int pred_x, int pred_y, int center_x, int center_y,
                        int lambda, int64_t min_mcost, int R, int16_t *mv_out)
{
  int bsx = k_bs[blocktype][0], bsy = k_bs[blocktype][1];
  int max_pos = (2 * R + 1) * (2 * R + 1), best = 0;
  uint16_t src[256];
  int16_t *sp = (int16_t *)malloc(sizeof(int16_t) * 2 * max_pos);
  jmo_spiral(R, sp);
  get_block(cur, cur_stride, pos_x, pos_y, bsx, bsy, src);
  int ccx = (pos_x << 2) + center_x, ccy = (pos_y << 2) + center_y;
  int ppx = (pos_x << 2) + pred_x, ppy = (pos_y << 2) + pred_y;
  for (int pos = 0; pos < max_pos; pos++) {
    int cx = ccx + 4 * sp[2 * pos], cy = ccy + 4 * sp[2 * pos + 1];
    int64_t mc = mvcost(lambda, cx, cy, ppx, ppy);
    if (mc >= min_mcost) continue;
    mc += (int64_t)jmo_dist(r, src, bsx, bsy, cx, cy, JMO_SAD, 0) << 5;
    if (mc < min_mcost) { min_mcost = mc; best = pos; }
  }
  mv_out[0] = (int16_t)(center_x + 4 * sp[2 * best]);
  mv_out[1] = (int16_t)(center_y + 4 * sp[2 * best + 1]);
  free(sp);
  return min_mcost;
}

/* sub_pel_motion_estimation, lencod/src/me_fullsearch.c:186-289 (rdopt on) */
int64_t jmo_sub_pel(const jmo_ref *r, const uint16_t *cur, int cur_stride, int blocktype,
                    int pos_x, int pos_y, int pred_x, int pred_y, int mv_x, int mv_y,
                    const int *lambda3, int64_t min_mcost, int metric_h, int metric_q,
                    int start_hp, int start_qp, int test8x8, int16_t *mv_out)
{
  int bsx = k_bs[blocktype][0], bsy = k_bs[blocktype][1];
  uint16_t src[256];
  int16_t sp[18];
  jmo_spiral(1, sp);
  get_block(cur, cur_stride, pos_x, pos_y, bsx, bsy, src);
  int best = 0;
  for (int pos = start_hp; pos < 9; pos++) {
    int cx = mv_x + 2 * sp[2 * pos], cy = mv_y + 2 * sp[2 * pos + 1];
    int64_t mc = mvcost(lambda3[1], cx, cy, pred_x, pred_y);
    if (mc >= min_mcost) continue;
    mc += (int64_t)jmo_dist(r, src, bsx, bsy, cx + (pos_x << 2), cy + (pos_y << 2), metric_h, test8x8) << 5;
    if (mc < min_mcost) { min_mcost = mc; best = pos; }
  }
  if (best) { mv_x += 2 * sp[2 * best]; mv_y += 2 * sp[2 * best + 1]; }
  if (!start_qp) min_mcost = (int64_t)0x7fffffff << 5;   /* DISTBLK_MAX, defines.h:136 */
  best = 0;
  for (int pos = start_qp; pos < 9; pos++) {
    int cx = mv_x + sp[2 * pos], cy = mv_y + sp[2 * pos + 1];
    int64_t mc = mvcost(lambda3[2], cx, cy, pred_x, pred_y);
    if (mc >= min_mcost) continue;
    mc += (int64_t)jmo_dist(r, src, bsx, bsy, cx + (pos_x << 2), cy + (pos_y << 2), metric_q, test8x8) << 5;
    if (mc < min_mcost) { min_mcost = mc; best = pos; }
  }
  if (best) { mv_x += sp[2 * best]; mv_y += sp[2 * best + 1]; }
  mv_out[0] = (int16_t)mv_x; mv_out[1] = (int16_t)mv_y;
  return min_mcost;
}

/* search centre of the fast full search, lencod/src/me_fullfast.c:309-327 (rdopt on) */
void jmo_ffs_center(int pmv_x, int pmv_y, int R, const int *hq, const int *vq, int16_t *c)
{
  int sr = R << 2;
  c[0] = (int16_t)iclip(hq[0] + sr, hq[1] - sr, ((pmv_x + 2) >> 2) * 4);
  c[1] = (int16_t)iclip(vq[0] + sr, vq[1] - sr, ((pmv_y + 2) >> 2) * 4);
}

/* setup_fast_full_search, lencod/src/me_fullfast.c:492-556 (the MACROBLOCK origin is clamped once
 * per position, :498) + update_full_search_large_blocks :196-260.
 * block_sad layout [blocktype 0..7][16][max_pos]; slot numbering of each type as JM's. */
void jmo_ffs_setup(const jmo_ref *r, const uint16_t *cur, int cur_stride, int mb_x, int mb_y,
                   int center_x, int center_y, int R, uint32_t *bs)
{
  int max_pos = (2 * R + 1) * (2 * R + 1);
  int16_t *sp = (int16_t *)malloc(sizeof(int16_t) * 2 * max_pos);
  jmo_spiral(R, sp);
#define BS(t, i) (bs + ((size_t)(t) * 16 + (i)) * max_pos)
  for (int pos = 0; pos < max_pos; pos++) {
    int cx = (mb_x << 2) + center_x + 4 * sp[2 * pos], cy = (mb_y << 2) + center_y + 4 * sp[2 * pos + 1];
    const uint16_t *ref = umv_line(r, cy, cx);
    for (int b = 0; b < 16; b++) {
      int bx = (b & 3) * 4, by = (b >> 2) * 4, s = 0;
      for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++)
          s += iabs_((int)ref[(size_t)(by + y) * r->W + bx + x] - (int)cur[(size_t)(mb_y + by + y) * cur_stride + mb_x + bx + x]);
      BS(7, b)[pos] = (uint32_t)s;
    }
  }
  /* larger blocks: the exact slot arithmetic of me_fullfast.c:207-259 */
  for (int pos = 0; pos < max_pos; pos++) {
    for (int i = 0; i < 4; i++) { BS(6, i)[pos] = BS(7, i)[pos] + BS(7, i + 4)[pos]; BS(6, 8 + i)[pos] = BS(7, 8 + i)[pos] + BS(7, 12 + i)[pos]; }
    for (int i = 0; i < 16; i += 2) BS(5, i)[pos] = BS(7, i)[pos] + BS(7, i + 1)[pos];
    BS(4, 0)[pos] = BS(6, 0)[pos] + BS(6, 1)[pos];   BS(4, 2)[pos]  = BS(6, 2)[pos] + BS(6, 3)[pos];
    BS(4, 8)[pos] = BS(6, 8)[pos] + BS(6, 9)[pos];   BS(4, 10)[pos] = BS(6, 10)[pos] + BS(6, 11)[pos];
    BS(3, 0)[pos] = BS(4, 0)[pos] + BS(4, 8)[pos];   BS(3, 2)[pos]  = BS(4, 2)[pos] + BS(4, 10)[pos];
    BS(2, 0)[pos] = BS(4, 0)[pos] + BS(4, 2)[pos];   BS(2, 8)[pos]  = BS(4, 8)[pos] + BS(4, 10)[pos];
    BS(1, 0)[pos] = BS(3, 0)[pos] + BS(3, 2)[pos];
  }
#undef BS
  free(sp);
}

/* fast_full_search_motion_estimation, lencod/src/me_fullfast.c:618-689 (rdopt on) */
int64_t jmo_ffs_search(const uint32_t *bs, int R, int blocktype, int block_index, int center_x, int center_y,
                       int pred_x, int pred_y, int lambda, int64_t min_mcost, int max_mvd, int16_t *mv_out)
{
  int max_pos = (2 * R + 1) * (2 * R + 1), best = 0;
  int16_t *sp = (int16_t *)malloc(sizeof(int16_t) * 2 * max_pos);
  jmo_spiral(R, sp);
  const uint32_t *s = bs + ((size_t)blocktype * 16 + block_index) * max_pos;
  max_mvd -= 1;
  for (int pos = 0; pos < max_pos; pos++) {
    int64_t mc = (int64_t)s[pos] << 5;
    int cx = center_x + 4 * sp[2 * pos], cy = center_y + 4 * sp[2 * pos + 1];
    int mvd = iabs_(cx - pred_x) > iabs_(cy - pred_y) ? iabs_(cx - pred_x) : iabs_(cy - pred_y);
    if (mc < min_mcost && mvd < max_mvd) {
      mc += mvcost(lambda, cx, cy, pred_x, pred_y);
      if (mc < min_mcost) { min_mcost = mc; best = pos; }
    }
  }
  mv_out[0] = (int16_t)(center_x + 4 * sp[2 * best]);
  mv_out[1] = (int16_t)(center_y + 4 * sp[2 * best + 1]);
  free(sp);
  return min_mcost;
}

/* forward4x4, lcommon/src/transform.c:20-68 */
void jmo_forward4x4(int *b)
{
  int t[16];
  for (int i = 0; i < 4; i++) {
    int *p = b + 4 * i;
    int t0 = p[0] + p[3], t1 = p[1] + p[2], t2 = p[1] - p[2], t3 = p[0] - p[3];
    t[4 * i] = t0 + t1; t[4 * i + 1] = (t3 << 1) + t2; t[4 * i + 2] = t0 - t1; t[4 * i + 3] = t3 - (t2 << 1);
  }
  for (int i = 0; i < 4; i++) {
    int t0 = t[i] + t[12 + i], t1 = t[4 + i] + t[8 + i], t2 = t[4 + i] - t[8 + i], t3 = t[i] - t[12 + i];
    b[i] = t0 + t1; b[4 + i] = t2 + (t3 << 1); b[8 + i] = t0 - t1; b[12 + i] = t3 - (t2 << 1);
  }
}

static void fwd8_1d(const int *p, int s, int *o, int os)   /* lcommon/src/transform.c:365-402 */
{
  int a0 = p[0] + p[7 * s], a1 = p[s] + p[6 * s], a2 = p[2 * s] + p[5 * s], a3 = p[3 * s] + p[4 * s];
  int b0 = a0 + a3, b1 = a1 + a2, b2 = a0 - a3, b3 = a1 - a2;
  a0 = p[0] - p[7 * s]; a1 = p[s] - p[6 * s]; a2 = p[2 * s] - p[5 * s]; a3 = p[3 * s] - p[4 * s];
  int b4 = a1 + a2 + ((a0 >> 1) + a0), b5 = a0 - a3 - ((a2 >> 1) + a2);
  int b6 = a0 + a3 - ((a1 >> 1) + a1), b7 = a1 - a2 + ((a3 >> 1) + a3);
  o[0] = b0 + b1; o[os] = b4 + (b7 >> 2); o[2 * os] = b2 + (b3 >> 1); o[3 * os] = b5 + (b6 >> 2);
  o[4 * os] = b0 - b1; o[5 * os] = b6 - (b5 >> 2); o[6 * os] = (b2 >> 1) - b3; o[7 * os] = (b4 >> 2) - b7;
}

/* forward8x8, lcommon/src/transform.c:353-448 */
void jmo_forward8x8(int *b)
{
  int t[64];
  for (int i = 0; i < 8; i++) fwd8_1d(b + 8 * i, 1, t + 8 * i, 1);
  for (int i = 0; i < 8; i++) fwd8_1d(t + i, 8, b + i, 8);
}

/* quant_4x4_normal  lencod/src/quant4x4_normal.c:39-115,  quant_4x4_around  quant4x4_around.c:40-130,
 * quant_8x8_normal  quant8x8_normal.c:43-107, quant_8x8_around quant8x8_around.c:40-115,
 * quant_8x8cavlc_normal quant8x8_normal.c:123-202, quant_8x8cavlc_around quant8x8_around.c:133-223 */
int jmo_quant(int variant, int *coef, int qp, const int *qparams, const uint8_t *scan,
              const uint8_t *c_cost, int is_cavlc, int arw, int *levels, int *runs, int *fadjust, int *coeff_cost)
{
  int n = variant < 2 ? 4 : 8, nn = n * n;
  int around = variant & 1, cavlc8 = variant >= 4;
  int qp_per = qp / 6, q_bits = (n == 4 ? 15 : 16) + qp_per, dq = n == 4 ? 4 : 6;
  int nonzero = 0, nl[4] = {0, 0, 0, 0}, run[4] = {0, 0, 0, 0};
  int clip = (n == 4) ? is_cavlc : cavlc8;
  for (int k = 0; k < nn; k++) {
    int i = scan[2 * k], j = scan[2 * k + 1], idx = j * n + i;
    int s = cavlc8 ? k / 16 : 0;
    int *m7 = coef + idx;
    const int *q = qparams + 3 * idx;
    if (around && fadjust) fadjust[idx] = 0;
    if (*m7 == 0) { run[s]++; continue; }
    int scaled = iabs_(*m7) * q[1];
    int level = (scaled + q[0]) >> q_bits;
    if (level == 0) { *m7 = 0; run[s]++; continue; }
    if (clip && level > 2063) level = 2063;
    if (around && fadjust) fadjust[idx] = (arw * (scaled - (level << q_bits)) + (1 << q_bits)) >> (q_bits + 1);
    *coeff_cost += (level > 1) ? 999999 : c_cost[run[s]];
    if (*m7 < 0) level = -level;
    *m7 = (((level * q[2]) << qp_per) + (1 << (dq - 1))) >> dq;
    levels[17 * s * cavlc8 + nl[s]] = level;
    runs[17 * s * cavlc8 + nl[s]] = run[s];
    nl[s]++; run[s] = 0; nonzero = 1;
  }
  for (int s = 0; s < (cavlc8 ? 4 : 1); s++) levels[17 * s * cavlc8 + nl[s]] = 0;
  return nonzero;
}

/* ------------------------------------------------------------------------------------------
 * Inverse transforms.  lcommon/src/transform.c:70-119 (inverse4x4), :450-547 (inverse8x8).
 * ---------------------------------------------------------------------------------------- */
void jmo_inverse4x4(int *b)
{
  int t[16];
  for (int i = 0; i < 4; i++) {                     /* horizontal */
    const int *p = b + 4 * i;
    int p0 = p[0] + p[2], p1 = p[0] - p[2], p2 = (p[1] >> 1) - p[3], p3 = p[1] + (p[3] >> 1);
    t[4 * i] = p0 + p3; t[4 * i + 1] = p1 + p2; t[4 * i + 2] = p1 - p2; t[4 * i + 3] = p0 - p3;
  }
  for (int i = 0; i < 4; i++) {                     /* vertical */
    int p0 = t[i] + t[8 + i], p1 = t[i] - t[8 + i], p2 = (t[4 + i] >> 1) - t[12 + i], p3 = t[4 + i] + (t[12 + i] >> 1);
    b[i] = p0 + p3; b[4 + i] = p1 + p2; b[8 + i] = p1 - p2; b[12 + i] = p0 - p3;
  }
}

static void inv8_1d(const int *p, int s, int *o, int os)
{
  int a0 = p[0] + p[4 * s], a1 = p[0] - p[4 * s], a2 = p[6 * s] - (p[2 * s] >> 1), a3 = p[2 * s] + (p[6 * s] >> 1);
  int b0 = a0 + a3, b2 = a1 - a2, b4 = a1 + a2, b6 = a0 - a3;
  a0 = -p[3 * s] + p[5 * s] - p[7 * s] - (p[7 * s] >> 1);
  a1 =  p[s] + p[7 * s] - p[3 * s] - (p[3 * s] >> 1);
  a2 = -p[s] + p[7 * s] + p[5 * s] + (p[5 * s] >> 1);
  a3 =  p[3 * s] + p[5 * s] + p[s] + (p[s] >> 1);
  int b1 = a0 + (a3 >> 2), b3 = a1 + (a2 >> 2), b5 = a2 - (a1 >> 2), b7 = a3 - (a0 >> 2);
  o[0] = b0 + b7; o[os] = b2 - b5; o[2 * os] = b4 + b3; o[3 * os] = b6 + b1;
  o[4 * os] = b6 - b1; o[5 * os] = b4 - b3; o[6 * os] = b2 + b5; o[7 * os] = b0 - b7;
}

void jmo_inverse8x8(int *b)
{
  int t[64];
  for (int i = 0; i < 8; i++) inv8_1d(b + 8 * i, 1, t + 8 * i, 1);
  for (int i = 0; i < 8; i++) inv8_1d(t + i, 8, b + i, 8);
}

/* ------------------------------------------------------------------------------------------
 * Inter luma residual coding of one macroblock: luma_residual_coding (lencod/src/macroblock.c:1182-1257) for a
 * non-skipped inter macroblock of a P slice, i.e. per 8x8 quadrant luma_residual_coding_8x8 / _16x16 (:832-1017):
 * residual -> forward transform -> quantisation (quant_4x4_normal / quant_8x8_normal) -> per block inverse transform +
 * sample_reconstruct (lcommon/src/blk_prediction.c:48, (r + 32) >> 6 + pred, clipped) when a level is nonzero, else the
 * prediction; quadrants whose coefficient cost is <= _LUMA_COEFF_COST_ (4) are reset (reset_block :806: levels zeroed,
 * cbp bits cleared, prediction copied); if the macroblock's summed cost is <= _LUMA_MB_COEFF_COST_ (5) the luma cbp is
 * cleared and the whole prediction copied (levels are left as they are: JM's memset is commented out, :1254).
 *   src, pred : 16x16 samples;  n = 4 | 8;  levels: 256 in scan order per block (4x4: block by*4+bx at [b*16], 8x8: b8 at [b8*64])
 *   out: cost8[4] (after the resets), *cbp (bits 0..3), *cbp_blk (bits 0..15), recon 16x16, return SSE(src, recon)
 * ---------------------------------------------------------------------------------------- */
long long jmo_luma_residual_coding(const uint16_t *src, const uint16_t *pred, int n, int qp, const int *qparams,
                                   const uint8_t *scan, const uint8_t *c_cost, int is_cavlc, int max_value,
                                   short *levels, int *cost8, int *cbp, int *cbp_blk, uint16_t *recon)
{
  int sum = 0;
  *cbp = 0; *cbp_blk = 0;
  memset(levels, 0, 256 * sizeof(short));
  for (int i = 0; i < 256; i++) recon[i] = pred[i];
  for (int b8 = 0; b8 < 4; b8++) {
    const int qy = (b8 >> 1) * 8, qx = (b8 & 1) * 8;
    int cost = 0;
    for (int sb = 0; sb < (n == 4 ? 4 : 1); sb++) {
      const int by = qy + (n == 4 ? (sb >> 1) * 4 : 0), bx = qx + (n == 4 ? (sb & 1) * 4 : 0);
      int blk[64], lv[68], rn[68];
      for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++) blk[y * n + x] = (int)src[(by + y) * 16 + bx + x] - (int)pred[(by + y) * 16 + bx + x];
      if (n == 4) jmo_forward4x4(blk); else jmo_forward8x8(blk);
      /* 8x8 + CAVLC: residual_transform_quant_luma_8x8_cavlc -> quant_8x8cavlc_normal, four interleaved lists (transform8x8.c:604) */
      const int variant = n == 4 ? 0 : (is_cavlc ? 4 : 2);
      int nz = jmo_quant(variant, blk, qp, qparams, scan, c_cost, is_cavlc, 0, lv, rn, NULL, &cost);
      short *out = levels + (n == 4 ? ((by >> 2) * 4 + (bx >> 2)) * 16 : b8 * 64);
      for (int s = 0; s < (variant == 4 ? 4 : 1); s++)
        for (int k = 16 * s * (variant == 4), i = 17 * s * (variant == 4); lv[i] != 0; i++) { k += rn[i]; out[k++] = (short)lv[i]; }
      if (nz) {
        if (n == 4) { jmo_inverse4x4(blk); *cbp_blk |= 1 << ((by >> 2) * 4 + (bx >> 2)); }
        else { jmo_inverse8x8(blk); *cbp_blk |= 51 << (4 * b8 - 2 * (b8 & 1)); }
        *cbp |= 1 << b8;
        for (int y = 0; y < n; y++)
          for (int x = 0; x < n; x++)
            recon[(by + y) * 16 + bx + x] = (uint16_t)iclip(0, max_value, ((blk[y * n + x] + 32) >> 6) + (int)pred[(by + y) * 16 + bx + x]);
      }
    }
    if (cost <= 4) {                                  /* reset_block */
      cost = 0;
      *cbp &= 63 - (1 << b8);
      *cbp_blk &= ~(51 << (4 * b8 - 2 * (b8 & 1)));
      for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) recon[(qy + y) * 16 + qx + x] = pred[(qy + y) * 16 + qx + x];
      if (n == 4) { for (int sb = 0; sb < 4; sb++) memset(levels + (((qy >> 2) + (sb >> 1)) * 4 + (qx >> 2) + (sb & 1)) * 16, 0, 16 * sizeof(short)); }
      else memset(levels + b8 * 64, 0, 64 * sizeof(short));
    }
    cost8[b8] = cost;
    sum += cost;
  }
  if (sum <= 5) {
    *cbp &= 0xfffff0; *cbp_blk &= 0xff0000;
    for (int i = 0; i < 256; i++) recon[i] = pred[i];
  }
  long long sse = 0;
  for (int i = 0; i < 256; i++) { int d = (int)src[i] - (int)recon[i]; sse += d * d; }
  return sse;
}

/* ------------------------------------------------------------------------------------------
 * List quantiser = quant_ac4x4_normal/_around (lencod/src/quant4x4_normal.c:117-196, quant4x4_around.c:132-215),
 * quant_dc4x4_normal (quant4x4_normal.c:200-259), quant_dc2x2_normal/_around and quant_dc4x2_normal/_around
 * (lencod/src/quantChroma_normal.c:37-170, quantChroma_around.c): one loop over m coefficients in scan order.
 *   params[k] = {Offset (doubled by the caller for the DC forms), Scale, InvScale};  dequant: 0 level, 1 (level*Inv)<<qp_per,
 *   2 ((level*Inv)<<qp_per + 8)>>4.  levels/runs [17], fadjust [m] (around), *coeff_cost accumulated when use_cost.
 * ---------------------------------------------------------------------------------------- */
int jmo_quant_list(int m, int q_bits, int qp_per, int dequant, int clip, int use_cost, int around, int arw,
                   const int *params, const uint8_t *c_cost, int *coef, int *levels, int *runs, int *fadjust, int *coeff_cost)
{
  int run = 0, n = 0, nonzero = 0;
  for (int k = 0; k < m; k++) {
    const int *q = params + 3 * k;
    if (around && fadjust) fadjust[k] = 0;
    if (coef[k] == 0) { run++; continue; }
    int scaled = iabs_(coef[k]) * q[1];
    int level = (scaled + q[0]) >> q_bits;
    if (level == 0) { coef[k] = 0; run++; continue; }
    if (clip && level > 2063) level = 2063;
    if (around && fadjust) fadjust[k] = (arw * (scaled - (level << q_bits)) + (1 << q_bits)) >> (q_bits + 1);
    if (use_cost) *coeff_cost += (level > 1) ? 999999 : c_cost[run];
    if (coef[k] < 0) level = -level;
    int dq = (level * q[2]) << qp_per;
    coef[k] = dequant == 0 ? level : (dequant == 1 ? dq : ((dq + 8) >> 4));
    levels[n] = level; runs[n] = run; n++; run = 0; nonzero = 1;
  }
  levels[n] = 0;
  return nonzero;
}

/* hadamard4x4 / ihadamard4x4 / hadamard4x2 / ihadamard4x2 / hadamard2x2 / ihadamard2x2, lcommon/src/transform.c:121-330.
 * kind 0..5 in that order; flat layouts as documented in include/jmb200.h (jmb_hadamard). */
void jmo_hadamard(int kind, int *b)
{
  if (kind <= 1) {
    int m[16];
    for (int i = 0; i < 4; i++) {
      int p0 = b[4 * i], p1 = b[4 * i + 1], p2 = b[4 * i + 2], p3 = b[4 * i + 3];
      if (kind == 0) { int t0 = p0 + p3, t1 = p1 + p2, t2 = p1 - p2, t3 = p0 - p3; m[4 * i] = t0 + t1; m[4 * i + 1] = t3 + t2; m[4 * i + 2] = t0 - t1; m[4 * i + 3] = t3 - t2; }
      else { int q0 = p0 + p2, q1 = p0 - p2, q2 = p1 - p3, q3 = p1 + p3; m[4 * i] = q0 + q3; m[4 * i + 1] = q1 + q2; m[4 * i + 2] = q1 - q2; m[4 * i + 3] = q0 - q3; }
    }
    for (int i = 0; i < 4; i++) {
      int p0 = m[i], p1 = m[4 + i], p2 = m[8 + i], p3 = m[12 + i];
      if (kind == 0) { int t0 = p0 + p3, t1 = p1 + p2, t2 = p1 - p2, t3 = p0 - p3; b[i] = (t0 + t1) >> 1; b[4 + i] = (t2 + t3) >> 1; b[8 + i] = (t0 - t1) >> 1; b[12 + i] = (t3 - t2) >> 1; }
      else { int q0 = p0 + p2, q1 = p0 - p2, q2 = p1 - p3, q3 = p1 + p3; b[i] = q0 + q3; b[4 + i] = q1 + q2; b[8 + i] = q1 - q2; b[12 + i] = q0 - q3; }
    }
  } else if (kind <= 3) {
    int m[8], o[8];
    for (int i = 0; i < 4; i++) { m[i] = b[i] + b[4 + i]; m[4 + i] = b[i] - b[4 + i]; }
    for (int i = 0; i < 2; i++) {
      int p0 = m[4 * i], p1 = m[4 * i + 1], p2 = m[4 * i + 2], p3 = m[4 * i + 3];
      if (kind == 2) { int t0 = p0 + p3, t1 = p1 + p2, t2 = p1 - p2, t3 = p0 - p3; o[4 * i] = t0 + t1; o[4 * i + 1] = t3 + t2; o[4 * i + 2] = t0 - t1; o[4 * i + 3] = t3 - t2; }
      else { int t0 = p0 + p2, t1 = p0 - p2, t2 = p1 - p3, t3 = p1 + p3; o[i] = t0 + t3; o[2 + i] = t1 + t2; o[4 + i] = t1 - t2; o[6 + i] = t0 - t3; }
    }
    for (int i = 0; i < 8; i++) b[i] = o[i];
  } else {
    int a = b[0], c = b[1], d = b[2], e = b[3];
    if (kind == 4) { int p0 = a + c, p1 = a - c, p2 = d + e, p3 = d - e; b[0] = p0 + p2; b[1] = p1 + p3; b[2] = p0 - p2; b[3] = p1 - p3; }
    else { int t0 = a + c, t1 = a - c, t2 = d + e, t3 = d - e; b[0] = t0 + t2; b[1] = t1 + t3; b[2] = t0 - t2; b[3] = t1 - t3; }
  }
}

/* ------------------------------------------------------------------------------------------
 * EPZS.  EPZS_integer_motion_estimation (lencod/src/me_epzs_int.c:42-426) and
 * EPZS_sub_pel_motion_estimation (lencod/src/me_epzs_sub.c:30-213), restated for a caller that
 * supplies what JM's host state would: the ordered predictor list in up to four segments, each
 * behind the cost gate its generator sits behind in JM (always / min_mcost > k * stopCriterion:
 * me_epzs_common.c:1556 temporal neighbours k=1, me_epzs_int.c:193-198 window predictors k=3,
 * :211 block-type predictors k=2), the stop criterion (EPZSDetermineStopCriterion,
 * me_epzs_common.c:1874), medthres / subthres (:454-457) and *prevSad.  Refinement patterns:
 * pattern_data, me_epzs_common.c:48-76 (offsets in quarter-pel, next start, next count) with the
 * chaining of EPZSInit (:178-230): sbdiamond and pmvfast hand over to the small diamond.
 * Distortions go through jmo_dist with JM's threshold form (d > thr>>5 ? thr : d<<5,
 * mv_search.h:19-23, me_distortion.c:349-426).
 * ---------------------------------------------------------------------------------------- */
static const short epzs_pat[6][12][4] = {
  {{0, 4, 3, 3}, {4, 0, 0, 3}, {0, -4, 1, 3}, {-4, 0, 2, 3}},
  {{0, 4, 7, 3}, {4, 4, 7, 5}, {4, 0, 1, 3}, {4, -4, 1, 5}, {0, -4, 3, 3}, {-4, -4, 3, 5}, {-4, 0, 5, 3}, {-4, 4, 5, 5}},
  {{-4, 4, 10, 5}, {0, 8, 10, 8}, {0, 4, 10, 7}, {4, 4, 1, 5}, {8, 0, 1, 8}, {4, 0, 1, 7}, {4, -4, 4, 5}, {0, -8, 4, 8},
   {0, -4, 4, 7}, {-4, -4, 7, 5}, {-8, 0, 7, 8}, {-4, 0, 7, 7}},
  {{0, 8, 6, 5}, {4, 4, 0, 3}, {8, 0, 0, 5}, {4, -4, 2, 3}, {0, -8, 2, 5}, {-4, -4, 4, 3}, {-8, 0, 4, 5}, {-4, 4, 6, 3}},
  {{0, 8, 6, 12}, {4, 4, 0, 12}, {8, 0, 0, 12}, {4, -4, 2, 12}, {0, -8, 2, 12}, {-4, -4, 4, 12}, {-8, 0, 4, 12}, {-4, 4, 6, 12},
   {0, 2, 6, 12}, {2, 0, 0, 12}, {0, -2, 2, 12}, {-2, 0, 4, 12}},
  {{0, 8, 6, 5}, {4, 4, 0, 3}, {8, 0, 0, 5}, {4, -4, 2, 3}, {0, -8, 2, 5}, {-4, -4, 4, 3}, {-8, 0, 4, 5}, {-4, 4, 6, 3}}};
static const int epzs_pat_n[6] = {4, 8, 12, 8, 12, 8}, epzs_pat_stop[6] = {1, 1, 1, 1, 0, 0}, epzs_pat_next[6] = {0, 1, 2, 3, 0, 0};
/* nextLast is TRUE for every pattern EPZSInit builds */

static const int bt_sx[8] = {16, 16, 16, 8, 8, 8, 4, 4}, bt_sy[8] = {16, 16, 8, 16, 8, 4, 8, 4};

typedef struct { const jmo_ref *r; const uint16_t *src; int bsx, bsy, px, py; int evals; } epzs_blk;
static int64_t epzs_dist(epzs_blk *b, int metric, int t8, int mvx, int mvy, int64_t thr)
{
  int d = jmo_dist(b->r, b->src, b->bsx, b->bsy, b->px * 4 + mvx, b->py * 4 + mvy, metric, t8);
  b->evals++;
  return (int64_t)d > (thr >> 5) ? thr : ((int64_t)d << 5);
}

void jmo_epzs(const jmo_ref *r, const uint16_t *cur, int cur_stride, const jmo_epzs_req *q, const int16_t *cands,
              const int *me /* metric_h, metric_q, start_hp, start_qp, search_pos2 */, jmo_epzs_res *o)
{
  const int64_t BIG = (int64_t)0x7fffffff << 5;
  uint16_t src[256];
  epzs_blk B = {r, src, bt_sx[q->blocktype], bt_sy[q->blocktype], q->pos_x, q->pos_y, 0};
  get_block(cur, cur_stride, q->pos_x, q->pos_y, B.bsx, B.bsy, src);
  const int sx = q->start_x, sy = q->start_y, px = q->pred_x, py = q->pred_y, rx = q->range_x, ry = q->range_y;
  int tx = sx, ty = sy;                       /* tmp */
  int64_t min_mcost = q->min_mcost, prev = q->prev_sad;
  int exit_code = 0;
  const int gt0 = (q->flags & JMO_EPZS_REF_GT0_FRAME) != 0;

  if (!(q->flags & JMO_EPZS_SKIP_INT)) {
    const int lam = q->lambda[0];
    const int64_t ld = 2 * (int64_t)lam, med = q->medthres, stop = q->stop;
    const int mw = 2 * rx + 1;
    uint8_t *map = (uint8_t *)calloc((size_t)mw * (2 * ry + 1), 1);
#define VIS(x, y) map[((y) - sy + ry) * mw + ((x) - sx + rx)]
#define INR(x, y) (iabs_((x) - sx) - rx <= 0 && iabs_((y) - sy) - ry <= 0)
    VIS(sx, sy) = 1;
    min_mcost = mvcost(lam, sx, sy, px, py);
    min_mcost += epzs_dist(&B, JMO_SAD, 0, sx, sy, BIG - min_mcost);
    if (gt0 && (prev < (med + ld < min_mcost ? med + ld : min_mcost) || prev * 8 < min_mcost)) exit_code = 1;   /* :103-117 */
    else if (min_mcost > med + ld) {                                                                    /* :121 */
      if (min_mcost < (stop >> 1)) {                                                                    /* :135-150 */
        if (q->jm_ref == 0 || prev > min_mcost) prev = min_mcost;
        exit_code = 2;
      } else {
        int64_t second = BIG, mcost;
        const int64_t centre_cost = min_mcost;      /* JM runs the generators (and their gates) before it checks any predictor */
        int check_median = 0, t2x = 0, t2y = 0;
        const int16_t *c = cands + 2 * (size_t)q->cand_off;
        for (int s = 0; s < 4; s++) {                                                                   /* predictor generators, :152-212 */
          const int on = q->gate[s] == 0 || centre_cost > (int64_t)q->gate[s] * stop;
          const int gen = s == 2 && (q->flags & JMO_EPZS_WINDOW_GEN);       /* EPZSWindowPredictorInit mode 0, me_epzs_common.c:352-371 */
          static const signed char ring[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};
          for (int i = 0; i < q->n_cand[s]; i++) {
            int vx, vy;
            if (gen) { const int rings = (q->n_cand[s] + 8) >> 3, sp = rx >> (rings - 1 - (i >> 3)); vx = sx + ring[i & 7][0] * sp; vy = sy + ring[i & 7][1] * sp; }
            else { vx = c[0]; vy = c[1]; c += 2; }                                                      /* :215-252 */
            if (!on) continue;
            if (!INR(vx, vy) || VIS(vx, vy)) continue;
            VIS(vx, vy) = 1;
            mcost = mvcost(lam, vx, vy, px, py);
            if (mcost < second) {
              mcost += epzs_dist(&B, JMO_SAD, 0, vx, vy, second - mcost);
              if (mcost < min_mcost) { t2x = tx; t2y = ty; tx = vx; ty = vy; second = min_mcost; min_mcost = mcost; check_median = 1; }
              else if (mcost < second) { t2x = vx; t2y = vy; second = mcost; check_median = 1; }
            }
          }
        }
        if (gt0 && prev * 3 < min_mcost) exit_code = 3;                                                 /* :254-273 */
        else if (min_mcost > stop) {                                                                    /* :279 */
          int pat = q->pattern, cx, cy;
          if (q->flags & JMO_EPZS_ADAPT_PATTERN) {                                                      /* :286-300 */
            if (min_mcost < stop + ((3 * med) >> 1))
              pat = ((tx == 0 && ty == 0) || (iabs_(tx - sx) < 10 && iabs_(ty - sy) < 10)) ? 0 : 1;
            else if (q->flags & JMO_EPZS_SQUARE_HINT) pat = 1;
          }
          cx = tx; cy = ty;
          for (;;) {
            int pattern_stop = 0, point = 0, next_last = 0, total = epzs_pat_n[pat], dir = 0;
            do {                                                                                        /* :307-360 */
              int check = total;
              do {
                const int vx = cx + epzs_pat[pat][point][0], vy = cy + epzs_pat[pat][point][1];
                if (INR(vx, vy) && !VIS(vx, vy)) {
                  VIS(vx, vy) = 1;
                  mcost = mvcost(lam, vx, vy, px, py);
                  if (mcost < min_mcost) {
                    mcost += epzs_dist(&B, JMO_SAD, 0, vx, vy, min_mcost - mcost);
                    if (mcost < min_mcost) { tx = vx; ty = vy; min_mcost = mcost; dir = point; }
                  }
                }
                if (++point >= epzs_pat_n[pat]) point -= epzs_pat_n[pat];
              } while (--check > 0);
              if (next_last || (tx == cx && ty == cy)) {
                pattern_stop = epzs_pat_stop[pat];
                pat = epzs_pat_next[pat];
                total = epzs_pat_n[pat];
                next_last = 1;
                dir = 0; point = 0;
              } else {
                total = epzs_pat[pat][dir][3];
                point = epzs_pat[pat][dir][2];
                cx = tx; cy = ty;
              }
            } while (pattern_stop != 1);
            if (gt0 && (4 * prev < min_mcost || (3 * prev < min_mcost && prev <= stop))) { exit_code = 4; break; }   /* :362-376 */
            if (!(check_median && (q->jm_ref == 0 || min_mcost < 2 * prev) && min_mcost > ((3 * stop) >> 1) && (q->flags & JMO_EPZS_DUAL))) break;   /* :379-384 */
            if ((tx == 0 && ty == 0) || (tx == sx && ty == sy)) pat = (iabs_(tx - sx) < 10 && iabs_(ty - sy) < 10) ? 0 : 1;   /* :391-399 */
            else pat = q->pattern_dual;
            cx = t2x; cy = t2y;
            check_median = 0;
          }
        }
      }
    }
    if (!exit_code) { if (q->jm_ref == 0 || prev > min_mcost) prev = min_mcost; exit_code = 5; }           /* :409-410 */
    free(map);
#undef VIS
#undef INR
  }
  o->imv_x = (int16_t)tx; o->imv_y = (int16_t)ty; o->icost = min_mcost; o->prev_sad = prev; o->exit_code = exit_code;
  int mvx = tx, mvy = ty;

  /* BlockMotionSearch between the two stages (mv_search.c:964-976): sub-pel only for ref 0, or when the integer cost is below
   * 3.5 x the (updated) previous distortion; DISTBLK_MAX on entry unless start_me_refinement_hp */
  if ((q->flags & JMO_EPZS_SUBPEL) && ((q->flags & JMO_EPZS_SKIP_INT) || !gt0 || 2 * min_mcost < 7 * prev)) {
    static const signed char hp[10][2] = {{0, 0}, {-2, 0}, {0, 2}, {2, 0}, {0, -2}, {-2, 2}, {2, 2}, {2, -2}, {-2, -2}, {-2, 2}};
    static const int ns[5][5] = {{0, 8, 5, 6, 7}, {8, 0, 5, 8, 8}, {5, 5, 0, 6, 5}, {6, 6, 6, 0, 7}, {7, 8, 7, 7, 0}};
    static const int ne[5][5] = {{0, 10, 7, 8, 9}, {10, 0, 6, 10, 9}, {7, 6, 0, 7, 7}, {8, 8, 7, 0, 8}, {9, 9, 9, 8, 0}};
    const int metric_h = me[0], metric_q = me[1], start_hp = me[2], start_qp = me[3], search_pos2 = me[4];
    const int t8 = (q->flags & JMO_EPZS_TEST8X8) != 0;
    const int max_pos2 = (!start_hp || !start_qp) ? (search_pos2 > 1 ? search_pos2 : 1) : search_pos2;
    int lam = q->lambda[1], pos, best = 0, second_pos = 0, done = 0;
    int64_t second = BIG, mcost, sub_thr = q->subthres + 2 * (int64_t)lam;
    if (!(q->flags & JMO_EPZS_SKIP_INT) && !start_hp) min_mcost = BIG;
    for (pos = start_hp; pos < (5 < max_pos2 ? 5 : max_pos2); pos++) {                                   /* me_epzs_sub.c:66-90 */
      const int vx = mvx + hp[pos][0], vy = mvy + hp[pos][1];
      mcost = mvcost(lam, vx, vy, px, py);
      if (mcost < second) {
        mcost += epzs_dist(&B, metric_h, t8, vx, vy, second - mcost);
        if (mcost < min_mcost) { second = min_mcost; second_pos = best; min_mcost = mcost; best = pos; }
        else if (mcost < second) { second = mcost; second_pos = pos; }
      }
    }
    if (best == 0 && px == mvx && py == mvy && min_mcost < sub_thr) done = 1;                            /* :92-95 */
    if (!done) {
      if (search_pos2 >= 9 && (best != 0 || (iabs_(px - mvx) + iabs_(py - mvy)))) {                      /* :97-122 */
        const int p0 = ns[best][second_pos], p1 = ne[best][second_pos];
        for (pos = p0; pos < p1; pos++) {
          const int vx = mvx + hp[pos][0], vy = mvy + hp[pos][1];
          mcost = mvcost(lam, vx, vy, px, py);
          if (mcost < min_mcost) {
            mcost += epzs_dist(&B, metric_h, t8, vx, vy, min_mcost - mcost);
            if (mcost < min_mcost) { min_mcost = mcost; best = pos; }
          }
        }
      }
      if (best) { mvx += hp[best][0]; mvy += hp[best][1]; }
      const int end_pos = (min_mcost < sub_thr) ? 1 : 5;                                                 /* :135-170 */
      second = BIG;
      if (!start_qp) { best = -1; min_mcost = BIG; } else best = 0;
      lam = q->lambda[2];
      for (pos = start_qp; pos < end_pos; pos++) {
        const int vx = mvx + hp[pos][0] / 2, vy = mvy + hp[pos][1] / 2;
        mcost = mvcost(lam, vx, vy, px, py);
        if (mcost < second) {
          mcost += epzs_dist(&B, metric_q, t8, vx, vy, second - mcost);
          if (mcost < min_mcost) { second = min_mcost; second_pos = best; min_mcost = mcost; best = pos; }
          else if (mcost < second) { second = mcost; second_pos = pos; }
        }
      }
      if (min_mcost > sub_thr && (best != 0 || (iabs_(px - mvx) + iabs_(py - mvy)))) {                   /* :173-200 */
        const int p0 = ns[best][second_pos], p1 = ne[best][second_pos];
        for (pos = p0; pos < p1; pos++) {
          const int vx = mvx + hp[pos][0] / 2, vy = mvy + hp[pos][1] / 2;
          mcost = mvcost(lam, vx, vy, px, py);
          if (mcost < min_mcost) {
            mcost += epzs_dist(&B, metric_q, t8, vx, vy, min_mcost - mcost);
            if (mcost < min_mcost) { min_mcost = mcost; best = pos; }
          }
        }
      }
      if (best > 0) { mvx += hp[best][0] / 2; mvy += hp[best][1] / 2; }
    }
  }
  o->mv_x = (int16_t)mvx; o->mv_y = (int16_t)mvy; o->cost = min_mcost; o->n_evals = B.evals;
}

/* ------------------------------------------------------------------------------------------
 * Batch forms used by bench.py's CPU legs (kind "port") and by the picture-sized parity checks.
 * ---------------------------------------------------------------------------------------- */
void jmo_epzs_batch(const jmo_ref *r, const uint16_t *cur, int cur_stride, const jmo_epzs_req *reqs, int n, const int16_t *cands,
                    const int *me, jmo_epzs_res *res)
{
  for (int i = 0; i < n; i++) jmo_epzs(r, cur, cur_stride, &reqs[i], cands, me, &res[i]);
}

/* prediction -> residual -> forward transform -> quantisation of the partition modes in mode_mask for one macroblock, from the
 * 41 motion vectors of its searches: luma_prediction (mc_prediction.c:117-236, one origin clamp per prediction unit:
 * macroblock.c:946-971, :1225) -> forward4x4 / forward8x8 -> quant_4x4_normal / quant_8x8_normal / quant_8x8cavlc_normal
 * (variants 0 / 2 / 4 of jmo_quant).  levels: [7][256] dense in scan order, layout of jmb_mc_tq_modes. */
void jmo_mc_tq_modes_mb(const jmo_ref *r, const uint16_t *cur, int cur_stride, int mb_x, int mb_y, const int16_t *mv41, int n, int qp,
                        const int *qparams, const uint8_t *scan, const uint8_t *c_cost, int is_cavlc, unsigned mode_mask, int16_t *levels)
{
  static const int base[8] = {0, 0, 1, 3, 5, 9, 17, 25}, w4[8] = {4, 4, 4, 2, 2, 2, 1, 1}, h4[8] = {4, 4, 2, 4, 2, 1, 2, 1};
  const int per_mb = (n == 4) ? 16 : 4, nn = n * n, cavlc8 = (n == 8 && is_cavlc);
  memset(levels, 0, 7 * 256 * sizeof(int16_t));
  for (int mode = 1; mode <= 7; mode++) {
    if (!((mode_mask >> (mode - 1)) & 1)) continue;
    for (int b = 0; b < per_mb; b++) {
      const int bx4 = (n == 4) ? (b & 3) : (b & 1) * 2, by4 = (n == 4) ? (b >> 2) : (b >> 1) * 2;
      int ux4 = bx4, uy4 = by4;
      if (mode < 5 || n == 8) { ux4 &= ~1; uy4 &= ~1; }
      if (mode == 1) { ux4 = 0; uy4 = 0; }
      const int slot = base[mode] + (uy4 / h4[mode]) * (4 / w4[mode]) + ux4 / w4[mode];
      const int qx = ((mb_x + ux4 * 4) << 2) + mv41[2 * slot], qy = ((mb_y + uy4 * 4) << 2) + mv41[2 * slot + 1];
      const uint16_t *rl = umv_line(r, qy, qx) + (by4 - uy4) * 4 * r->W + (bx4 - ux4) * 4;
      int blk[64], lv[68], rn[68], fa[64], cost = 0;
      for (int y = 0; y < n; y++)
        for (int x = 0; x < n; x++)
          blk[y * n + x] = (int)cur[(size_t)(mb_y + by4 * 4 + y) * cur_stride + mb_x + bx4 * 4 + x] - (int)rl[y * r->W + x];
      if (n == 4) jmo_forward4x4(blk); else jmo_forward8x8(blk);
      jmo_quant(n == 4 ? 0 : (cavlc8 ? 4 : 2), blk, qp, qparams, scan, c_cost, is_cavlc, 0, lv, rn, fa, &cost);
      int16_t *o = levels + (mode - 1) * 256 + b * nn;
      if (cavlc8) {
        for (int s = 0; s < 4; s++)
          for (int i = 0, k = 0; lv[17 * s + i]; i++) { k += rn[17 * s + i]; o[16 * s + k++] = (int16_t)lv[17 * s + i]; }
      } else
        for (int i = 0, k = 0; lv[i]; i++) { k += rn[i]; o[k++] = (int16_t)lv[i]; }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Chroma of an inter macroblock: motion-compensated prediction and residual coding.
 *  - prediction: OneComponentChromaPrediction4x4_regenerate (lencod/src/mc_prediction.c:292-352): every chroma sample takes the
 *    motion vector of the luma 4x4 block it lies under; bilinear interpolation at 1/8 (1/4 vertically for 4:2:2) sample
 *    precision, source coordinates clamped to the picture; C integer division (toward zero) as in JM.
 *  - residual coding: residual_transform_quant_chroma_4x4 (lencod/src/block.c:954-1202) for intra = 0, frame scan, no adaptive
 *    rounding: forward4x4 of the 4 (4:2:0) or 8 (4:2:2) blocks, DC through hadamard2x2 + quant_dc2x2_normal or hadamard4x2 +
 *    quant_dc4x2_normal at qp + 3 (SCAN_YUV420 / SCAN_YUV422, block.c:78-94), quant_ac4x4_normal per block with one running
 *    coeff_cost per component, the _CHROMA_COEFF_COST_ (4) threshold, cbp_blk bits (cbp_blk_chroma, block.c:158), inverse4x4 and
 *    sample_reconstruct.  The DC / AC quantisers and Hadamards are the pinned list quantiser / jmo_hadamard of this file.
 * ---------------------------------------------------------------------------------------- */
void jmo_chroma_pred(const uint8_t *ref_c, int wc, int hc, int stride, int yuv, int mb_cx, int mb_cy, const int16_t *mv16 /* [16][2] luma 4x4 mvs */,
                     uint8_t *pred /* [hc_mb][8] */)
{
  const int hmb = (yuv == 1) ? 8 : 16, f1x = 8, f1y = 64 / hmb, f2x = f1x - 1, f2y = f1y - 1, f3 = f1x * f1y, f4 = f3 >> 1;
  const int ydiv = hmb >> 2;
  for (int j = 0; j < hmb; j++)
    for (int i = 0; i < 8; i++) {
      const int16_t *mv = mv16 + 2 * ((j / ydiv) * 4 + i / 2);
      const int ii = (i + mb_cx) * f1x + mv[0], jj = (j + mb_cy) * f1y + mv[1];
      const int ii0 = iclip(0, wc - 1, ii / f1x), jj0 = iclip(0, hc - 1, jj / f1y);
      const int ii1 = iclip(0, wc - 1, (ii + f2x) / f1x), jj1 = iclip(0, hc - 1, (jj + f2y) / f1y);
      const int if1 = ii & f2x, if0 = f1x - if1, jf1 = jj & f2y, jf0 = f1y - jf1;
      pred[j * 8 + i] = (uint8_t)((if0 * jf0 * ref_c[jj0 * stride + ii0] + if1 * jf0 * ref_c[jj0 * stride + ii1] +
                                   if0 * jf1 * ref_c[jj1 * stride + ii0] + if1 * jf1 * ref_c[jj1 * stride + ii1] + f4) / f3);
    }
}

/* one component of one macroblock.  src / pred [hmb][8]; qp_ac / qp_dc scaled chroma qps; params_ac [16][3] (row-major j*4+i),
 * params_dc [3]; out: dc_levels[8] and ac_levels[nb][15] dense in scan order, *cbp_bits = the component's bits of cbp_blk
 * shifted down by 16 + uv * nb (bit b = block b raster, 2 blocks per row), recon [hmb][8]; returns cr_cbp (0, 1, 2). */
int jmo_chroma_rc(const uint8_t *src, const uint8_t *pred, int yuv, int qp_ac, int qp_dc, const int *params_ac, const int *params_dc,
                  const uint8_t *c_cost, int is_cavlc, int16_t *dc_levels, int16_t *ac_levels, unsigned *cbp_bits, uint8_t *recon)
{
  static const uint8_t scan4[16][2] = {{0,0},{1,0},{0,1},{0,2},{1,1},{2,0},{3,0},{2,1},{1,2},{0,3},{1,3},{2,2},{3,1},{3,2},{2,3},{3,3}};
  static const uint8_t scan422[8][2] = {{0,0},{0,1},{1,0},{0,2},{0,3},{1,1},{1,2},{1,3}};
  const int hmb = (yuv == 1) ? 8 : 16, nb = hmb / 2;
  int rres[16][8], blk[16], lv[17], rn[17], fa[16], cost = 0, dczero, nonzero[8] = {0}, any_ac = 0, cr_cbp = 0, cr_tmp = 0;
  memset(dc_levels, 0, 8 * sizeof(int16_t)); memset(ac_levels, 0, (size_t)nb * 15 * sizeof(int16_t));
  *cbp_bits = 0;
  for (int b = 0; b < nb; b++) {      /* integer transform (block.c:1011-1027) */
    const int n1 = (b & 1) * 4, n2 = (b >> 1) * 4;
    for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) blk[y * 4 + x] = (int)src[(n2 + y) * 8 + n1 + x] - (int)pred[(n2 + y) * 8 + n1 + x];
    jmo_forward4x4(blk);
    for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) rres[n2 + y][n1 + x] = blk[y * 4 + x];
  }
  if (yuv == 1) {                     /* CHROMA DC YUV420 (:1029-1054) */
    int m1[4] = {rres[0][0], rres[0][4], rres[4][0], rres[4][4]}, pdc[4 * 3];
    jmo_hadamard(4, m1);
    for (int k = 0; k < 4; k++) { pdc[3 * k] = params_dc[0] << 1; pdc[3 * k + 1] = params_dc[1]; pdc[3 * k + 2] = params_dc[2]; }
    dczero = jmo_quant_list(4, 15 + qp_dc / 6 + 1, qp_dc / 6, 1, is_cavlc, 0, 0, 0, pdc, c_cost, m1, lv, rn, fa, NULL);
    for (int i = 0, k = 0; lv[i]; i++) { k += rn[i]; dc_levels[k++] = (int16_t)lv[i]; }
    jmo_hadamard(5, m1);
    rres[0][0] = m1[0] >> 5; rres[0][4] = m1[1] >> 5; rres[4][0] = m1[2] >> 5; rres[4][4] = m1[3] >> 5;
  } else {                            /* CHROMA DC YUV422 (:1055-1092): tblk[x][y] = DC of block (x, y), transposed */
    int t[8], list[8], pdc[8 * 3], o[8];
    for (int x = 0; x < 2; x++) for (int y = 0; y < 4; y++) t[x * 4 + y] = rres[y * 4][x * 4];
    jmo_hadamard(2, t);
    for (int k = 0; k < 8; k++) { list[k] = t[scan422[k][0] * 4 + scan422[k][1]]; pdc[3 * k] = params_dc[0] << 1; pdc[3 * k + 1] = params_dc[1]; pdc[3 * k + 2] = params_dc[2]; }
    dczero = jmo_quant_list(8, 15 + qp_dc / 6 + 1, qp_dc / 6, 1, is_cavlc, 0, 0, 0, pdc, c_cost, list, lv, rn, fa, NULL);
    for (int i = 0, k = 0; lv[i]; i++) { k += rn[i]; dc_levels[k++] = (int16_t)lv[i]; }
    for (int k = 0; k < 8; k++) t[scan422[k][0] * 4 + scan422[k][1]] = list[k];
    jmo_hadamard(3, t);               /* out: 4 rows x 2 */
    for (int k = 0; k < 8; k++) o[k] = t[k];
    for (int j = 0; j < 4; j++) {
      rres[j << 2][0] = (o[2 * j] + 32) >> 6;
      rres[j << 2][4] = (o[2 * j + 1] + 32) >> 6;
    }
  }
  if (dczero) { *cbp_bits = (1u << nb) - 1; cr_cbp = 1; }
  for (int b = 0; b < nb; b++) {      /* chroma AC (:1094-1132) */
    const int n1 = (b & 1) * 4, n2 = (b >> 1) * 4;
    int list[15], pac[15 * 3];
    for (int k = 1; k < 16; k++) {
      const int i = scan4[k][0], j = scan4[k][1];
      list[k - 1] = rres[n2 + j][n1 + i];
      pac[3 * (k - 1)] = params_ac[3 * (j * 4 + i)]; pac[3 * (k - 1) + 1] = params_ac[3 * (j * 4 + i) + 1]; pac[3 * (k - 1) + 2] = params_ac[3 * (j * 4 + i) + 2];
    }
    nonzero[b] = jmo_quant_list(15, 15 + qp_ac / 6, qp_ac / 6, 2, is_cavlc, 1, 0, 0, pac, c_cost, list, lv, rn, fa, &cost);
    for (int k = 1; k < 16; k++) rres[n2 + scan4[k][1]][n1 + scan4[k][0]] = list[k - 1];
    for (int i = 0, k = 0; lv[i]; i++) { k += rn[i]; ac_levels[b * 15 + k++] = (int16_t)lv[i]; }
    if (nonzero[b]) { *cbp_bits |= 1u << b; cr_tmp = 2; any_ac = 1; }
  }
  if (any_ac && cost < 4) {           /* _CHROMA_COEFF_COST_ (:1134-1170) */
    cr_tmp = 0;
    for (int b = 0; b < nb; b++)
      if (nonzero[b]) {
        const int n1 = (b & 1) * 4, n2 = (b >> 1) * 4;
        nonzero[b] = 0;
        if (!dczero) *cbp_bits = 0;
        for (int k = 1; k < 16; k++) rres[n2 + scan4[k][1]][n1 + scan4[k][0]] = 0;
        memset(ac_levels + b * 15, 0, 15 * sizeof(int16_t));
      }
  }
  if (cr_tmp == 2) cr_cbp = 2;
  int any = 0;
  for (int b = 0; b < nb; b++) {      /* inverse transform + reconstruction (:1176-1199) */
    const int n1 = (b & 1) * 4, n2 = (b >> 1) * 4;
    if (rres[n2][n1] != 0 || nonzero[b]) {
      for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) blk[y * 4 + x] = rres[n2 + y][n1 + x];
      jmo_inverse4x4(blk);
      for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) rres[n2 + y][n1 + x] = blk[y * 4 + x];
      any = 1;
    }
  }
  for (int y = 0; y < hmb; y++)
    for (int x = 0; x < 8; x++)
      recon[y * 8 + x] = any ? (uint8_t)iclip(0, 255, ((rres[y][x] + 32) >> 6) + pred[y * 8 + x]) : pred[y * 8 + x];
  return cr_cbp;
}

/* ---- deblocking: DeblockFrame (lencod/src/loopFilter.c:63-299) + loop_filter_normal.c, non-MBAFF frame pictures, 8 bit -------
 * Sequential restatement, macroblock after macroblock in raster order; pinned against JM's own DeblockFrame by
 * tests/test_oracle_vs_ref.py::test_deblock_matches_jm (oracle/ref_harness.c::jmref_deblock). */
static const unsigned char DB_ALPHA[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,4,4,5,6,7,8,9,10,12,13,15,17,20,22,25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255};
static const unsigned char DB_BETA[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18};
static const unsigned char DB_CLIP[52][3] = {      /* CLIP_TAB[indexA][1..3] (loop_filter.h:36-45); [4] equals [3] and is never reached with strength < 4 */
  {0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},
  {0,0,1},{0,0,1},{0,0,1},{0,0,1},{0,1,1},{0,1,1},{1,1,1},{1,1,1},{1,1,1},{1,1,1},{1,1,2},{1,1,2},{1,1,2},{1,1,2},{1,2,3},{1,2,3},{2,2,3},
  {2,2,4},{2,3,4},{2,3,4},{3,3,5},{3,4,6},{3,4,6},{4,5,7},{4,5,8},{4,6,9},{5,7,10},{6,8,11},{6,8,13},{7,10,14},{8,11,16},{9,12,18},{10,13,20},
  {11,15,23},{13,17,25}};

static int db_intra(int t) { return t == 9 || t == 10 || t == 13 || t == 14; }
static int db_mvdiff(const int16_t *a, const int16_t *b) { return (abs(a[0] - b[0]) >= 4) | (abs(a[1] - b[1]) >= 4); }      /* compare_mvs, mvlimit 4 (frame) */

/* strength of the 4-sample segment k of edge `edge` (0..3) in direction dir (0 vertical edge, 1 horizontal); Q = this macroblock,
 * P = the macroblock on the other side (Q itself for inner edges): GetStrengthVer / GetStrengthHor */
static int db_strength(int dir, int edge, int k, const jmo_db_mb *Q, const jmo_db_mb *P)
{
  if (db_intra(Q->mb_type) || db_intra(P->mb_type)) return edge == 0 ? 4 : 3;
  int bq = dir ? edge * 4 + k : k * 4 + edge;                                       /* 4x4 block of Q at the edge */
  int bp = edge ? (dir ? bq - 4 : bq - 1) : (dir ? 12 + k : k * 4 + 3);             /* the block across it */
  if (((Q->cbp_blk >> bq) & 1) || ((P->cbp_blk >> bp) & 1)) return 2;
  if (edge && (Q->mb_type == 1 || Q->mb_type == (dir ? 3 : 2))) return 0;
  int p0 = Q->ref_id[0][bq], p1 = Q->ref_id[1][bq], q0 = P->ref_id[0][bp], q1 = P->ref_id[1][bp];      /* (JM names the Q-side block "p": symmetric) */
  if (!((p0 == q0 && p1 == q1) || (p0 == q1 && p1 == q0))) return 1;
  const int16_t *mp0 = Q->mv[0][bq], *mp1 = Q->mv[1][bq], *mq0 = P->mv[0][bp], *mq1 = P->mv[1][bp];
  if (p0 != p1) return p0 == q0 ? (db_mvdiff(mp0, mq0) | db_mvdiff(mp1, mq1)) : (db_mvdiff(mp0, mq1) | db_mvdiff(mp1, mq0));
  return (db_mvdiff(mp0, mq0) | db_mvdiff(mp1, mq1)) && (db_mvdiff(mp0, mq1) | db_mvdiff(mp1, mq0));
}

static int db_clip(int lo, int hi, int v) { return v < lo ? lo : v > hi ? hi : v; }

/* one line of samples across a luma edge, in the standard's notation (8.7.2.3 / 8.7.2.4; loop_filter_normal.c:310-575):
 * q points at q0, `st` steps away from the edge, p0 = q[-st] */
static void db_luma_line(uint8_t *q, int st, int bS, int alpha, int beta, int tc0)
{
  int p0 = q[-st], q0 = q[0], p1 = q[-2 * st], q1 = q[st];
  if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
  int p2 = q[-3 * st], q2 = q[2 * st];
  int ap = abs(p2 - p0) < beta, aq = abs(q2 - q0) < beta;
  if (bS == 4) {
    int strong = abs(p0 - q0) < (alpha >> 2) + 2;
    if (ap && strong) {
      int p3 = q[-4 * st];
      q[-st] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      q[-2 * st] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
      q[-3 * st] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
    } else q[-st] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
    if (aq && strong) {
      int q3 = q[3 * st];
      q[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      q[st] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
      q[2 * st] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
    } else q[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
  } else {
    int tc = tc0 + ap + aq, avg = (p0 + q0 + 1) >> 1;
    int delta = db_clip(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
    if (ap) q[-2 * st] = (uint8_t)(p1 + db_clip(-tc0, tc0, (p2 + avg - 2 * p1) >> 1));
    if (aq) q[st] = (uint8_t)(q1 + db_clip(-tc0, tc0, (q2 + avg - 2 * q1) >> 1));
    if (delta) { q[-st] = (uint8_t)db_clip(0, 255, p0 + delta); q[0] = (uint8_t)db_clip(0, 255, q0 - delta); }
  }
}

/* chroma: only p0 and q0 change (loop_filter_normal.c:585-758) */
static void db_chroma_line(uint8_t *q, int st, int bS, int alpha, int beta, int tc0)
{
  int p0 = q[-st], q0 = q[0], p1 = q[-2 * st], q1 = q[st];
  if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
  if (bS == 4) { q[-st] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2); q[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2); }
  else {
    int tc = tc0 + 1, delta = db_clip(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
    if (delta) { q[-st] = (uint8_t)db_clip(0, 255, p0 + delta); q[0] = (uint8_t)db_clip(0, 255, q0 - delta); }
  }
}

void jmo_deblock(uint8_t *luma, int pitch, uint8_t *cb, uint8_t *cr, int pitch_c, int w, int h, int yuv, int slice_type, int direct8x8inf,
                 const jmo_db_mb *mbs)
{
  /* chroma_edge[dir][edge][yuv_format] and pelnum_cr (loop_filter.h:47-58), for 4:2:0 (1) and 4:2:2 (2) */
  static const int cedge[2][4][3] = {{{-4, 0, 0}, {-4, -4, -4}, {-4, 4, 4}, {-4, -4, -4}}, {{-4, 0, 0}, {-4, -4, 4}, {-4, 4, 8}, {-4, -4, 12}}};
  static const int pelnum[2][3] = {{0, 8, 16}, {0, 8, 8}};
  const int mbw = w / 16, mbh = h / 16;
  for (int my = 0; my < mbh; my++)
    for (int mx = 0; mx < mbw; mx++) {
      const jmo_db_mb *Q = &mbs[my * mbw + mx];
      if (Q->df_disable_idc == 1) continue;
      const int t8 = Q->flags & 1, cbp = Q->flags & 2;
      int edge0[2] = {mx != 0, my != 0};
      if (Q->df_disable_idc == 2) { edge0[0] = (Q->flags & 4) != 0; edge0[1] = (Q->flags & 8) != 0; }
      for (int dir = 0; dir < 2; dir++)
        for (int edge = 0; edge < 4; edge++) {
          const int luma_on = !(t8 && (edge & 1));
          if (!cbp) {      /* loopFilter.c:153-164, :209-220 */
            if (!luma_on && (dir == 0 || yuv == 1)) continue;
            if (edge > 0 && (slice_type == 0 || slice_type == 1)) {
              if ((Q->mb_type == 0 && slice_type == 0) || Q->mb_type == 1 || Q->mb_type == (dir ? 3 : 2)) continue;
              if ((edge & 1) && (Q->mb_type == (dir ? 2 : 3) || (Q->mb_type == 0 && slice_type == 1 && direct8x8inf))) continue;
            }
          }
          if (!(edge || edge0[dir])) continue;
          const jmo_db_mb *P = edge ? Q : (dir ? Q - mbw : Q - 1);
          int bs[4], any = 0;
          for (int k = 0; k < 4; k++) { bs[k] = db_strength(dir, edge, k, Q, P); any |= bs[k]; }
          if (!any) continue;
          if (luma_on) {
            const int qp = (P->qp + Q->qp + 1) >> 1, ia = db_clip(0, 51, qp + Q->df_alpha_c0_offset), ib = db_clip(0, 51, qp + Q->df_beta_offset);
            const int alpha = DB_ALPHA[ia], beta = DB_BETA[ib];
            if (alpha | beta)
              for (int i = 0; i < 16; i++) {
                const int s = bs[i >> 2];
                if (!s) continue;
                uint8_t *q = dir ? luma + (size_t)(my * 16 + edge * 4) * pitch + mx * 16 + i : luma + (size_t)(my * 16 + i) * pitch + mx * 16 + edge * 4;
                db_luma_line(q, dir ? pitch : 1, s, alpha, beta, s < 4 ? DB_CLIP[ia][s - 1] : 0);
              }
          }
          if (yuv == 1 || yuv == 2) {
            const int ec = cedge[dir][edge][yuv];
            if (ec < 0) continue;
            const int n = pelnum[dir][yuv], cw = 8, chh = yuv == 1 ? 8 : 16;
            for (int uv = 0; uv < 2; uv++) {
              uint8_t *pl = uv ? cr : cb;
              const int qp = (P->qpc[uv] + Q->qpc[uv] + 1) >> 1, ia = db_clip(0, 51, qp + Q->df_alpha_c0_offset), ib = db_clip(0, 51, qp + Q->df_beta_offset);
              const int alpha = DB_ALPHA[ia], beta = DB_BETA[ib];
              if (!(alpha | beta)) continue;
              for (int i = 0; i < n; i++) {
                const int s = bs[n == 8 ? i >> 1 : i >> 2];      /* Strength[(PelNum == 8) ? ((pel >> 1) << 2) + (pel & 1) : pel] */
                if (!s) continue;
                uint8_t *q = dir ? pl + (size_t)(my * chh + ec) * pitch_c + mx * cw + i : pl + (size_t)(my * chh + i) * pitch_c + mx * cw + ec;
                db_chroma_line(q, dir ? pitch_c : 1, s, alpha, beta, s < 4 ? DB_CLIP[ia][s - 1] : 0);
              }
            }
          }
        }
    }
}
