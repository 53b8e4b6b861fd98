// k_search.cu -- K1 + K2 + K3: integer-pel full search for every partition of a macroblock.
//
// One CTA per (macroblock, reference) "group".  The CTA stages the reference search window and the
// 16x16 source block in shared memory, evaluates the sixteen 4x4 SADs of every displacement with
// packed-byte VABSDIFF4 (4 samples per instruction), sums them to the 41 partition SADs exactly as
// update_full_search_large_blocks does (lencod/src/me_fullfast.c:196-260) and keeps, per
// partition, the minimum of  J = (SAD << 5) + lambda * (mvbits[dx] + mvbits[dy])  with JM's
// tie-break: the first position in spiral order wins (strict '<' at me_fullsearch.c:89 and
// me_fullfast.c:677).  The 4x4 SADs are shared by all partitions of the macroblock; nothing but
// the 41 (mv, cost) results leaves the SM.
//
// Semantics reproduced:
//  * FULL search  (full_search_motion_estimation, lencod/src/me_fullsearch.c:39-103): each
//    partition has its own centre; computeSAD clamps the PARTITION origin (me_distortion.c:367).
//  * FAST FULL search (setup_fast_full_search + fast_full_search_motion_estimation,
//    lencod/src/me_fullfast.c:269-689): one centre per macroblock; the MACROBLOCK origin is clamped
//    (:498); candidates with GetMaxMVD >= max_mvd-1 are skipped (:671).
//  Both clamps are "displacement clamps": the block(s) are read at D = clamp(d, Dlo, Dhi) per axis,
//  so a thread that owns displacement D serves every candidate d that clamps onto it.
//  JM's early terminations only replace a losing cost by min_mcost, so complete sums decide alike.
#include "jmb_internal.h"

namespace {

constexpr int NPART = 41;
constexpr int CW = 128;            // chunk of displacements handled per staging pass
constexpr int CH = 72;
constexpr int WIN_PITCH = CW + 16 + 4;
constexpr int WIN_ROWS = CH + 15;
constexpr int IDX_BITS = 13;       // (2*64+1)^2 = 16641 > 8192: search_range <= 45 keeps idx < 8192

struct PartGeom { unsigned char type, bx, by, w4, h4; };
// canonical partition order: by type, then raster order of the partitions inside the macroblock
__constant__ PartGeom c_part[NPART] = {
  {1,0,0,4,4},
  {2,0,0,4,2},{2,0,2,4,2},
  {3,0,0,2,4},{3,2,0,2,4},
  {4,0,0,2,2},{4,2,0,2,2},{4,0,2,2,2},{4,2,2,2,2},
  {5,0,0,2,1},{5,2,0,2,1},{5,0,1,2,1},{5,2,1,2,1},{5,0,2,2,1},{5,2,2,2,1},{5,0,3,2,1},{5,2,3,2,1},
  {6,0,0,1,2},{6,1,0,1,2},{6,2,0,1,2},{6,3,0,1,2},{6,0,2,1,2},{6,1,2,1,2},{6,2,2,1,2},{6,3,2,1,2},
  {7,0,0,1,1},{7,1,0,1,1},{7,2,0,1,1},{7,3,0,1,1},{7,0,1,1,1},{7,1,1,1,1},{7,2,1,1,1},{7,3,1,1,1},
  {7,0,2,1,1},{7,1,2,1,1},{7,2,2,1,1},{7,3,2,1,1},{7,0,3,1,1},{7,1,3,1,1},{7,2,3,1,1},{7,3,3,1,1}};

struct ReqS {
  int active;
  int cx, cy;              // search centre, integer-pel displacement
  int px, py;              // predictor, quarter-pel
  int dlo_x, dhi_x, dlo_y, dhi_y;   // displacement clamp range of this request
  int lam;
  int ffs;                 // FAST_FULL semantics (max_mvd guard)
  int req;                 // index of the request in the caller's array
  unsigned long long init; // min_mcost << IDX_BITS
};

__device__ __forceinline__ unsigned sad4(unsigned a, unsigned b, unsigned c) {
  unsigned d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// all candidates d that clamp onto displacement (Dx,Dy); rare (picture borders only)
__device__ __noinline__ void eval_border(const ReqS *q, unsigned sad, int Dx, int Dy, int R, int max_mvd_m1,
                                         unsigned long long *best) {
  if (Dx < q->dlo_x || Dx > q->dhi_x || Dy < q->dlo_y || Dy > q->dhi_y) return;
  int x0 = Dx, x1 = Dx, y0 = Dy, y1 = Dy;
  if (Dx == q->dlo_x) x0 = q->cx - R;     // every d <= Dlo clamps to Dlo
  if (Dx == q->dhi_x) x1 = q->cx + R;
  if (Dy == q->dlo_y) y0 = q->cy - R;
  if (Dy == q->dhi_y) y1 = q->cy + R;
  x0 = max(x0, q->cx - R); x1 = min(x1, q->cx + R);
  y0 = max(y0, q->cy - R); y1 = min(y1, q->cy + R);
  unsigned long long k = ~0ull;
  for (int dy = y0; dy <= y1; dy++)
    for (int dx = x0; dx <= x1; dx++) {
      int mx = 4 * dx - q->px, my = 4 * dy - q->py;
      if (q->ffs && max(abs(mx), abs(my)) >= max_mvd_m1) continue;
      unsigned long long cost = ((unsigned long long)sad << 5) + (unsigned long long)((long long)q->lam * (jmb_mvbits(mx) + jmb_mvbits(my)));
      k = min(k, (cost << IDX_BITS) | (unsigned)jmb_spiral_index(dx - q->cx, dy - q->cy));
    }
  if (k < *best) atomicMin(best, k);
}

__device__ __forceinline__ void eval(const ReqS *q, unsigned sad, int Dx, int Dy, int R, int max_mvd_m1,
                                     unsigned long long *best) {
  if (!q->active) return;
  unsigned long long cur = *(volatile unsigned long long *)best;
  if (((unsigned long long)sad << (5 + IDX_BITS)) > cur) return;     // cost >= SAD<<5 cannot win
  int ex = Dx - q->cx, ey = Dy - q->cy;
  bool interior = Dx > q->dlo_x && Dx < q->dhi_x && Dy > q->dlo_y && Dy < q->dhi_y;
  if (interior) {
    if (abs(ex) > R || abs(ey) > R) return;
    int mx = 4 * Dx - q->px, my = 4 * Dy - q->py;
    if (q->ffs && max(abs(mx), abs(my)) >= max_mvd_m1) return;
    unsigned long long cost = ((unsigned long long)sad << 5) + (unsigned long long)((long long)q->lam * (jmb_mvbits(mx) + jmb_mvbits(my)));
    unsigned long long k = (cost << IDX_BITS) | (unsigned)jmb_spiral_index(ex, ey);
    if (k < cur) atomicMin(best, k);
  } else {
    eval_border(q, sad, Dx, Dy, R, max_mvd_m1, best);
  }
}

// inverse of jmb_spiral_index
__device__ void spiral_xy(int idx, int *dx, int *dy) {
  if (idx == 0) { *dx = 0; *dy = 0; return; }
  int s = (int)sqrtf((float)idx);
  while (s * s > idx) s--;
  while ((s + 1) * (s + 1) <= idx) s++;
  int l = (s + 1) / 2, base = (2 * l - 1) * (2 * l - 1), off = idx - base;
  if (off < 2 * (2 * l - 1)) { *dx = (off >> 1) - l + 1; *dy = (off & 1) ? l : -l; }
  else { off -= 2 * (2 * l - 1); *dy = (off >> 1) - l; *dx = (off & 1) ? l : -l; }
}

// groups == nullptr: frame layout, group g = requests [41g, 41g+41) in canonical partition order.
__global__ void __launch_bounds__(256)
k_int_search(const jmb_me_req *__restrict__ reqs, const int *__restrict__ groups, jmb_me_res *__restrict__ res,
             const uint8_t *__restrict__ cur, int cur_pitch,
             const uint8_t *const *__restrict__ ref_planes, int ref_pitch, int w, int h, int R, int max_mvd_m1) {
  __shared__ ReqS rq[NPART];
  __shared__ unsigned long long best[NPART];
  __shared__ __align__(16) uint8_t win[WIN_ROWS * WIN_PITCH];
  __shared__ unsigned ssrc[16 * 4];
  __shared__ int sbox[8];

  const int tid = threadIdx.x, g = blockIdx.x;
  const int W = w + 2 * JMB_PAD_X, H = h + 2 * JMB_PAD_Y;

  if (tid < NPART) {
    int ri = groups ? groups[g * NPART + tid] : g * NPART + tid;
    ReqS q; q.active = 0; q.req = ri;
    if (ri >= 0) {
      jmb_me_req r = reqs[ri];
      if (!(r.flags & JMB_REQ_SKIP_INT)) {
        q.active = 1;
        q.ffs = (r.mode == JMB_SEARCH_FAST_FULL);
        int ox = q.ffs ? (r.pos_x & ~15) : r.pos_x, oy = q.ffs ? (r.pos_y & ~15) : r.pos_y;   // clamped origin
        q.dlo_x = -JMB_PAD_X - ox; q.dhi_x = (w + JMB_PAD_X - 1 - 16) - ox;      // UMVLine4X, refbuf.h:25
        q.dlo_y = -JMB_PAD_Y - oy; q.dhi_y = (h + JMB_PAD_Y - 1 - 16) - oy;
        q.cx = r.center_x >> 2; q.cy = r.center_y >> 2;
        q.px = r.pred_x; q.py = r.pred_y; q.lam = r.lambda[0];
        q.init = (unsigned long long)r.min_mcost << IDX_BITS;
      }
    }
    rq[tid] = q;
    best[tid] = q.active ? q.init : 0ull;
  }
  __syncthreads();
  if (tid == 0) {
    int x0 = 1 << 30, x1 = -(1 << 30), y0 = 1 << 30, y1 = -(1 << 30), mbx = 0, mby = 0, rf = 0, any = 0;
    for (int p = 0; p < NPART; p++) if (rq[p].active) {
      const ReqS &q = rq[p];
      x0 = min(x0, jmb_clip(q.dlo_x, q.dhi_x, q.cx - R)); x1 = max(x1, jmb_clip(q.dlo_x, q.dhi_x, q.cx + R));
      y0 = min(y0, jmb_clip(q.dlo_y, q.dhi_y, q.cy - R)); y1 = max(y1, jmb_clip(q.dlo_y, q.dhi_y, q.cy + R));
      jmb_me_req r = reqs[q.req];
      mbx = r.pos_x & ~15; mby = r.pos_y & ~15; rf = r.ref; any = 1;
    }
    sbox[0] = x0; sbox[1] = x1; sbox[2] = y0; sbox[3] = y1; sbox[4] = mbx; sbox[5] = mby; sbox[6] = rf; sbox[7] = any;
  }
  __syncthreads();
  if (!sbox[7]) return;   // nothing but sub-pel-only requests in this group
  const int bx0 = sbox[0], bx1 = sbox[1], by0 = sbox[2], by1 = sbox[3], mbx = sbox[4], mby = sbox[5];
  const uint8_t *ref = ref_planes[sbox[6]];

  if (tid < 64) {   // source macroblock, one 32-bit word = 4 samples
    int r = tid >> 2, c = tid & 3;
    ssrc[tid] = *(const unsigned *)(cur + (size_t)(mby + r) * cur_pitch + mbx + 4 * c);
  }

  for (int cy0 = by0; cy0 <= by1; cy0 += CH) {
    const int ch = min(CH, by1 - cy0 + 1);
    for (int cx0 = bx0; cx0 <= bx1; cx0 += CW) {
      const int cw = min(CW, bx1 - cx0 + 1);
      __syncthreads();
      // stage the window: rows cy0 .. cy0+ch+14, columns cx0 .. cx0+cw+14 (+3), relative to the MB origin;
      // coordinates outside the padded plane are clamped (such samples are never used by a valid candidate)
      const int wcols = ((cw + 3) & ~3) + 16, wrows = ch + 15;
      for (int i = tid; i < wrows * (wcols >> 2); i += 256) {
        int r = i / (wcols >> 2), c4 = (i - r * (wcols >> 2)) * 4;
        int py = jmb_clip(0, H - 1, mby + cy0 + r + JMB_PAD_Y);
        unsigned v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          int px = jmb_clip(0, W - 1, mbx + cx0 + c4 + k + JMB_PAD_X);
          v |= (unsigned)ref[(size_t)py * ref_pitch + px] << (8 * k);
        }
        *(unsigned *)(win + r * WIN_PITCH + c4) = v;
      }
      __syncthreads();

      const int nx4 = (cw + 3) >> 2, mid = ch >> 1;
      for (int it = tid; it < nx4 * ch; it += 256) {
        const int k = it / nx4, ix4 = it - k * nx4;
        const int row = (k & 1) ? mid - ((k + 1) >> 1) : mid + (k >> 1);   // centre rows first: tight bounds early
        unsigned acc[4][16];
#pragma unroll
        for (int s = 0; s < 4; s++)
#pragma unroll
          for (int b = 0; b < 16; b++) acc[s][b] = 0;
#pragma unroll
        for (int r = 0; r < 16; r++) {
          const unsigned *wp = (const unsigned *)(win + (row + r) * WIN_PITCH + ix4 * 4);
          unsigned w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
          unsigned wv[5] = {w0, w1, w2, w3, w4};
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const unsigned sv = ssrc[r * 4 + c];
            const int b = (r >> 2) * 4 + c;
            acc[0][b] = sad4(wv[c], sv, acc[0][b]);
            acc[1][b] = sad4(__byte_perm(wv[c], wv[c + 1], 0x4321), sv, acc[1][b]);
            acc[2][b] = sad4(__byte_perm(wv[c], wv[c + 1], 0x5432), sv, acc[2][b]);
            acc[3][b] = sad4(__byte_perm(wv[c], wv[c + 1], 0x6543), sv, acc[3][b]);
          }
        }
        const int Dy = cy0 + row;
#pragma unroll
        for (int s = 0; s < 4; s++) {
          const int Dx = cx0 + ix4 * 4 + s;
          if (Dx > bx1) break;
          const unsigned *a = acc[s];
          // partition sums, lencod/src/me_fullfast.c:207-259
          unsigned s6[8], s5[8], s4[4];
#pragma unroll
          for (int i = 0; i < 4; i++) { s6[i] = a[i] + a[4 + i]; s6[4 + i] = a[8 + i] + a[12 + i]; }
#pragma unroll
          for (int i = 0; i < 8; i++) s5[i] = a[2 * i] + a[2 * i + 1];
          s4[0] = s6[0] + s6[1]; s4[1] = s6[2] + s6[3]; s4[2] = s6[4] + s6[5]; s4[3] = s6[6] + s6[7];
          const unsigned s3a = s4[0] + s4[2], s3b = s4[1] + s4[3], s2a = s4[0] + s4[1], s2b = s4[2] + s4[3];
#define EV(p, v) eval(&rq[p], (v), Dx, Dy, R, max_mvd_m1, &best[p])
          EV(0, s2a + s2b);
          EV(1, s2a); EV(2, s2b);
          EV(3, s3a); EV(4, s3b);
          EV(5, s4[0]); EV(6, s4[1]); EV(7, s4[2]); EV(8, s4[3]);
#pragma unroll
          for (int i = 0; i < 8; i++) EV(9 + i, s5[i]);     // 8x4: (bx 0|2, by 0..3) raster = pairs (2i, 2i+1)
#pragma unroll
          for (int i = 0; i < 8; i++) EV(17 + i, s6[i]);    // 4x8: (bx 0..3, by 0|2)
#pragma unroll
          for (int i = 0; i < 16; i++) EV(25 + i, a[i]);
#undef EV
        }
      }
    }
  }
  __syncthreads();
  if (tid < NPART && rq[tid].active) {
    const ReqS &q = rq[tid];
    unsigned long long k = best[tid];
    int dx = 0, dy = 0;
    long long cost = (long long)(k >> IDX_BITS);
    if (k != q.init) spiral_xy((int)(k & ((1u << IDX_BITS) - 1)), &dx, &dy);
    jmb_me_res o;
    o.imv_x = o.mv_x = (int16_t)(4 * (q.cx + dx));
    o.imv_y = o.mv_y = (int16_t)(4 * (q.cy + dy));
    o.icost = o.cost = cost;
    res[q.req] = o;
  }
}

// BlockSAD surfaces of one macroblock in JM's layout and spiral order (jmb_ffs_surfaces).
__global__ void k_ffs_surfaces(const uint8_t *__restrict__ cur, int cur_pitch, const uint8_t *__restrict__ ref, int ref_pitch,
                               int w, int h, int mbx, int mby, int cx, int cy, int R, uint32_t *__restrict__ out) {
  const int max_pos = (2 * R + 1) * (2 * R + 1);
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= max_pos) return;
  int l = 0, dx = 0, dy = 0;
  if (pos) {
    int s = (int)sqrtf((float)pos);
    while (s * s > pos) s--;
    while ((s + 1) * (s + 1) <= pos) s++;
    l = (s + 1) / 2;
    int base = (2 * l - 1) * (2 * l - 1), off = pos - base;
    if (off < 2 * (2 * l - 1)) { dx = (off >> 1) - l + 1; dy = (off & 1) ? l : -l; }
    else { off -= 2 * (2 * l - 1); dy = (off >> 1) - l; dx = (off & 1) ? l : -l; }
  }
  // macroblock origin clamped once per position, me_fullfast.c:498
  int X = jmb_clip(-JMB_PAD_X, w + JMB_PAD_X - 1 - 16, mbx + cx + dx);
  int Y = jmb_clip(-JMB_PAD_Y, h + JMB_PAD_Y - 1 - 16, mby + cy + dy);
  unsigned a[16];
  for (int b = 0; b < 16; b++) {
    int bx = (b & 3) * 4, by = (b >> 2) * 4;
    unsigned s = 0;
    for (int y = 0; y < 4; y++) {
      unsigned sv = *(const unsigned *)(cur + (size_t)(mby + by + y) * cur_pitch + mbx + bx);
      const uint8_t *rp = ref + (size_t)(Y + by + y + JMB_PAD_Y) * ref_pitch + (X + bx + JMB_PAD_X);
      unsigned rv = rp[0] | (rp[1] << 8) | (rp[2] << 16) | ((unsigned)rp[3] << 24);
      s = __vsadu4(sv, rv) + s;
    }
    a[b] = s;
  }
#define O(t, i) out[((size_t)(t) * 16 + (i)) * max_pos + pos]
  for (int i = 0; i < 16; i++) O(7, i) = a[i];
  unsigned s6[16], s4[16];
  for (int i = 0; i < 4; i++) { s6[i] = a[i] + a[i + 4]; s6[8 + i] = a[8 + i] + a[12 + i]; O(6, i) = s6[i]; O(6, 8 + i) = s6[8 + i]; }
  for (int i = 0; i < 16; i += 2) O(5, i) = a[i] + a[i + 1];
  s4[0] = s6[0] + s6[1]; s4[2] = s6[2] + s6[3]; s4[8] = s6[8] + s6[9]; s4[10] = s6[10] + s6[11];
  O(4, 0) = s4[0]; O(4, 2) = s4[2]; O(4, 8) = s4[8]; O(4, 10) = s4[10];
  O(3, 0) = s4[0] + s4[8]; O(3, 2) = s4[2] + s4[10];
  O(2, 0) = s4[0] + s4[2]; O(2, 8) = s4[8] + s4[10];
  O(1, 0) = s4[0] + s4[2] + s4[8] + s4[10];
#undef O
}

}  // namespace

int jmb_launch_refine(jmb_ctx *ctx, const jmb_me_req *d_reqs, jmb_me_res *d_res, int n, const uint8_t *const *d_ref_planes);

// canonical slot of a request inside its macroblock group (order of c_part)
static int part_slot(const jmb_me_req &r) {
  static const int base[8] = {0, 0, 1, 3, 5, 9, 17, 25};
  static const int w4[8] = {4, 4, 4, 2, 2, 2, 1, 1}, h4[8] = {4, 4, 2, 4, 2, 1, 2, 1};
  int t = r.blocktype, bx = (r.pos_x & 15) >> 2, by = (r.pos_y & 15) >> 2;
  return base[t] + (by / h4[t]) * (4 / w4[t]) + bx / w4[t];
}

static int validate_req(jmb_ctx *ctx, const jmb_me_req &r, int i) {
  static const int bsx[8] = {0, 16, 16, 8, 8, 8, 4, 4}, bsy[8] = {0, 16, 8, 16, 8, 4, 8, 4};
  if (r.blocktype < 1 || r.blocktype > 7) return jmb_fail(ctx, JMB_ERR_ARG, "request %d: blocktype %d", i, r.blocktype);
  if (r.ref >= ctx->nref) return jmb_fail(ctx, JMB_ERR_ARG, "request %d: ref %d of %d", i, r.ref, ctx->nref);
  if (r.pos_x < 0 || r.pos_y < 0 || r.pos_x + bsx[r.blocktype] > ctx->cur_w || r.pos_y + bsy[r.blocktype] > ctx->cur_h ||
      (r.pos_x % bsx[r.blocktype]) || (r.pos_y % bsy[r.blocktype]))
    return jmb_fail(ctx, JMB_ERR_ARG, "request %d: block (%d,%d) type %d outside / misaligned in %dx%d", i, r.pos_x, r.pos_y,
                    r.blocktype, ctx->cur_w, ctx->cur_h);
  if (!(r.flags & JMB_REQ_SKIP_INT) && ((r.center_x | r.center_y) & 3))
    return jmb_fail(ctx, JMB_ERR_ARG, "request %d: search centre (%d,%d) is not integer-pel", i, r.center_x, r.center_y);
  if (r.mode > JMB_SEARCH_FAST_FULL) return jmb_fail(ctx, JMB_ERR_ARG, "request %d: mode %d", i, r.mode);
  if (r.min_mcost < 0 || r.min_mcost > ((int64_t)1 << 48)) return jmb_fail(ctx, JMB_ERR_ARG, "request %d: min_mcost out of range", i);
  return 0;
}

static int upload_ref_table(jmb_ctx *ctx, const uint8_t *const **d_tab, int plane) {
  // table of plane[0][0] (or all-16 base) pointers of the picture's reference list, in d_groups' tail
  (void)plane;
  const uint8_t *tab[JMB_MAX_REFS];
  for (int i = 0; i < JMB_MAX_REFS; i++) tab[i] = i < ctx->nref ? ctx->refs[ctx->ref_list[i]].planes : nullptr;
  static_assert(sizeof(tab) == JMB_MAX_REFS * sizeof(void *), "");
  int rc = jmb_reserve_dev(ctx, &ctx->d_reftab, &ctx->d_reftab_cap, sizeof(tab));
  if (rc) return rc;
  JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_reftab, tab, sizeof(tab), cudaMemcpyHostToDevice, ctx->stream));
  *d_tab = (const uint8_t *const *)ctx->d_reftab;
  return 0;
}

static int me_search_impl(jmb_ctx *ctx, const jmb_me_req *reqs, int n, jmb_me_res *res, int loc, bool frame_layout) {
  if (n <= 0) return JMB_OK;
  if (!ctx->cur || ctx->nref == 0) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_me_search: call jmb_pic_begin with >= 1 reference first");
  if (ctx->me.search_range > 45) return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "search_range %d > 45", ctx->me.search_range);
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const jmb_ref &r0 = ctx->refs[ctx->ref_list[0]];
  const uint8_t *const *d_tab = nullptr;
  int rc = upload_ref_table(ctx, &d_tab, 0);
  if (rc) return rc;

  const jmb_me_req *d_reqs = reqs; jmb_me_res *d_res = res;
  const int *d_groups = nullptr; int n_groups = 0;
  bool any_subpel = false;
  const jmb_me_req *h_reqs = nullptr;

  if (loc == JMB_HOST) h_reqs = reqs;
  else if (!frame_layout) {
    // grouping needs the request headers on the host
    rc = jmb_reserve_host(ctx, &ctx->h_stage, &ctx->h_stage_cap, (size_t)n * sizeof(jmb_me_req)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->h_stage, reqs, (size_t)n * sizeof(jmb_me_req), cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    h_reqs = (const jmb_me_req *)ctx->h_stage;
  }
  if (h_reqs) {
    for (int i = 0; i < n; i++) { rc = validate_req(ctx, h_reqs[i], i); if (rc) return rc; any_subpel |= (h_reqs[i].flags & JMB_REQ_SUBPEL) != 0; }
  } else any_subpel = true;

  if (frame_layout) {
    if (n % NPART) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_search_frame: n=%d is not a multiple of 41", n);
    n_groups = n / NPART;
    if (h_reqs)
      for (int i = 0; i < n; i++)
        if (part_slot(h_reqs[i]) != i % NPART || (h_reqs[i].pos_x & ~15) != (h_reqs[i - i % NPART].pos_x & ~15) ||
            (h_reqs[i].pos_y & ~15) != (h_reqs[i - i % NPART].pos_y & ~15) || h_reqs[i].ref != h_reqs[i - i % NPART].ref)
          return jmb_fail(ctx, JMB_ERR_ARG, "jmb_me_search_frame: request %d is not partition %d of its macroblock", i, i % NPART);
  } else {
    // consecutive requests of one (macroblock, ref) form a group; a repeated partition starts a new one
    rc = jmb_reserve_host(ctx, &ctx->h_groups, &ctx->h_groups_cap, (size_t)n * NPART * sizeof(int)); if (rc) return rc;
    int *hg = (int *)ctx->h_groups;
    int kx = -1, ky = -1, kr = -1;
    for (int i = 0; i < n; i++) {
      const jmb_me_req &q = h_reqs[i];
      int slot = part_slot(q);
      bool fresh = n_groups == 0 || (q.pos_x & ~15) != kx || (q.pos_y & ~15) != ky || q.ref != kr || hg[(n_groups - 1) * NPART + slot] >= 0;
      if (fresh) {
        for (int p = 0; p < NPART; p++) hg[n_groups * NPART + p] = -1;
        n_groups++; kx = q.pos_x & ~15; ky = q.pos_y & ~15; kr = q.ref;
      }
      hg[(n_groups - 1) * NPART + slot] = i;
    }
    rc = jmb_reserve_dev(ctx, &ctx->d_groups, &ctx->d_groups_cap, (size_t)n_groups * NPART * sizeof(int)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_groups, hg, (size_t)n_groups * NPART * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    d_groups = (const int *)ctx->d_groups;
  }
  if (loc == JMB_HOST) {
    rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, (size_t)n * sizeof(jmb_me_req)); if (rc) return rc;
    rc = jmb_reserve_dev(ctx, &ctx->d_stage2, &ctx->d_stage2_cap, (size_t)n * sizeof(jmb_me_res)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, reqs, (size_t)n * sizeof(jmb_me_req), cudaMemcpyHostToDevice, ctx->stream));
    d_reqs = (const jmb_me_req *)ctx->d_stage; d_res = (jmb_me_res *)ctx->d_stage2;
  }
  jmb_time_begin(ctx, JMB_K_INT_SEARCH);
  k_int_search<<<n_groups, 256, 0, ctx->stream>>>(d_reqs, d_groups, d_res, ctx->cur, ctx->cur_pitch, d_tab, r0.pitch,
                                                  ctx->cur_w, ctx->cur_h, ctx->me.search_range, ctx->me.max_mvd - 1);
  jmb_time_end(ctx, JMB_K_INT_SEARCH);
  JMB_LAUNCH_CHECK(ctx);
  if (any_subpel) { rc = jmb_launch_refine(ctx, d_reqs, d_res, n, d_tab); if (rc) return rc; }
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(res, d_res, (size_t)n * sizeof(jmb_me_res), cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

extern "C" {

int jmb_me_search(jmb_ctx *ctx, const jmb_me_req *reqs, int n, jmb_me_res *res, int loc) {
  return me_search_impl(ctx, reqs, n, res, loc, false);
}

int jmb_me_search_frame(jmb_ctx *ctx, const jmb_me_req *reqs, int n_mb, jmb_me_res *res, int loc) {
  return me_search_impl(ctx, reqs, n_mb * NPART, res, loc, true);
}

int jmb_ffs_surfaces(jmb_ctx *ctx, int ref, int mb_x, int mb_y, int center_x, int center_y, uint32_t *out, int loc) {
  if (!ctx->cur || ref < 0 || ref >= ctx->nref) return jmb_fail(ctx, JMB_ERR_STATE, "jmb_ffs_surfaces: no picture / bad ref %d", ref);
  if ((mb_x & 15) || (mb_y & 15) || mb_x < 0 || mb_y < 0 || mb_x + 16 > ctx->cur_w || mb_y + 16 > ctx->cur_h || ((center_x | center_y) & 3))
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_ffs_surfaces: macroblock (%d,%d) centre (%d,%d)", mb_x, mb_y, center_x, center_y);
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const jmb_ref &r = ctx->refs[ctx->ref_list[ref]];
  const int R = ctx->me.search_range, max_pos = (2 * R + 1) * (2 * R + 1);
  size_t bytes = (size_t)8 * 16 * max_pos * sizeof(uint32_t);
  uint32_t *d_out = out;
  if (loc == JMB_HOST) {
    int rc = jmb_reserve_dev(ctx, &ctx->d_stage2, &ctx->d_stage2_cap, bytes); if (rc) return rc;
    d_out = (uint32_t *)ctx->d_stage2;
  }
  jmb_time_begin(ctx, JMB_K_FFS_SURF);
  k_ffs_surfaces<<<(max_pos + 127) / 128, 128, 0, ctx->stream>>>(ctx->cur, ctx->cur_pitch, r.planes, r.pitch, r.w, r.h, mb_x, mb_y,
                                                                 center_x >> 2, center_y >> 2, R, d_out);
  jmb_time_end(ctx, JMB_K_FFS_SURF);
  JMB_LAUNCH_CHECK(ctx);
  if (loc == JMB_HOST) {
    JMB_CUDA(ctx, cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    JMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return JMB_OK;
}

}  // extern "C"
