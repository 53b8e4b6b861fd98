// k_deblock.cu -- DeblockFrame (lencod/src/loopFilter.c:63-299) with the non-MBAFF strength and edge functions of
// lencod/src/loop_filter_normal.c: frame pictures, 8 bit, 4:0:0 / 4:2:0 / 4:2:2.
//
// The standard filters macroblock after macroblock in raster order, vertical edges before horizontal ones, and the order shows
// in the result: the left edge of a macroblock reads samples its left neighbour's horizontal edges have changed, the top edge
// reads samples the top-right neighbour's left edge has changed.  So macroblock (x, y) needs (x-1, y) and (x+1, y-1) finished
// (the latter implies (x, y-1)) and nothing else: a wavefront x + 2y (JM's own JM_PARALLEL_DEBLOCK walks the same diagonals).
// One CTA of two warps per macroblock.  CTAs take macroblock numbers from a ticket counter in raster order -- whatever a CTA
// waits for was handed out before it, to a CTA that is running or done -- and wait on per-macroblock flags in global memory.
// A CTA copies its macroblock and the 4 samples left of / above it into shared memory, derives the 32 edge strengths (one per
// lane: direction x edge x 4-sample segment), filters the vertical edges row-parallel and the horizontal ones column-parallel
// there (luma in one warp, chroma in the other), and writes back what it may have changed.  Nobody else touches that area in
// between.
#include "jmb_internal.h"

namespace {

__constant__ unsigned char c_db_alpha[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,4,4,5,6,7,8,9,10,12,13,15,17,20,22,25,28,32,36,40,45,50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255};
__constant__ unsigned char c_db_beta[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13,14,14,15,15,16,16,17,17,18,18};
// CLIP_TAB[indexA][1..3] (lencod/inc/loop_filter.h:36-45)
__constant__ unsigned char c_db_clip[52][3] = {
  {0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},
  {0,0,1},{0,0,1},{0,0,1},{0,0,1},{0,1,1},{0,1,1},{1,1,1},{1,1,1},{1,1,1},{1,1,1},{1,1,2},{1,1,2},{1,1,2},{1,1,2},{1,2,3},{1,2,3},{2,2,3},
  {2,2,4},{2,3,4},{2,3,4},{3,3,5},{3,4,6},{3,4,6},{4,5,7},{4,5,8},{4,6,9},{5,7,10},{6,8,11},{6,8,13},{7,10,14},{8,11,16},{9,12,18},{10,13,20},
  {11,15,23},{13,17,25}};

struct DbArgs {
  uint8_t *luma, *cb, *cr;
  int pitch, pitch_c, mbw, mbh, yuv, slice_type, d8;
  const jmb_db_mb *mbs;
  unsigned *ticket;          // next macroblock to hand out
  int *done;                 // done[mb] == serial: finished in this call
  int serial;
};

constexpr int LP = 24, CP = 12;      // shared tile pitches: luma rows hold columns -4..15 (+4 spare), chroma rows columns -4..7

__device__ __forceinline__ bool db_intra(int t) { return t == 9 || t == 10 || t == 13 || t == 14; }
__device__ __forceinline__ int db_mvdiff(const int16_t *a, const int16_t *b) { return (abs(a[0] - b[0]) >= 4) | (abs(a[1] - b[1]) >= 4); }      // compare_mvs, mvlimit 4

// GetStrengthVer / GetStrengthHor (loop_filter_normal.c:53-300): segment k of edge `edge` in direction dir; P = the macroblock across the edge
__device__ int db_strength(int dir, int edge, int k, const jmb_db_mb &Q, const jmb_db_mb &P) {
  if (db_intra(Q.mb_type) || db_intra(P.mb_type)) return edge == 0 ? 4 : 3;
  const int bq = dir ? edge * 4 + k : k * 4 + edge;
  const int bp = edge ? (dir ? bq - 4 : bq - 1) : (dir ? 12 + k : k * 4 + 3);
  if (((Q.cbp_blk >> bq) & 1) || ((P.cbp_blk >> bp) & 1)) return 2;
  if (edge && (Q.mb_type == 1 || Q.mb_type == (dir ? 3 : 2))) return 0;
  const int p0 = Q.ref_id[0][bq], p1 = Q.ref_id[1][bq], q0 = P.ref_id[0][bp], q1 = P.ref_id[1][bp];
  if (!((p0 == q0 && p1 == q1) || (p0 == q1 && p1 == q0))) return 1;
  const int16_t *mp0 = Q.mv[0][bq], *mp1 = Q.mv[1][bq], *mq0 = P.mv[0][bp], *mq1 = P.mv[1][bp];
  if (p0 != p1) return p0 == q0 ? (db_mvdiff(mp0, mq0) | db_mvdiff(mp1, mq1)) : (db_mvdiff(mp0, mq1) | db_mvdiff(mp1, mq0));
  return (db_mvdiff(mp0, mq0) | db_mvdiff(mp1, mq1)) && (db_mvdiff(mp0, mq1) | db_mvdiff(mp1, mq0));
}

// One line of samples across a luma edge in the standard's notation (H.264 8.7.2.3 / 8.7.2.4; JM: EdgeLoopLumaVer / Hor,
// loop_filter_normal.c:310-575): q points at q0, `st` is the step away from the edge, p0 = q[-st].
__device__ __forceinline__ void db_luma_line(uint8_t *q, int st, int bS, int alpha, int beta, int tc0) {
  const int p0 = q[-st], q0 = q[0], p1 = q[-2 * st], q1 = q[st];
  if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;      // filterSamplesFlag
  const int p2 = q[-3 * st], q2 = q[2 * st];
  const bool ap = abs(p2 - p0) < beta, aq = abs(q2 - q0) < beta;
  if (bS == 4) {
    const bool strong = abs(p0 - q0) < (alpha >> 2) + 2;
    if (ap && strong) {
      const int p3 = q[-4 * st];
      q[-st] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      q[-2 * st] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
      q[-3 * st] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
    } else q[-st] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
    if (aq && strong) {
      const int q3 = q[3 * st];
      q[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      q[st] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
      q[2 * st] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
    } else q[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
  } else {
    const int tc = tc0 + ap + aq, avg = (p0 + q0 + 1) >> 1;
    const int delta = jmb_clip(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
    if (ap) q[-2 * st] = (uint8_t)(p1 + jmb_clip(-tc0, tc0, (p2 + avg - 2 * p1) >> 1));
    if (aq) q[st] = (uint8_t)(q1 + jmb_clip(-tc0, tc0, (q2 + avg - 2 * q1) >> 1));
    if (delta) { q[-st] = (uint8_t)jmb_clip(0, 255, p0 + delta); q[0] = (uint8_t)jmb_clip(0, 255, q0 - delta); }
  }
}

// the same for a chroma edge (chromaStyleFilteringFlag: only p0 and q0 change; JM: EdgeLoopChromaVer / Hor, :585-758)
__device__ __forceinline__ void db_chroma_line(uint8_t *q, int st, int bS, int alpha, int beta, int tc0) {
  const int p0 = q[-st], q0 = q[0], p1 = q[-2 * st], q1 = q[st];
  if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
  if (bS == 4) { q[-st] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2); q[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2); }
  else {
    const int tc = tc0 + 1, delta = jmb_clip(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
    if (delta) { q[-st] = (uint8_t)jmb_clip(0, 255, p0 + delta); q[0] = (uint8_t)jmb_clip(0, 255, q0 - delta); }
  }
}

// Two warps per macroblock: warp 0 filters luma, warp 1 the two chroma planes (independent data, the same strengths).
// What does not depend on the neighbours -- the macroblock records, the macroblock's own samples (nobody filters into them before
// its turn) and the 32 strengths -- is fetched and worked out BEFORE the wait; after it only the 4 columns left of and the
// 4 rows above the macroblock are read.
__global__ void __launch_bounds__(64)
k_deblock(const DbArgs A) {
  __shared__ jmb_db_mb M[3];                        // this macroblock, its left and its upper neighbour
  __shared__ __align__(4) uint8_t L[20 * LP];       // luma rows -4..15
  __shared__ __align__(4) uint8_t C[2][20 * CP];    // chroma rows -4..15 (4:2:0 uses -4..7)
  __shared__ unsigned char sbs[32];                 // strength of [dir][edge][segment]; 0 where DeblockMb passes the edge over
  __shared__ int s_mb;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_mb = (int)atomicAdd(A.ticket, 1u);
  __syncthreads();
  const int mb = s_mb, x = mb % A.mbw, y = mb / A.mbw;
  const int chh = A.yuv == 1 ? 8 : 16;
  // ---- before the wait ----
  {
    const unsigned *src[3] = {(const unsigned *)&A.mbs[mb], (const unsigned *)&A.mbs[x > 0 ? mb - 1 : mb], (const unsigned *)&A.mbs[y > 0 ? mb - A.mbw : mb]};
    for (int i = tid; i < 3 * 44; i += 64) ((unsigned *)M)[i] = __ldg(src[i / 44] + i % 44);
  }
  if (warp == 0) {
    for (int i = lane; i < 16 * 4; i += 32) {      // own luma samples: rows 0..15, words 0..3
      const int r = i >> 2, wd = i & 3;
      *(unsigned *)&L[(r + 4) * LP + (wd + 1) * 4] = __ldcg((const unsigned *)(A.luma + (size_t)(y * 16 + r) * A.pitch + x * 16 + wd * 4));
    }
  } else if (A.yuv) {
    for (int i = lane; i < 2 * chh * 2; i += 32) {      // own chroma samples: rows 0..chh-1, words 0..1 of both planes
      const int pl = i / (chh * 2), j = i % (chh * 2), r = j >> 1, wd = j & 1;
      *(unsigned *)&C[pl][(r + 4) * CP + (wd + 1) * 4] = __ldcg((const unsigned *)((pl ? A.cr : A.cb) + (size_t)(y * chh + r) * A.pitch_c + x * 8 + wd * 4));
    }
  }
  __syncthreads();
  const jmb_db_mb &Q = M[0];
  const bool filter_on = Q.df_disable_idc != 1;
  const bool t8 = Q.flags & JMB_DB_T8X8, cbp = Q.flags & JMB_DB_CBP;
  if (warp == 0 && filter_on) {      // the 32 strengths, one per lane; DeblockMb's reasons to pass an edge over (loopFilter.c:150-166, :206-222) make it 0
    const int dir = lane >> 4, edge = (lane >> 2) & 3, k = lane & 3;
    bool on = edge ? true : (Q.df_disable_idc == 2 ? (Q.flags & (dir ? JMB_DB_AVAIL_B : JMB_DB_AVAIL_A)) != 0 : (dir ? y : x) != 0);
    if (!cbp) {
      const bool luma_on = !(t8 && (edge & 1));
      if (!luma_on && (dir == 0 || A.yuv == 1)) on = false;
      else if (edge > 0 && (A.slice_type == 0 || A.slice_type == 1)) {
        if ((Q.mb_type == 0 && A.slice_type == 0) || Q.mb_type == 1 || Q.mb_type == (dir ? 3 : 2)) on = false;
        else if ((edge & 1) && (Q.mb_type == (dir ? 2 : 3) || (Q.mb_type == 0 && A.slice_type == 1 && A.d8))) on = false;
      }
    }
    sbs[lane] = on ? (unsigned char)db_strength(dir, edge, k, Q, edge ? Q : M[1 + dir]) : 0;
  }
  // ---- the wait: (x-1, y) and (x+1, y-1) -- (x, y-1) in the last column -- must be through ----
  if (tid == 0) {
    if (x > 0) while (*(volatile int *)&A.done[mb - 1] != A.serial) __nanosleep(20);
    if (y > 0) { const int dep = mb - A.mbw + (x < A.mbw - 1 ? 1 : 0); while (*(volatile int *)&A.done[dep] != A.serial) __nanosleep(20); }
    __threadfence();
  }
  __syncthreads();
  if (filter_on) {
    // the strips the neighbours have just finished: every read goes to L2 (another SM wrote them)
    if (warp == 0) {
      if (lane < 16) { if (x > 0) *(unsigned *)&L[(lane + 4) * LP] = __ldcg((const unsigned *)(A.luma + (size_t)(y * 16 + lane) * A.pitch + x * 16 - 4)); }
      else if (y > 0) { const int r = (lane - 16) >> 2, wd = lane & 3; *(unsigned *)&L[r * LP + (wd + 1) * 4] = __ldcg((const unsigned *)(A.luma + (size_t)(y * 16 - 4 + r) * A.pitch + x * 16 + wd * 4)); }
    } else if (A.yuv) {
      if (x > 0) for (int i = lane; i < 2 * chh; i += 32) { const int pl = i / chh, r = i % chh; *(unsigned *)&C[pl][(r + 4) * CP] = __ldcg((const unsigned *)((pl ? A.cr : A.cb) + (size_t)(y * chh + r) * A.pitch_c + x * 8 - 4)); }
      if (y > 0 && lane < 16) { const int pl = lane >> 3, r = (lane >> 1) & 3, wd = lane & 1; *(unsigned *)&C[pl][r * CP + (wd + 1) * 4] = __ldcg((const unsigned *)((pl ? A.cr : A.cb) + (size_t)(y * chh - 4 + r) * A.pitch_c + x * 8 + wd * 4)); }
    }
    __syncwarp();
    // chroma_edge[dir][edge][yuv_format] (loop_filter.h:47-56): where a luma edge's strengths are used in the chroma planes
    auto cedge = [&](int dir, int edge) { return edge == 0 ? 0 : edge == 2 ? (dir && A.yuv == 2 ? 8 : 4) : (dir && A.yuv == 2 ? edge * 4 : -4); };
#pragma unroll 1
    for (int dir = 0; dir < 2; dir++) {
      if (warp == 0) {      // luma: lane = the row (vertical edges) / the column (horizontal edges)
        if (lane < 16)
          for (int edge = 0; edge < 4; edge++) {
            const int s = sbs[dir * 16 + edge * 4 + (lane >> 2)];
            if (!s || (t8 && (edge & 1))) continue;
            const jmb_db_mb &P = edge ? Q : M[1 + dir];
            const int qp = (P.qp + Q.qp + 1) >> 1, ia = jmb_clip(0, 51, qp + Q.df_alpha_c0_offset), ib = jmb_clip(0, 51, qp + Q.df_beta_offset);
            const int alpha = c_db_alpha[ia], beta = c_db_beta[ib];
            if (!(alpha | beta)) continue;
            uint8_t *q = dir ? &L[(4 + edge * 4) * LP + 4 + lane] : &L[(4 + lane) * LP + 4 + edge * 4];
            db_luma_line(q, dir ? LP : 1, s, alpha, beta, s < 4 ? c_db_clip[ia][s - 1] : 0);
          }
      } else if (A.yuv) {      // chroma: vertical edges run over chh rows of each plane, horizontal ones over 8 columns
        const int n = dir ? 8 : chh;
        for (int i = lane; i < 2 * n; i += 32) {
          const int pl = i / n, j = i % n;
          for (int edge = 0; edge < 4; edge++) {
            const int ec = cedge(dir, edge);
            if (ec < 0) continue;
            const int s = sbs[dir * 16 + edge * 4 + (n == 8 ? j >> 1 : j >> 2)];
            if (!s) continue;
            const jmb_db_mb &P = edge ? Q : M[1 + dir];
            const int qp = (P.qpc[pl] + Q.qpc[pl] + 1) >> 1, ia = jmb_clip(0, 51, qp + Q.df_alpha_c0_offset), ib = jmb_clip(0, 51, qp + Q.df_beta_offset);
            const int alpha = c_db_alpha[ia], beta = c_db_beta[ib];
            if (!(alpha | beta)) continue;
            uint8_t *q = dir ? &C[pl][(4 + ec) * CP + 4 + j] : &C[pl][(4 + j) * CP + 4 + ec];
            db_chroma_line(q, dir ? CP : 1, s, alpha, beta, s < 4 ? c_db_clip[ia][s - 1] : 0);
          }
        }
      }
      __syncwarp();
    }
    // write back: the macroblock, the 4 columns left of it and the 4 rows above it (not the corner: nothing there was touched)
    if (warp == 0) {
      for (int i = lane; i < 20 * 5; i += 32) {
        const int r = i / 5 - 4, wd = i % 5 - 1;
        if ((r < 0 && wd < 0) || (r < 0 && y == 0) || (wd < 0 && x == 0)) continue;
        __stcg((unsigned *)(A.luma + (size_t)(y * 16 + r) * A.pitch + x * 16 + wd * 4), *(const unsigned *)&L[(r + 4) * LP + (wd + 1) * 4]);
      }
    } else if (A.yuv) {
      for (int i = lane; i < 2 * (chh + 4) * 3; i += 32) {
        const int pl = i / ((chh + 4) * 3), j = i % ((chh + 4) * 3), r = j / 3 - 4, wd = j % 3 - 1;
        if ((r < 0 && wd < 0) || (r < 0 && y == 0) || (wd < 0 && x == 0)) continue;
        __stcg((unsigned *)((pl ? A.cr : A.cb) + (size_t)(y * chh + r) * A.pitch_c + x * 8 + wd * 4), *(const unsigned *)&C[pl][(r + 4) * CP + (wd + 1) * 4]);
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) *(volatile int *)&A.done[mb] = A.serial;
}

}  // namespace

extern "C" int jmb_deblock_picture(jmb_ctx *ctx, uint8_t *luma, int pitch, uint8_t *cb, uint8_t *cr, int pitch_c, int width, int height, int yuv_format,
                                   int slice_type, int direct_8x8_inference, const jmb_db_mb *mbs, int loc) {
  static_assert(sizeof(jmb_db_mb) == 176, "jmb_db_mb layout");
  if (!luma || !mbs || width < 16 || height < 16 || (width & 15) || (height & 15) || pitch < width)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_deblock_picture: picture %dx%d pitch %d", width, height, pitch);
  if (yuv_format < 0 || yuv_format > 2 || (yuv_format && (!cb || !cr || pitch_c < width / 2)))
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_deblock_picture: yuv_format %d (0 = 4:0:0, 1 = 4:2:0, 2 = 4:2:2) with its two chroma planes", yuv_format);
  if (slice_type < 0 || slice_type > 2) return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "jmb_deblock_picture: slice type %d (SP / SI slices are not handled)", slice_type);
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const int mbw = width / 16, mbh = height / 16, n = mbw * mbh, hc = yuv_format == 1 ? height / 2 : height, wc = width / 2;
  const bool host = jmb_is_host(loc);
  DbArgs A;
  A.luma = luma; A.cb = cb; A.cr = cr; A.pitch = pitch; A.pitch_c = pitch_c; A.mbs = mbs;
  if (host) {
    if (loc == JMB_HOST) for (int i = 0; i < n; i++) {
      const jmb_db_mb &m = mbs[i];
      if (m.df_disable_idc < 0 || m.df_disable_idc > 2 || m.qp < 0 || m.qp > 51 || m.qpc[0] < 0 || m.qpc[0] > 51 || m.qpc[1] < 0 || m.qpc[1] > 51 ||
          m.df_alpha_c0_offset < -12 || m.df_alpha_c0_offset > 12 || m.df_beta_offset < -12 || m.df_beta_offset > 12)
        return jmb_fail(ctx, JMB_ERR_ARG, "jmb_deblock_picture: macroblock %d: qp %d/%d/%d idc %d offsets %d/%d", i, m.qp, m.qpc[0], m.qpc[1], m.df_disable_idc,
                        m.df_alpha_c0_offset, m.df_beta_offset);
    }
    // device copies: planes at a 128-byte pitch, then the macroblock records
    const int dp = (width + 127) & ~127, dpc = (wc + 127) & ~127;
    const size_t lb = (size_t)dp * height, cbytes = yuv_format ? (size_t)dpc * hc : 0, mb_bytes = (size_t)n * sizeof(jmb_db_mb);
    int rc = jmb_reserve_dev(ctx, &ctx->d_stage3, &ctx->d_stage3_cap, lb + 2 * cbytes + mb_bytes + 256); if (rc) return rc;
    uint8_t *d = (uint8_t *)ctx->d_stage3;
    A.luma = d; A.cb = d + lb; A.cr = d + lb + cbytes; A.pitch = dp; A.pitch_c = dpc; A.mbs = (const jmb_db_mb *)(d + lb + 2 * cbytes);
    JMB_CUDA(ctx, cudaMemcpy2DAsync(A.luma, dp, luma, pitch, width, height, cudaMemcpyHostToDevice, ctx->stream));
    if (yuv_format) {
      JMB_CUDA(ctx, cudaMemcpy2DAsync(A.cb, dpc, cb, pitch_c, wc, hc, cudaMemcpyHostToDevice, ctx->stream));
      JMB_CUDA(ctx, cudaMemcpy2DAsync(A.cr, dpc, cr, pitch_c, wc, hc, cudaMemcpyHostToDevice, ctx->stream));
    }
    JMB_CUDA(ctx, cudaMemcpyAsync((void *)A.mbs, mbs, mb_bytes, cudaMemcpyHostToDevice, ctx->stream));
  } else if ((pitch & 3) || (pitch_c & 3) || ((size_t)luma & 3) || ((size_t)cb & 3) || ((size_t)cr & 3))
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_deblock_picture: device planes must be 4-byte aligned with pitches that are multiples of 4");
  // ticket + completion flags
  if ((size_t)(n + 1) * sizeof(int) > ctx->d_db_cap) {
    int rc = jmb_reserve_dev(ctx, &ctx->d_db, &ctx->d_db_cap, (size_t)(n + 1) * sizeof(int)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemsetAsync(ctx->d_db, 0, ctx->d_db_cap, ctx->stream));
    ctx->db_serial = 0;
  }
  JMB_CUDA(ctx, cudaMemsetAsync(ctx->d_db, 0, sizeof(int), ctx->stream));
  A.ticket = (unsigned *)ctx->d_db; A.done = (int *)ctx->d_db + 1; A.serial = ++ctx->db_serial;
  A.mbw = mbw; A.mbh = mbh; A.yuv = yuv_format; A.slice_type = slice_type; A.d8 = direct_8x8_inference != 0;
  jmb_time_begin(ctx, JMB_K_DEBLOCK);
  k_deblock<<<n, 64, 0, ctx->stream>>>(A);
  jmb_time_end(ctx, JMB_K_DEBLOCK);
  JMB_LAUNCH_CHECK(ctx);
  if (host) {
    JMB_CUDA(ctx, cudaMemcpy2DAsync(luma, pitch, A.luma, A.pitch, width, height, cudaMemcpyDeviceToHost, ctx->stream));
    if (yuv_format) {
      JMB_CUDA(ctx, cudaMemcpy2DAsync(cb, pitch_c, A.cb, A.pitch_c, wc, hc, cudaMemcpyDeviceToHost, ctx->stream));
      JMB_CUDA(ctx, cudaMemcpy2DAsync(cr, pitch_c, A.cr, A.pitch_c, wc, hc, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (loc == JMB_HOST) return jmb_check_device_errors(ctx);
  }
  return JMB_OK;
}
