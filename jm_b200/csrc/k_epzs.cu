// k_epzs.cu -- EPZS (SearchMode 3) on the device: EPZS_integer_motion_estimation (lencod/src/me_epzs_int.c:42-426) and
// EPZS_sub_pel_motion_estimation (lencod/src/me_epzs_sub.c:30-213), one CTA per 41 searches (a macroblock's partitions).
//
// EPZS is a short, data-dependent walk: check the start mv, stop early against the previous distortions, check an ordered
// predictor list, walk a refinement pattern until its centre wins (optionally again from the second-best predictor), then a
// two-step half-/quarter-pel pattern.  10-60 distortions per block instead of the 4225 of the full search.  What is
// data-parallel in it are the distortions of one step -- and the 41 searches of a macroblock, which are independent once
// their predictors are given: the CTA advances all of them round by round (see k_epzs_int).  The sub-pel stage is a second launch.
#include "jmb_dist_dev.cuh"

namespace {

constexpr int EG = 41;         // searches per CTA: consecutive requests (one macroblock's 41 partitions in the picture form)
constexpr int ET = 128;        // threads per CTA
constexpr int CB = 48;         // candidates a search puts up per round (a multiple of 4: IntS stays word-sized)
#ifndef JMB_EPZS_SUB_MINB
#define JMB_EPZS_SUB_MINB 8
#endif
#ifndef JMB_EPZS_INT_MINB
#define JMB_EPZS_INT_MINB 6
#endif

// pattern_data (me_epzs_common.c:48-76): {mv_x, mv_y, start_nmbr, next_points}, quarter-pel; chaining as EPZSInit (:178-230):
// every pattern stops on itself except sbdiamond / pmvfast, which hand over to the small diamond; nextLast is TRUE for all
__constant__ short c_pat[6][12][4] = {
  {{0, 4, 3, 3}, {4, 0, 0, 3}, {0, -4, 1, 3}, {-4, 0, 2, 3}},
  {{0, 4, 7, 3}, {4, 4, 7, 5}, {4, 0, 1, 3}, {4, -4, 1, 5}, {0, -4, 3, 3}, {-4, -4, 3, 5}, {-4, 0, 5, 3}, {-4, 4, 5, 5}},
  {{-4, 4, 10, 5}, {0, 8, 10, 8}, {0, 4, 10, 7}, {4, 4, 1, 5}, {8, 0, 1, 8}, {4, 0, 1, 7}, {4, -4, 4, 5}, {0, -8, 4, 8},
   {0, -4, 4, 7}, {-4, -4, 7, 5}, {-8, 0, 7, 8}, {-4, 0, 7, 7}},
  {{0, 8, 6, 5}, {4, 4, 0, 3}, {8, 0, 0, 5}, {4, -4, 2, 3}, {0, -8, 2, 5}, {-4, -4, 4, 3}, {-8, 0, 4, 5}, {-4, 4, 6, 3}},
  {{0, 8, 6, 12}, {4, 4, 0, 12}, {8, 0, 0, 12}, {4, -4, 2, 12}, {0, -8, 2, 12}, {-4, -4, 4, 12}, {-8, 0, 4, 12}, {-4, 4, 6, 12},
   {0, 2, 6, 12}, {2, 0, 0, 12}, {0, -2, 2, 12}, {-2, 0, 4, 12}},
  {{0, 8, 6, 5}, {4, 4, 0, 3}, {8, 0, 0, 5}, {4, -4, 2, 3}, {0, -8, 2, 5}, {-4, -4, 4, 3}, {-8, 0, 4, 5}, {-4, 4, 6, 3}}};
__constant__ unsigned char c_pat_n[6] = {4, 8, 12, 8, 12, 8}, c_pat_stop[6] = {1, 1, 1, 1, 0, 0}, c_pat_next[6] = {0, 1, 2, 3, 0, 0};
// search_point_hp / next_start_pos / next_end_pos, lencod/inc/me_epzs.h:23-42 (search_point_qp = half of these)
__constant__ signed char c_hp[10][2] = {{0, 0}, {-2, 0}, {0, 2}, {2, 0}, {0, -2}, {-2, 2}, {2, 2}, {2, -2}, {-2, -2}, {-2, 2}};
__constant__ unsigned char c_ns[5][5] = {{0, 8, 5, 6, 7}, {8, 0, 5, 8, 8}, {5, 5, 0, 6, 5}, {6, 6, 6, 0, 7}, {7, 8, 7, 7, 0}};
__constant__ unsigned char c_ne[5][5] = {{0, 10, 7, 8, 9}, {10, 0, 6, 10, 9}, {7, 6, 0, 7, 7}, {8, 8, 7, 0, 8}, {9, 9, 9, 8, 0}};
// window predictor ring, EPZSWindowPredictorInit mode 0 (me_epzs_common.c:352-371): i = +1 then -1
__constant__ signed char c_ring[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};


__device__ __forceinline__ unsigned mv_cost32(int lam, int vx, int vy, int px, int py) {      // lambda <= 65535, mvbits <= 33 each: < 2^23
  return (unsigned)lam * (unsigned)(jmb_mvbits(vx - px) + jmb_mvbits(vy - py));
}

enum { EPZS_ERR_FIELD = 512 };      // a request field out of range (reported through d_err like the JMB_REQERR_* codes)

__device__ __forceinline__ int epzs_check(const jmb_epzs_req &q, int w, int h, int nref, int n_cands, int max_range) {
  if (q.blocktype < 1 || q.blocktype > 7) return JMB_REQERR_BLOCKTYPE;
  const int bsx = c_bsx[q.blocktype], bsy = c_bsy[q.blocktype];
  int e = 0;
  if (q.ref >= nref) e |= JMB_REQERR_REF;
  if (q.pos_x < 0 || q.pos_y < 0 || q.pos_x + bsx > w || q.pos_y + bsy > h || (q.pos_x % bsx) || (q.pos_y % bsy)) e |= JMB_REQERR_POS;
  if ((unsigned)q.lambda[0] > 65535u || (unsigned)q.lambda[1] > 65535u || (unsigned)q.lambda[2] > 65535u) e |= JMB_REQERR_LAMBDA;
  const long long lim = 1ll << 48;
  if (q.stop < 0 || q.stop > lim || q.medthres < 0 || q.medthres > lim || q.prev_sad < 0 || q.prev_sad > lim || q.subthres < 0 ||
      q.subthres > lim || q.min_mcost < 0 || q.min_mcost > lim) e |= JMB_REQERR_MINCOST;
  const int tot = q.n_cand[0] + q.n_cand[1] + ((q.flags & JMB_EPZS_WINDOW_GEN) ? 0 : q.n_cand[2]) + q.n_cand[3];
  if (q.pattern > 5 || q.pattern_dual > 5 || q.range_x < 1 || q.range_y < 1 || q.range_x > max_range || q.range_y > max_range ||
      q.cand_off < 0 || q.cand_off + tot > n_cands || ((q.flags & JMB_EPZS_TEST8X8) && q.blocktype > 4) ||
      ((q.flags & JMB_EPZS_WINDOW_GEN) && q.n_cand[2] > 63)) e |= EPZS_ERR_FIELD;
  return e;
}

// ---- integer stage: EPZS_integer_motion_estimation (me_epzs_int.c:42-426) ---------------------------------------------
// One CTA walks EG searches TOGETHER, round by round.  Thread s < EG directs search s: it replays JM's control flow (early
// exits, predictor list, pattern walk, dual refinement) on COMPLETE costs and puts up the candidates of its next step; then
// all 128 threads evaluate everything that was put up -- one 4x4 block of one search per thread (the blocks of a macroblock's
// 41 partitions are 112 of them), walking that search's candidates and adding mv cost + (SAD << 5) into tot[search][candidate]
// -- and the directors read the totals.  JM's early-terminated distortion returns the bound it was given (mv_search.h:19-23),
// which can never win a strict '<', so complete sums decide alike.
// JM's visited map (EPZSMap) is not needed.  A position met again costs what it cost before, which is at least the current
// minimum (the minimum only falls), and only a strict '<' moves the minimum: the pattern walk decides alike with or without
// the map.  In the predictor list a repeated position could do one thing only -- enter the best / second-best pair a second
// time, as the second best behind itself -- and that is the one case the merge below excludes (a candidate equal to the
// current best is passed over; one that was second best, or neither, fails both '<' on its own).  Repeats cost a few
// redundant distortions (n_evals counts them) and nothing else.
struct IntS {
  short pos_x, pos_y, px, py;
  int lam, first;                 // first: index of the search's first 4x4 block among the CTA's
  unsigned char blocktype, ref, nc, lbx;   // nc: candidates put up this round; lbx: log2(4x4 blocks per row)
  short2 cand[CB];
  unsigned char cidx[CB];         // pattern rounds: which point of the round
};

__global__ void __launch_bounds__(ET, JMB_EPZS_INT_MINB)
k_epzs_int(const jmb_epzs_req *__restrict__ reqs, int n, const short2 *__restrict__ cands, int n_cands, jmb_epzs_res *__restrict__ res,
           const uint8_t *__restrict__ cur, int cur_pitch, const uint8_t *const *__restrict__ ref_planes, size_t plane_bytes, int ref_pitch,
           int w, int h, int nref, int max_range, int *__restrict__ err) {
  __shared__ IntS S[EG];
  __shared__ int tot[EG][CB + 1];      // (+1: the directors read their rows side by side)
  __shared__ int s_first[EG + 1];
  const int tid = threadIdx.x, base = blockIdx.x * EG, cnt = min(EG, n - base);
  const long long BIG = (long long)0x7fffffff << 5;      // DISTBLK_MAX
  enum { PH_START, PH_PRED, PH_PAT, PH_DONE };
  int phase = PH_DONE;
  // director state
  int sx = 0, sy = 0, rx = 0, ry = 0, tx = 0, ty = 0, t2x = 0, t2y = 0, check_median = 0, exit_code = 0, evals = 0, seg = 0, i0 = 0, off = 0;
  int pat = 0, cx = 0, cy = 0, point = 0, total = 0, next_last = 0, pattern_stop = 0, dir = 0, pending = 0;
  unsigned flags = 0, ncw = 0, gtw = 0, jm_ref = 0, pats = 0;
  long long minc = 0, second = 0, prev = 0, med = 0, stop = 0, ld = 0, centre_cost = 0;
  if (tid < cnt) {
    const jmb_epzs_req q = reqs[base + tid];
    IntS &me_ = S[tid];
    me_.nc = 0; me_.first = 0; me_.lbx = 0; me_.blocktype = 7;
    const int bad = epzs_check(q, w, h, nref, n_cands, max_range);
    jmb_epzs_res o;
    o.mv_x = o.imv_x = q.start_x; o.mv_y = o.imv_y = q.start_y; o.cost = o.icost = q.min_mcost; o.prev_sad = q.prev_sad; o.exit_code = 0; o.n_evals = 0;
    if (bad) { jmb_req_report(err, bad, base + tid); res[base + tid] = o; }
    else if (q.flags & JMB_EPZS_SKIP_INT) res[base + tid] = o;      // sub-pel only: the integer-stage fields of the result echo the request
    else {
      me_.pos_x = q.pos_x; me_.pos_y = q.pos_y; me_.px = q.pred_x; me_.py = q.pred_y; me_.lam = q.lambda[0];
      me_.blocktype = q.blocktype; me_.ref = q.ref; me_.lbx = (unsigned char)(c_bsx[q.blocktype] >> 3);      // 4 -> 0, 8 -> 1, 16 -> 2
      sx = q.start_x; sy = q.start_y; rx = q.range_x; ry = q.range_y; tx = sx; ty = sy;
      flags = q.flags; jm_ref = q.jm_ref; pats = q.pattern | (q.pattern_dual << 8);
      ncw = q.n_cand[0] | (q.n_cand[1] << 8) | (q.n_cand[2] << 16) | ((unsigned)q.n_cand[3] << 24);
      gtw = q.gate[0] | (q.gate[1] << 8) | (q.gate[2] << 16) | ((unsigned)q.gate[3] << 24);
      off = q.cand_off; prev = q.prev_sad; med = q.medthres; stop = q.stop; ld = 2ll * q.lambda[0];
      phase = PH_START;
      me_.cand[0] = make_short2((short)sx, (short)sy); me_.nc = 1; tot[tid][0] = 0;      // the start mv (:93-100)
    }
  }
  __syncthreads();
  if (tid == 0) {      // the CTA's 4x4 blocks, search after search (searches that take no part own none)
    int f = 0;
    for (int i = 0; i < cnt; i++) { s_first[i] = f; if (S[i].nc) f += (c_bsx[S[i].blocktype] >> 2) * (c_bsy[S[i].blocktype] >> 2); }
    s_first[cnt] = f;
  }
  __syncthreads();
  const int nblk = s_first[cnt];
  auto in_range = [&](int vx, int vy) { return abs(vx - sx) <= rx && abs(vy - sy) <= ry; };
  // the search a 4x4 block belongs to: the last one whose first block is <= item (searches without blocks share a first index with
  // the next one that has some).  The thread's first block is the same in every round (a macroblock has 112 of them).
  auto owner_of = [&](int item) { int lo = 0, hi = cnt - 1; while (lo < hi) { const int m = (lo + hi + 1) >> 1; if (s_first[m] <= item) lo = m; else hi = m - 1; } return lo; };
  const int lo0 = tid < nblk ? owner_of(tid) : 0;

  for (;;) {
    if (!__syncthreads_or(phase != PH_DONE)) break;
    // ---- everybody: the distortions of what was put up ----
    for (int item = tid; item < nblk; item += ET) {
      const int lo = item == tid ? lo0 : owner_of(item);
      const IntS &q = S[lo];
      const int nc = q.nc;
      if (!nc) continue;
      const int blk = item - s_first[lo], bx = blk & ((1 << q.lbx) - 1), by = blk >> q.lbx;
      const uint8_t *sp = cur + (size_t)(q.pos_y + by * 4) * cur_pitch + q.pos_x + bx * 4;
      const unsigned s0 = *(const unsigned *)sp, s1 = *(const unsigned *)(sp + cur_pitch), s2 = *(const unsigned *)(sp + 2 * (size_t)cur_pitch),
                     s3 = *(const unsigned *)(sp + 3 * (size_t)cur_pitch);
      const RefView rv{ref_planes[q.ref], plane_bytes, ref_pitch, w, h};
      const int bqx = q.pos_x << 2, bqy = q.pos_y << 2;
      const size_t boff = (size_t)(by * 4) * ref_pitch + bx * 4;
      for (int c = 0; c < nc; c++) {
        const short2 v = q.cand[c];
        const uint8_t *rp = umv(rv, bqy + v.y, bqx + v.x) + boff;      // computeSAD (me_distortion.c:349): the PARTITION origin is clamped
        const RowsAt R = rows_at(rp, ref_pitch);
        unsigned d = (__vsadu4(s0, row4(R, 0)) + __vsadu4(s1, row4(R, 1)) + __vsadu4(s2, row4(R, 2)) + __vsadu4(s3, row4(R, 3))) << 5;
        if (blk == 0) d += mv_cost32(q.lam, v.x, v.y, q.px, q.py);
        atomicAdd(&tot[lo][c], (int)d);
      }
    }
    __syncthreads();
    // ---- directors: read the totals, decide, put up the next candidates ----
    if (phase != PH_DONE) {
      IntS &me_ = S[tid];
      int have = me_.nc;      // totals of `have` candidates are in tot[tid][]
      me_.nc = 0;
      bool posted = false;
      while (!posted && phase != PH_DONE) {
        if (phase == PH_START) {
          evals++;
          minc = (unsigned)tot[tid][0];
          const bool gt0 = (flags & JMB_EPZS_REF_GT0_FRAME) != 0;
          if (gt0 && (prev < min(med + ld, minc) || prev * 8 < minc)) { exit_code = 1; phase = PH_DONE; }               // :103-117
          else if (!(minc > med + ld)) phase = PH_DONE;                                                                 // :121
          else if (minc < (stop >> 1)) { if (jm_ref == 0 || prev > minc) prev = minc; exit_code = 2; phase = PH_DONE; } // :135-150
          else {
            second = BIG; centre_cost = minc;      // JM runs the predictor generators (and their gates) before it checks any predictor
            seg = 0; i0 = 0; have = 0; phase = PH_PRED;
          }
        } else if (phase == PH_PRED) {
          // JM's rule (:226-251) keeps the two cheapest candidates so far, the earlier one first on a tie
          for (int c = 0; c < have; c++) {
            const long long v = (unsigned)tot[tid][c];
            const short2 m = me_.cand[c];
            if (m.x == tx && m.y == ty) continue;      // the current best named again (see above)
            if (v < minc) { second = minc; t2x = tx; t2y = ty; minc = v; tx = m.x; ty = m.y; check_median = 1; }
            else if (v < second) { second = v; t2x = m.x; t2y = m.y; check_median = 1; }
          }
          have = 0;
          int k = 0;
          while (seg < 4 && k < CB) {      // the next stretch of the list (:215-252), across its segments: what is in range
            const int ns = (ncw >> (8 * seg)) & 255, gate = (gtw >> (8 * seg)) & 255;
            const bool gen = seg == 2 && (flags & JMB_EPZS_WINDOW_GEN);
            if (i0 >= ns || !(gate == 0 || centre_cost > (long long)gate * stop)) { if (!gen) off += ns; seg++; i0 = 0; continue; }
            const int nb = min(CB - k, ns - i0);
            for (int i = i0; i < i0 + nb; i++) {
              short2 v;
              if (gen) {      // window predictors around the start mv: rings of size range >> k, k descending (EPZSWindowPredictorInit)
                const int rings = (ns + 8) >> 3, spc = rx >> (rings - 1 - (i >> 3));
                v = make_short2((short)(sx + c_ring[i & 7][0] * spc), (short)(sy + c_ring[i & 7][1] * spc));
              } else v = cands[off + i];
              if (in_range(v.x, v.y)) me_.cand[k++] = v;
            }
            i0 += nb;
          }
          if (k) { for (int c = 0; c < k; c++) tot[tid][c] = 0; me_.nc = (unsigned char)k; evals += k; posted = true; }
          else {      // the list is through
            const bool gt0 = (flags & JMB_EPZS_REF_GT0_FRAME) != 0;
            if (gt0 && prev * 3 < minc) { exit_code = 3; phase = PH_DONE; }                                             // :254-273
            else if (!(minc > stop)) phase = PH_DONE;                                                                   // :279
            else {
              pat = pats & 255;
              if (flags & JMB_EPZS_ADAPT_PATTERN) {                                                                     // :286-300
                if (minc < stop + ((3 * med) >> 1))
                  pat = ((tx == 0 && ty == 0) || (abs(tx - sx) < 10 && abs(ty - sy) < 10)) ? JMB_EPZS_PAT_SDIAMOND : JMB_EPZS_PAT_SQUARE;
                else if (flags & JMB_EPZS_SQUARE_HINT) pat = JMB_EPZS_PAT_SQUARE;
              }
              cx = tx; cy = ty;
              pattern_stop = 0; point = 0; next_last = 0; total = c_pat_n[pat]; dir = 0; pending = 0;
              phase = PH_PAT;
            }
          }
        } else {      // PH_PAT: the pattern walk (:307-360)
          if (pending) {
            const int np = c_pat_n[pat];
            long long bestv = minc; int bc = -1;
            for (int c = 0; c < have; c++) { const long long v = (unsigned)tot[tid][c]; if (v < bestv) { bestv = v; bc = c; } }      // the first of the cheapest
            if (bc >= 0) { tx = me_.cand[bc].x; ty = me_.cand[bc].y; minc = bestv; dir = point + me_.cidx[bc]; if (dir >= np) dir -= np; }
            have = 0; pending = 0;
            if (next_last || (tx == cx && ty == cy)) {
              pattern_stop = c_pat_stop[pat];
              pat = c_pat_next[pat];
              total = c_pat_n[pat];
              next_last = 1; dir = 0; point = 0;
            } else {
              total = c_pat[pat][dir][3];
              point = c_pat[pat][dir][2];
              cx = tx; cy = ty;
            }
            if (pattern_stop == 1) {
              const bool gt0 = (flags & JMB_EPZS_REF_GT0_FRAME) != 0;
              if (gt0 && (4 * prev < minc || (3 * prev < minc && prev <= stop))) { exit_code = 4; phase = PH_DONE; continue; }      // :362-376
              if (!(check_median && (jm_ref == 0 || minc < 2 * prev) && minc > ((3 * stop) >> 1) && (flags & JMB_EPZS_DUAL))) { phase = PH_DONE; continue; }   // :379-384
              if ((tx == 0 && ty == 0) || (tx == sx && ty == sy))                                                        // :391-399
                pat = (abs(tx - sx) < 10 && abs(ty - sy) < 10) ? JMB_EPZS_PAT_SDIAMOND : JMB_EPZS_PAT_SQUARE;
              else pat = (pats >> 8) & 255;
              cx = t2x; cy = t2y;
              check_median = 0;
              pattern_stop = 0; point = 0; next_last = 0; total = c_pat_n[pat]; dir = 0;
            }
          }
          const int np = c_pat_n[pat];
          int k = 0;
          for (int l = 0; l < total; l++) {
            int pi = point + l; if (pi >= np) pi -= np;
            const int vx = cx + c_pat[pat][pi][0], vy = cy + c_pat[pat][pi][1];
            if (in_range(vx, vy)) { me_.cand[k] = make_short2((short)vx, (short)vy); me_.cidx[k] = (unsigned char)l; k++; }
          }
          pending = 1;
          if (k) { for (int c = 0; c < k; c++) tot[tid][c] = 0; me_.nc = (unsigned char)k; evals += k; posted = true; }
        }
      }
      if (phase == PH_DONE) {
        if (!exit_code) { if (jm_ref == 0 || prev > minc) prev = minc; exit_code = 5; }                                  // :409-410
        jmb_epzs_res o;
        o.mv_x = o.imv_x = (int16_t)tx; o.mv_y = o.imv_y = (int16_t)ty;
        o.cost = o.icost = minc; o.prev_sad = prev; o.exit_code = exit_code; o.n_evals = evals;
        res[base + tid] = o;
      }
    }
  }
}

// ---- sub-pel stage: BlockMotionSearch's gate (mv_search.c:964-976), then EPZS_sub_pel_motion_estimation (me_epzs_sub.c:30-213).
// Same shape: EG searches per CTA, thread s directs search s through the four steps (half-pel diamond, its refinement towards
// the second best, the same two at quarter-pel), all threads evaluate -- one 4x4 (or, with the 8x8 Hadamard, 8x8) sub-block of
// one search per work item, walking that search's candidates of the step.
struct SubS {
  short pos_x, pos_y;
  int first;
  unsigned char blocktype, ref, t8, nc, nsub, nn, nsx, lsx;      // nsub = sub-blocks (a power of two), nsx per row = 1 << lsx
  short nitems; unsigned char lsub, pad_;
  short2 cand[5];
};

__global__ void __launch_bounds__(ET, JMB_EPZS_SUB_MINB)
k_epzs_sub(const jmb_epzs_req *__restrict__ reqs, int n, jmb_epzs_res *__restrict__ res, const uint8_t *__restrict__ cur, int cur_pitch,
           const uint8_t *const *__restrict__ ref_planes, size_t plane_bytes, int ref_pitch, int w, int h, jmb_me_config me, int nref) {
  __shared__ SubS S[EG];
  __shared__ int sums[EG][5];
  __shared__ int s_pre[ET / 32][EG + 24];
  const int tid = threadIdx.x, base = blockIdx.x * EG, cnt = min(EG, n - base);
  const long long BIG = (long long)0x7fffffff << 5;
  bool live = false, done = false;
  int px = 0, py = 0, mvx = 0, mvy = 0, evals = 0, lam = 0, lam_q = 0, best = 0, second_pos = 0, p0 = 0;
  long long minc = 0, second = BIG, sub_thr = 0;
  if (tid < cnt) {
    const jmb_epzs_req q = reqs[base + tid];
    SubS &me_ = S[tid];
    me_.nc = 0; me_.nsub = 0; me_.nitems = 0; me_.blocktype = 7;
    if ((q.flags & JMB_EPZS_SUBPEL) && q.blocktype >= 1 && q.blocktype <= 7 && q.ref < nref) {      // (rejected requests were reported by the integer stage)
      const jmb_epzs_res r0 = res[base + tid];
      const bool gt0 = (q.flags & JMB_EPZS_REF_GT0_FRAME) != 0, skip = (q.flags & JMB_EPZS_SKIP_INT) != 0;
      if ((r0.exit_code != 0 || skip) && (skip || !gt0 || 2 * r0.icost < 7 * r0.prev_sad)) {      // mv_search.c:964-976
        live = true;
        me_.pos_x = q.pos_x; me_.pos_y = q.pos_y; me_.blocktype = q.blocktype; me_.ref = q.ref; me_.t8 = (q.flags & JMB_EPZS_TEST8X8) != 0;
        px = q.pred_x; py = q.pred_y; mvx = r0.imv_x; mvy = r0.imv_y; evals = r0.n_evals; minc = r0.icost;
        lam = q.lambda[1]; lam_q = q.lambda[2];
        sub_thr = q.subthres + 2ll * lam;
        if (!skip && !me.start_hp) minc = BIG;
      }
    }
  }
  const int max_pos2 = (!me.start_hp || !me.start_qp) ? max(1, me.search_pos2) : me.search_pos2;

#pragma unroll 1
  for (int st = 0; st < 4; st++) {
    const int metric = me.metric[1 + (st >> 1)], div = 1 + (st >> 1);
    const bool wide = !(st & 1);
    if (tid < cnt) {      // directors: the candidates of this step
      SubS &me_ = S[tid];
      int p1 = 0;
      p0 = 0;
      if (live && !done) {
        if (st == 0) { p0 = me.start_hp; p1 = min(5, max_pos2); }                                                       // me_epzs_sub.c:66-90
        else if (st == 1) { if (me.search_pos2 >= 9 && (best != 0 || (abs(px - mvx) + abs(py - mvy)))) { p0 = c_ns[best][second_pos]; p1 = c_ne[best][second_pos]; } }   // :97-122
        else if (st == 2) { p0 = me.start_qp; p1 = (minc < sub_thr) ? 1 : 5; }                                          // :135-170
        else if (minc > sub_thr && (best != 0 || (abs(px - mvx) + abs(py - mvy)))) { p0 = c_ns[best][second_pos]; p1 = c_ne[best][second_pos]; }               // :173-200
      }
      const int nb = max(0, p1 - p0);
      for (int c = 0; c < nb; c++) { me_.cand[c] = make_short2((short)(mvx + c_hp[p0 + c][0] / div), (short)(mvy + c_hp[p0 + c][1] / div)); sums[tid][c] = 0; }
      me_.nc = (unsigned char)nb;
      const int nn = (metric == JMB_SATD && me_.t8) ? 8 : 4;
      me_.nn = (unsigned char)nn; me_.nsx = (unsigned char)(c_bsx[me_.blocktype] / nn);
      me_.nsub = nb ? (unsigned char)(me_.nsx * (c_bsy[me_.blocktype] / nn)) : 0;
      me_.nitems = (short)(nn == 8 ? me_.nsub * nb : me_.nsub);
      me_.lsx = (unsigned char)(me_.nsx >> 1); me_.lsub = (unsigned char)(31 - __clz(max(1, (int)me_.nsub)));      // nsx = 1, 2, 4
    }
    __syncthreads();
    // work items: a 4x4 sub-block walks its search's candidates; an 8x8 one (a Hadamard of 64 differences) takes one candidate per
    // item.  Every warp adds up the searches' item counts for itself (two shuffle scans; a shared prefix array would cost a barrier
    // more) and an item finds its search by bisection.
    int *const pre = s_pre[tid >> 5];
    {
      const int lane = tid & 31;
      int a = lane < cnt ? S[lane].nitems : 0, b = lane + 32 < cnt ? S[lane + 32].nitems : 0;
#pragma unroll
      for (int sh = 1; sh < 32; sh <<= 1) { const int u = __shfl_up_sync(0xffffffffu, a, sh); if (lane >= sh) a += u; }
      const int t32 = __shfl_sync(0xffffffffu, a, 31);
#pragma unroll
      for (int sh = 1; sh < 16; sh <<= 1) { const int u = __shfl_up_sync(0xffffffffu, b, sh); if (lane >= sh) b += u; }
      pre[lane + 1] = a;
      if (lane + 33 <= EG) pre[lane + 33] = t32 + b;
      if (lane == 0) pre[0] = 0;
      __syncwarp();
    }
    const int nitems = pre[cnt];
    for (int item = tid; item < nitems; item += ET) {
      int lo = 0, hi = cnt - 1;                        // last search whose first item is <= item (searches without items share it with the next that has some)
      while (lo < hi) { const int m = (lo + hi + 1) >> 1; if (pre[m] <= item) lo = m; else hi = m - 1; }
      const SubS &q = S[lo];
      const int nn = q.nn, nc = q.nc, it = item - pre[lo], c8 = nn == 8 ? it >> q.lsub : 0, sb = nn == 8 ? it & (q.nsub - 1) : it;
      const int sbx = sb & (q.nsx - 1), sby = sb >> q.lsx;
      const RefView rv{ref_planes[q.ref], plane_bytes, ref_pitch, w, h};
      SrcBlk src;
      load_src(src, cur, cur_pitch, q.pos_x + sbx * nn, q.pos_y + sby * nn, nn);
      const int bqx = q.pos_x << 2, bqy = q.pos_y << 2;
      if (nn == 4) {
        for (int c = 0; c < nc; c++) {
          unsigned rw[4];
          load_ref4(rv, bqx + q.cand[c].x, bqy + q.cand[c].y, sbx, sby, metric, rw);
          atomicAdd(&sums[lo][c], dist4(src, rw, metric));
        }
      } else atomicAdd(&sums[lo][c8], subblock_dist(rv, src, bqx + q.cand[c8].x, bqy + q.cand[c8].y, sbx, sby, 8, metric));
    }
    __syncthreads();
    if (tid < cnt && live && !done) {      // directors: JM's sequential selection on the complete distortions
      const SubS &me_ = S[tid];
      const int nb = me_.nc;
      evals += nb;
      for (int c = 0; c < nb; c++) {
        const int pos = p0 + c;
        long long mcost = mv_cost32(lam, me_.cand[c].x, me_.cand[c].y, px, py);
        if (wide) {
          if (mcost < second) {
            mcost += (long long)sums[tid][c] << 5;
            if (mcost < minc) { second = minc; second_pos = best; minc = mcost; best = pos; }
            else if (mcost < second) { second = mcost; second_pos = pos; }
          }
        } else if (mcost < minc) {
          mcost += (long long)sums[tid][c] << 5;
          if (mcost < minc) { minc = mcost; best = pos; }
        }
      }
      if (st == 0 && best == 0 && px == mvx && py == mvy && minc < sub_thr) done = true;                               // :92-95
      if (st == 1 && !done) {
        if (best) { mvx += c_hp[best][0]; mvy += c_hp[best][1]; }
        second = BIG; best = 0;      // start_me_refinement_qp == 1 (checked by the host side); second_pos carries over as in JM
        lam = lam_q;
      }
      if (st == 3 && best > 0) { mvx += c_hp[best][0] / 2; mvy += c_hp[best][1] / 2; }
    }
  }
  if (tid < cnt && live) { res[base + tid].mv_x = (int16_t)mvx; res[base + tid].mv_y = (int16_t)mvy; res[base + tid].cost = minc; res[base + tid].n_evals = evals; }
}

// requests of a whole picture for jmb_epzs_search_frame (88-byte records staged in shared memory, written as 16-byte words)
__global__ void __launch_bounds__(256)
k_gen_epzs(const jmb_mb_mvpred *__restrict__ pred, int n_mb, int mb_w, jmb_epzs_frame_params fp, jmb_epzs_req *__restrict__ reqs) {
  __shared__ __align__(16) jmb_epzs_req sq[256];
  static_assert(sizeof(jmb_epzs_req) == 88, "staging copies 256 x 88 bytes as 1408 x 16");
  const int t0 = blockIdx.x * 256, t = t0 + threadIdx.x, n = n_mb * 41;
  if (t < n) {
    const int mb = t / 41, p = t - mb * 41;
    // canonical partition order: type, then raster order inside the macroblock
    const int type = p < 1 ? 1 : p < 3 ? 2 : p < 5 ? 3 : p < 9 ? 4 : p < 17 ? 5 : p < 25 ? 6 : 7;
    const int first = type == 1 ? 0 : type == 2 ? 1 : type == 3 ? 3 : type == 4 ? 5 : type == 5 ? 9 : type == 6 ? 17 : 25;
    const int bsx = c_bsx[type], bsy = c_bsy[type], k = p - first, per_row = 16 / bsx;
    jmb_epzs_req q;
    memset(&q, 0, sizeof(q));
    q.pos_x = (int16_t)((mb % mb_w) * 16 + (k % per_row) * bsx); q.pos_y = (int16_t)((mb / mb_w) * 16 + (k / per_row) * bsy);
    const int px = pred[mb].pred[p][0], py = pred[mb].pred[p][1];
    q.pred_x = (int16_t)px; q.pred_y = (int16_t)py;
    // EPZSSubPelGrid: the search starts at the predictor itself (mv_search.c:925-928), clipped to the mv range (:957)
    q.start_x = (int16_t)jmb_clip(fp.mv_min_x, fp.mv_max_x, px); q.start_y = (int16_t)jmb_clip(fp.mv_min_y, fp.mv_max_y, py);
    q.blocktype = (uint8_t)type; q.ref = q.jm_ref = (uint8_t)fp.ref;
    q.flags = (uint8_t)((fp.flags & (JMB_EPZS_ADAPT_PATTERN | JMB_EPZS_DUAL | JMB_EPZS_SUBPEL | (type <= 4 ? JMB_EPZS_TEST8X8 : 0))) | (fp.window ? JMB_EPZS_WINDOW_GEN : 0));
    q.pattern = (uint8_t)fp.pattern; q.pattern_dual = (uint8_t)fp.pattern_dual;
    q.n_cand[0] = (uint8_t)fp.n_shared; q.n_cand[2] = (uint8_t)(fp.window ? 8 * fp.window - 1 : 0);
    q.gate[2] = 3;                                                       // me_epzs_int.c:193-198
    q.cand_off = mb * fp.n_shared;
    q.lambda[0] = fp.lambda[0]; q.lambda[1] = fp.lambda[1]; q.lambda[2] = fp.lambda[2];
    q.range_x = q.range_y = (int16_t)fp.range;
    // EPZSDetermineStopCriterion (me_epzs_common.c:1874) with no neighbour distortion known (sadA = sadB = sadC = DISTBLK_MAX)
    const long long ld = 2ll * fp.lambda[0], med = fp.medthres[type];
    long long stop = (long long)0x7fffffff << 5;
    stop = max(stop, (long long)fp.minthres[type]);
    stop = min(stop, (long long)fp.maxthres[type] + ld);
    stop = (8 * max(med + ld, stop) + med) >> 3;
    q.stop = stop + ld; q.medthres = med; q.subthres = fp.subthres[type];
    q.prev_sad = (long long)0x7fffffff << 5;
    q.min_mcost = (long long)0x7fffffff << 5;
    sq[threadIdx.x] = q;
  }
  __syncthreads();
  const int cnt = min(256, n - t0);
  uint4 *dst = (uint4 *)(reqs + t0);
  for (int i = threadIdx.x; i < cnt * 88 / 16; i += 256) dst[i] = ((const uint4 *)sq)[i];
  for (int i = (cnt * 88 / 16) * 16 + threadIdx.x * 8; i < cnt * 88; i += 256 * 8) *(uint2 *)((char *)dst + i) = *(const uint2 *)((const char *)sq + i);
}

// 8-byte results (+ the final clip of the mv, mv_search.c:981) and the 24-byte form the residual coder reads
__global__ void k_epzs_pack(const jmb_epzs_res *__restrict__ res, int n, jmb_epzs_frame_params fp, jmb_me_res *__restrict__ keep, jmb_me_res8 *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const jmb_epzs_res r = res[t];
  const int mx = jmb_clip(fp.mv_min_x, fp.mv_max_x, r.mv_x), my = jmb_clip(fp.mv_min_y, fp.mv_max_y, r.mv_y);
  jmb_me_res k;
  k.mv_x = (int16_t)mx; k.mv_y = (int16_t)my; k.imv_x = r.imv_x; k.imv_y = r.imv_y; k.cost = r.cost; k.icost = r.icost;
  keep[t] = k;
  if (out) {
    jmb_me_res8 o;
    o.mv_x = (int16_t)mx; o.mv_y = (int16_t)my; o.cost = r.cost > 0x7fffffffLL ? 0x7fffffff : (int32_t)r.cost;
    out[t] = o;
  }
}

}  // namespace

static int epzs_launch(jmb_ctx *ctx, const jmb_epzs_req *d_reqs, int n, const int16_t *d_cands, int n_cands, jmb_epzs_res *d_res, int max_range,
                       bool any_subpel) {
  const jmb_ref &r0 = ctx->refs[ctx->ref_list[0]];
  const uint8_t *tab[JMB_MAX_REFS];
  for (int i = 0; i < JMB_MAX_REFS; i++) tab[i] = i < ctx->nref ? ctx->refs[ctx->ref_list[i]].planes : nullptr;
  int rc = jmb_reserve_dev(ctx, &ctx->d_reftab, &ctx->d_reftab_cap, sizeof(tab)); if (rc) return rc;
  JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_reftab, tab, sizeof(tab), cudaMemcpyHostToDevice, ctx->stream));
  const int blocks = (n + EG - 1) / EG;
  jmb_time_begin(ctx, JMB_K_EPZS);
  k_epzs_int<<<blocks, ET, 0, ctx->stream>>>(d_reqs, n, (const short2 *)d_cands, n_cands, d_res, ctx->cur, ctx->cur_pitch,
                                               (const uint8_t *const *)ctx->d_reftab, r0.plane_bytes, r0.pitch, ctx->cur_w, ctx->cur_h, ctx->nref,
                                               max_range, ctx->d_err);
  jmb_time_end(ctx, JMB_K_EPZS);
  JMB_LAUNCH_CHECK(ctx);
  if (any_subpel) {
    jmb_time_begin(ctx, JMB_K_REFINE);      // reported as "subpel_refine", like the refinement that follows the full search
    k_epzs_sub<<<blocks, ET, 0, ctx->stream>>>(d_reqs, n, d_res, ctx->cur, ctx->cur_pitch, (const uint8_t *const *)ctx->d_reftab, r0.plane_bytes,
                                               r0.pitch, ctx->cur_w, ctx->cur_h, ctx->me, ctx->nref);
    jmb_time_end(ctx, JMB_K_REFINE);
    JMB_LAUNCH_CHECK(ctx);
  }
  return JMB_OK;
}

static int epzs_common_checks(jmb_ctx *ctx, const char *who) {
  if (!ctx->cur || ctx->nref == 0) return jmb_fail(ctx, JMB_ERR_STATE, "%s: call jmb_pic_begin with >= 1 reference first", who);
  if (ctx->me.start_qp != 1)
    return jmb_fail(ctx, JMB_ERR_UNSUPPORTED, "%s: EPZS sub-pel needs start_me_refinement_qp = 1 (MEDistortionHPel == MEDistortionQPel); "
                    "JM itself reads next_start_pos[][-1] otherwise (me_epzs_sub.c:141,182)", who);
  return 0;
}

extern "C" {

int jmb_epzs_search(jmb_ctx *ctx, const jmb_epzs_req *reqs, int n, const int16_t *cands, int n_cands, jmb_epzs_res *res, int loc) {
  if (n <= 0) return JMB_OK;
  if (!reqs || !res || n_cands < 0 || (n_cands && !cands)) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_epzs_search: NULL buffer");
  int rc = epzs_common_checks(ctx, "jmb_epzs_search"); if (rc) return rc;
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool host = jmb_is_host(loc);
  const jmb_epzs_req *d_reqs = reqs; const int16_t *d_cands = cands; jmb_epzs_res *d_res = res;
  int max_range = 4 * JMB_MAX_SEARCH_RANGE;
  if (host) {
    max_range = 1;
    for (int i = 0; i < n; i++) max_range = max(max_range, (int)max(reqs[i].range_x, reqs[i].range_y));
    if (max_range > 4 * JMB_MAX_SEARCH_RANGE) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_epzs_search: search range %d quarter-pel", max_range);
    const size_t rb = (size_t)n * sizeof(jmb_epzs_req), cb = (size_t)max(1, n_cands) * 4, ob = (size_t)n * sizeof(jmb_epzs_res);
    rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, rb); if (rc) return rc;
    rc = jmb_reserve_dev(ctx, &ctx->d_stage2, &ctx->d_stage2_cap, cb); if (rc) return rc;
    rc = jmb_reserve_dev(ctx, &ctx->d_stage5, &ctx->d_stage5_cap, ob); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage, reqs, rb, cudaMemcpyHostToDevice, ctx->stream));
    if (n_cands) JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage2, cands, (size_t)n_cands * 4, cudaMemcpyHostToDevice, ctx->stream));
    d_reqs = (const jmb_epzs_req *)ctx->d_stage; d_cands = (const int16_t *)ctx->d_stage2; d_res = (jmb_epzs_res *)ctx->d_stage5;
  } else max_range = 4 * ctx->me.search_range;      // device-resident requests: ranges are checked on the device against the configured one
  rc = epzs_launch(ctx, d_reqs, n, d_cands, n_cands, d_res, max_range, true); if (rc) return rc;
  if (host) {
    JMB_CUDA(ctx, cudaMemcpyAsync(res, d_res, (size_t)n * sizeof(jmb_epzs_res), cudaMemcpyDeviceToHost, ctx->stream));
    if (loc == JMB_HOST) return jmb_check_device_errors(ctx);
  }
  return JMB_OK;
}

int jmb_epzs_search_frame(jmb_ctx *ctx, const jmb_mb_mvpred *pred, const int16_t *shared, int n_mb, const jmb_epzs_frame_params *fp,
                          jmb_me_res8 *res, int loc) {
  if (n_mb <= 0) return JMB_OK;
  if (!pred || !fp) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_epzs_search_frame: NULL argument");
  int rc = epzs_common_checks(ctx, "jmb_epzs_search_frame"); if (rc) return rc;
  const int mb_w = ctx->cur_w / 16, mb_total = mb_w * (ctx->cur_h / 16);
  if (n_mb > mb_total) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_epzs_search_frame: n_mb %d (picture has %d)", n_mb, mb_total);
  if (fp->n_shared < 0 || fp->n_shared > 32 || (fp->n_shared && !shared) || fp->window < 0 || fp->window > 8 || fp->range < 1 ||
      fp->range > 4 * JMB_MAX_SEARCH_RANGE || fp->ref < 0 || fp->ref >= ctx->nref || fp->pattern < 0 || fp->pattern > 5 ||
      fp->pattern_dual < 0 || fp->pattern_dual > 5)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_epzs_search_frame: n_shared %d window %d range %d ref %d patterns %d/%d", fp->n_shared, fp->window,
                    fp->range, fp->ref, fp->pattern, fp->pattern_dual);
  if (fp->mv_min_x > fp->mv_max_x || fp->mv_min_y > fp->mv_max_y || fp->mv_min_x < -32768 + fp->range || fp->mv_max_x > 32767 - fp->range ||
      fp->mv_min_y < -32768 + fp->range || fp->mv_max_y > 32767 - fp->range)
    return jmb_fail(ctx, JMB_ERR_ARG, "jmb_epzs_search_frame: mv range x %d..%d y %d..%d", fp->mv_min_x, fp->mv_max_x, fp->mv_min_y, fp->mv_max_y);
  for (int k = 0; k < 3; k++)
    if (fp->lambda[k] < 0 || fp->lambda[k] > 65535) return jmb_fail(ctx, JMB_ERR_ARG, "jmb_epzs_search_frame: lambda[%d]=%d", k, fp->lambda[k]);
  for (int t = 1; t < 8; t++)
    if (fp->medthres[t] < 0 || fp->minthres[t] < 0 || fp->maxthres[t] < 0 || fp->subthres[t] < 0)
      return jmb_fail(ctx, JMB_ERR_ARG, "jmb_epzs_search_frame: negative threshold for block type %d", t);
  JMB_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool host = jmb_is_host(loc);
  const int n = n_mb * 41, n_cands = n_mb * fp->n_shared;
  const jmb_mb_mvpred *d_pred = pred; const int16_t *d_shared = shared;
  if (host) {
    rc = jmb_reserve_dev(ctx, &ctx->d_mvpred, &ctx->d_mvpred_cap, (size_t)n_mb * sizeof(jmb_mb_mvpred)); if (rc) return rc;
    JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_mvpred, pred, (size_t)n_mb * sizeof(jmb_mb_mvpred), cudaMemcpyHostToDevice, ctx->stream));
    d_pred = (const jmb_mb_mvpred *)ctx->d_mvpred;
    if (n_cands) {
      rc = jmb_reserve_dev(ctx, &ctx->d_stage2, &ctx->d_stage2_cap, (size_t)n_cands * 4); if (rc) return rc;
      JMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage2, shared, (size_t)n_cands * 4, cudaMemcpyHostToDevice, ctx->stream));
      d_shared = (const int16_t *)ctx->d_stage2;
    }
  }
  rc = jmb_reserve_dev(ctx, &ctx->d_stage, &ctx->d_stage_cap, (size_t)n * sizeof(jmb_epzs_req)); if (rc) return rc;
  rc = jmb_reserve_dev(ctx, &ctx->d_stage5, &ctx->d_stage5_cap, (size_t)n * sizeof(jmb_epzs_res)); if (rc) return rc;
  rc = jmb_reserve_dev(ctx, &ctx->d_res_keep, &ctx->d_res_keep_cap, (size_t)n * sizeof(jmb_me_res)); if (rc) return rc;
  jmb_epzs_req *d_reqs = (jmb_epzs_req *)ctx->d_stage; jmb_epzs_res *d_eres = (jmb_epzs_res *)ctx->d_stage5;
  jmb_time_begin(ctx, JMB_K_GEN);
  k_gen_epzs<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_pred, n_mb, mb_w, *fp, d_reqs);
  jmb_time_end(ctx, JMB_K_GEN);
  JMB_LAUNCH_CHECK(ctx);
  rc = epzs_launch(ctx, d_reqs, n, d_shared, n_cands, d_eres, fp->range, (fp->flags & JMB_EPZS_SUBPEL) != 0); if (rc) return rc;
  jmb_me_res8 *d_out = res;
  if (host && res) {
    rc = jmb_reserve_dev(ctx, &ctx->d_res8, &ctx->d_res8_cap, (size_t)n * sizeof(jmb_me_res8)); if (rc) return rc;
    d_out = (jmb_me_res8 *)ctx->d_res8;
  }
  jmb_time_begin(ctx, JMB_K_GEN);
  k_epzs_pack<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_eres, n, *fp, (jmb_me_res *)ctx->d_res_keep, d_out);
  jmb_time_end(ctx, JMB_K_GEN);
  JMB_LAUNCH_CHECK(ctx);
  ctx->last_res = (const jmb_me_res *)ctx->d_res_keep; ctx->last_res_n = n;
  if (host) {
    if (res) JMB_CUDA(ctx, cudaMemcpyAsync(res, d_out, (size_t)n * sizeof(jmb_me_res8), cudaMemcpyDeviceToHost, ctx->stream));
    if (loc == JMB_HOST) return jmb_check_device_errors(ctx);
  }
  return JMB_OK;
}

}  // extern "C"
