// k_epzs.cu -- EPZS (SearchMode 3) on the device: EPZS_integer_motion_estimation (lencod/src/me_epzs_int.c:42-426) and
// EPZS_sub_pel_motion_estimation (lencod/src/me_epzs_sub.c:30-213), one warp per search.
//
// EPZS is a short, data-dependent walk: check the start mv, stop early against the previous distortions, check an ordered
// predictor list, walk a refinement pattern until its centre wins (optionally again from the second-best predictor), then a
// two-step half-/quarter-pel pattern.  10-60 distortions per block instead of the 4225 of the full search.  What is
// data-parallel in it are the distortions of one step: the warp evaluates the candidates of a step TOGETHER (a candidate's
// rows are split over 32 / n lanes and reduced by shuffle), then lane 0 replays JM's sequential selection on the complete
// distortions -- JM's early-terminated distortion returns the threshold it was given (mv_search.h:19-23), which can never win
// a strict '<', so complete sums decide alike.  JM's visited map (EPZSMap, stamped with BlkCount) is a per-warp bitmap in
// shared memory (a small hash set of the positions around the start mv; only the slots a search filled are cleared before the
// next one).  Warps are persistent: each walks requests warp, warp + #warps, ...  The sub-pel stage is a second launch.
#include "jmb_dist_dev.cuh"

namespace {

constexpr int EW = 4;          // warps per CTA
constexpr int TCAP = 384;      // touched bitmap words remembered per search (more: the whole map is cleared)

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

struct WarpS {
  short2 mv[32];
  int dist[32];
  unsigned short touched[TCAP];
};

struct Blk { RefView rv; const uint8_t *cur; int cur_pitch, pos_x, pos_y, bsx, bsy; };

__device__ __forceinline__ long long mv_cost(int lam, int vx, int vy, int px, int py) {
  return (long long)lam * (jmb_mvbits(vx - px) + jmb_mvbits(vy - py));
}

// SAD (computeSAD, me_distortion.c:349: partition-origin clamp) of the candidates ws.mv[0..n) whose bit is set in `valid`;
// a candidate's rows are split over 32 / pow2(n) lanes
__device__ __forceinline__ void eval_sad(const Blk &b, WarpS &ws, int n, unsigned valid, int lane) {
  const int p2 = n <= 1 ? 1 : 1 << (32 - __clz(n - 1));
  const int lpc = min(32 / p2, b.bsy);
  const int c = lane / lpc, sub = lane - c * lpc;
  int s = 0;
  if (c < n && ((valid >> c) & 1)) {
    const short2 v = ws.mv[c];
    const uint8_t *ref = umv(b.rv, (b.pos_y << 2) + v.y, (b.pos_x << 2) + v.x);
    const uint8_t *src = b.cur + (size_t)b.pos_y * b.cur_pitch + b.pos_x;
    for (int y = sub; y < b.bsy; y += lpc)
      for (int x = 0; x < b.bsx; x += 4)
        s += __vsadu4(*(const unsigned *)(src + (size_t)y * b.cur_pitch + x), ld4(ref + (size_t)y * b.rv.pitch + x));
  }
  for (int sh = 1; sh < lpc; sh <<= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
  if (c < n && sub == 0) ws.dist[c] = s;
  __syncwarp();
}

// distortion of the sub-pel candidates ws.mv[0..n) with the stage's metric (computeSAD / SSE / SATD incl. the 8x8 Hadamard):
// work item = (candidate, sub-block), summed with shared-memory atomics
__device__ __forceinline__ void eval_sub(const Blk &b, WarpS &ws, int n, int metric, int t8, int lane) {
  if (lane < n) ws.dist[lane] = 0;
  __syncwarp();
  const int nn = (metric == JMB_SATD && t8) ? 8 : 4, nsx = b.bsx / nn, nsub = nsx * (b.bsy / nn);
  for (int it = lane; it < n * nsub; it += 32) {
    const int c = it / nsub, sb = it - c * nsub, sbx = sb % nsx, sby = sb / nsx;
    SrcBlk src;
    load_src(src, b.cur, b.cur_pitch, b.pos_x + sbx * nn, b.pos_y + sby * nn, nn);
    const short2 v = ws.mv[c];
    atomicAdd(&ws.dist[c], subblock_dist(b.rv, src, (b.pos_x << 2) + v.x, (b.pos_y << 2) + v.y, sbx, sby, nn, metric));
  }
  __syncwarp();
}

__device__ __forceinline__ long long shfl_ll(long long v) {
  return (long long)(((unsigned long long)(unsigned)__shfl_sync(0xffffffffu, (int)((unsigned long long)v >> 32), 0) << 32) |
                     (unsigned)__shfl_sync(0xffffffffu, (int)(unsigned long long)v, 0));
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

// JM's visited map (EPZSMap) as a small open-addressing hash set per warp: a search visits at most a few hundred of the
// (2 range + 1)^2 positions, and a 4 KB table instead of a 8-33 KB bitmap lets four times as many searches be in flight.
constexpr int HS = 1024;       // slots (power of two)
struct Visited {
  unsigned *tab;               // HS keys, 0 = empty
  unsigned short *touched;     // slots filled by the current search (lane 0's bookkeeping)
  int n;
  __device__ __forceinline__ static unsigned key_of(int dx, int dy) { return (((unsigned)(dy + 2048) << 16) | (unsigned)(dx + 2048)) + 1u; }
  __device__ __forceinline__ static unsigned slot_of(unsigned k) { return (k * 2654435761u) >> 22; }
  __device__ __forceinline__ bool seen(int dx, int dy) const {
    const unsigned k = key_of(dx, dy);
    for (unsigned s = slot_of(k);; s = (s + 1) & (HS - 1)) { const unsigned v = tab[s]; if (v == k) return true; if (!v) return false; }
  }
  __device__ __forceinline__ int visit(int dx, int dy) {      // lane 0: 1 = visited before, 0 = new, -1 = table full
    const unsigned k = key_of(dx, dy);
    for (unsigned s = slot_of(k);; s = (s + 1) & (HS - 1)) {
      const unsigned v = tab[s];
      if (v == k) return 1;
      if (!v) {
        if (n >= HS - HS / 4) return -1;
        tab[s] = k;
        if (n < TCAP) touched[n] = (unsigned short)s;
        n++;
        return 0;
      }
    }
  }
};

// ---- integer stage: EPZS_integer_motion_estimation (me_epzs_int.c:42-426) ---------------------------------------------
__global__ void __launch_bounds__(EW * 32, 6)
k_epzs_int(const jmb_epzs_req *__restrict__ reqs, int n, const short2 *__restrict__ cands, int n_cands, jmb_epzs_res *__restrict__ res,
           const uint8_t *__restrict__ cur, int cur_pitch, const uint8_t *const *__restrict__ ref_planes, size_t plane_bytes, int ref_pitch,
           int w, int h, int nref, int max_range, int *__restrict__ err) {
  __shared__ WarpS wss[EW];
  __shared__ unsigned htab[EW][HS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpS &ws = wss[warp];
  Visited vis{htab[warp], ws.touched, 0};
  const long long BIG = (long long)0x7fffffff << 5;      // DISTBLK_MAX
  for (int i = lane; i < HS; i += 32) vis.tab[i] = 0;
  __syncwarp();

  for (int ri = blockIdx.x * EW + warp; ri < n; ri += gridDim.x * EW) {
    const jmb_epzs_req q = reqs[ri];
    {
      const int bad = epzs_check(q, w, h, nref, n_cands, max_range);
      if (bad) { if (lane == 0) jmb_req_report(err, bad, ri); continue; }
    }
    if (q.flags & JMB_EPZS_SKIP_INT) {      // sub-pel only: the integer-stage fields of the result echo the request
      if (lane == 0) {
        jmb_epzs_res o;
        o.mv_x = o.imv_x = q.start_x; o.mv_y = o.imv_y = q.start_y; o.cost = o.icost = q.min_mcost; o.prev_sad = q.prev_sad; o.exit_code = 0; o.n_evals = 0;
        res[ri] = o;
      }
      continue;
    }
    Blk b{RefView{ref_planes[q.ref], plane_bytes, ref_pitch, w, h}, cur, cur_pitch, q.pos_x, q.pos_y, c_bsx[q.blocktype], c_bsy[q.blocktype]};
    const int sx = q.start_x, sy = q.start_y, px = q.pred_x, py = q.pred_y, rx = q.range_x, ry = q.range_y;
    const bool gt0 = (q.flags & JMB_EPZS_REF_GT0_FRAME) != 0;
    int tx = sx, ty = sy, exit_code = 0, evals = 0, full = 0;
    long long minc, prev = q.prev_sad;
    vis.n = 0;
    auto in_range = [&](int vx, int vy) { return abs(vx - sx) <= rx && abs(vy - sy) <= ry; };
    const int lam = q.lambda[0];
    const long long ld = 2ll * lam, med = q.medthres, stop = q.stop;
    // ---- the start mv (:93-100) ----
    if (lane == 0) { ws.mv[0] = make_short2((short)sx, (short)sy); vis.visit(0, 0); }
    __syncwarp();
    eval_sad(b, ws, 1, 1u, lane);
    evals++;
    minc = mv_cost(lam, sx, sy, px, py) + ((long long)ws.dist[0] << 5);
    if (gt0 && (prev < min(med + ld, minc) || prev * 8 < minc)) exit_code = 1;                        // :103-117
    else if (minc > med + ld) {                                                                        // :121
      if (minc < (stop >> 1)) {                                                                        // :135-150
        if (q.jm_ref == 0 || prev > minc) prev = minc;
        exit_code = 2;
      } else {
        long long second = BIG;
        const long long centre_cost = minc;      // JM runs the predictor generators (and their gates) before it checks any predictor
        int check_median = 0, t2x = 0, t2y = 0;
        int off = q.cand_off;
        // ---- predictor list (:215-252) ----
        for (int s = 0; s < 4; s++) {
          const int ns = q.n_cand[s];
          const bool gen = s == 2 && (q.flags & JMB_EPZS_WINDOW_GEN);
          const bool on = q.gate[s] == 0 || centre_cost > (long long)q.gate[s] * stop;
          if (on)
            for (int i0 = 0; i0 < ns; i0 += 32) {
              const int nb = min(32, ns - i0);
              bool ok = false;
              if (lane < nb) {
                short2 v;
                if (gen) {      // window predictors around the start mv: rings of size range >> k, k descending (EPZSWindowPredictorInit)
                  const int i = i0 + lane, rings = (ns + 8) >> 3, sp = rx >> (rings - 1 - (i >> 3));
                  v = make_short2((short)(sx + c_ring[i & 7][0] * sp), (short)(sy + c_ring[i & 7][1] * sp));
                } else v = cands[off + i0 + lane];
                ws.mv[lane] = v;
                ok = in_range(v.x, v.y) && !vis.seen(v.x - sx, v.y - sy);
              }
              const unsigned valid = __ballot_sync(0xffffffffu, ok);
              __syncwarp();
              if (valid) eval_sad(b, ws, nb, valid, lane);
              evals += __popc(valid);
              if (lane == 0) {
                for (int c = 0; c < nb; c++) {
                  if (!((valid >> c) & 1)) continue;
                  const int vx = ws.mv[c].x, vy = ws.mv[c].y;
                  const int vs = vis.visit(vx - sx, vy - sy);
                  if (vs) { full |= vs < 0; continue; }
                  long long mcost = mv_cost(lam, vx, vy, px, py);
                  if (mcost < second) {
                    mcost += (long long)ws.dist[c] << 5;
                    if (mcost < minc) { t2x = tx; t2y = ty; tx = vx; ty = vy; second = minc; minc = mcost; check_median = 1; }
                    else if (mcost < second) { t2x = vx; t2y = vy; second = mcost; check_median = 1; }
                  }
                }
              }
              __syncwarp();
            }
          if (!gen) off += ns;
        }
        tx = __shfl_sync(0xffffffffu, tx, 0); ty = __shfl_sync(0xffffffffu, ty, 0);
        t2x = __shfl_sync(0xffffffffu, t2x, 0); t2y = __shfl_sync(0xffffffffu, t2y, 0);
        check_median = __shfl_sync(0xffffffffu, check_median, 0);
        minc = shfl_ll(minc);
        if (gt0 && prev * 3 < minc) exit_code = 3;                                                     // :254-273
        else if (minc > stop) {                                                                        // :279
          int pat = q.pattern, cx, cy;
          if (q.flags & JMB_EPZS_ADAPT_PATTERN) {                                                      // :286-300
            if (minc < stop + ((3 * med) >> 1))
              pat = ((tx == 0 && ty == 0) || (abs(tx - sx) < 10 && abs(ty - sy) < 10)) ? JMB_EPZS_PAT_SDIAMOND : JMB_EPZS_PAT_SQUARE;
            else if (q.flags & JMB_EPZS_SQUARE_HINT) pat = JMB_EPZS_PAT_SQUARE;
          }
          cx = tx; cy = ty;
          for (;;) {
            int pattern_stop = 0, point = 0, next_last = 0, total = c_pat_n[pat], dir = 0;
            do {                                                                                       // :307-360
              const int np = c_pat_n[pat];
              bool ok = false;
              if (lane < total) {
                int pi = point + lane; if (pi >= np) pi -= np;
                const short2 v = make_short2((short)(cx + c_pat[pat][pi][0]), (short)(cy + c_pat[pat][pi][1]));
                ws.mv[lane] = v;
                ok = in_range(v.x, v.y) && !vis.seen(v.x - sx, v.y - sy);
              }
              const unsigned valid = __ballot_sync(0xffffffffu, ok);
              __syncwarp();
              if (valid) eval_sad(b, ws, total, valid, lane);
              evals += __popc(valid);
              if (lane == 0) {
                for (int c = 0; c < total; c++) {
                  if (!((valid >> c) & 1)) continue;
                  const int vx = ws.mv[c].x, vy = ws.mv[c].y;
                  const int vs = vis.visit(vx - sx, vy - sy);
                  if (vs) { full |= vs < 0; continue; }
                  long long mcost = mv_cost(lam, vx, vy, px, py);
                  if (mcost < minc) {
                    mcost += (long long)ws.dist[c] << 5;
                    if (mcost < minc) { tx = vx; ty = vy; minc = mcost; dir = point + c; if (dir >= np) dir -= np; }
                  }
                }
              }
              __syncwarp();
              tx = __shfl_sync(0xffffffffu, tx, 0); ty = __shfl_sync(0xffffffffu, ty, 0); dir = __shfl_sync(0xffffffffu, dir, 0);
              minc = shfl_ll(minc);
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
            } while (pattern_stop != 1);
            if (gt0 && (4 * prev < minc || (3 * prev < minc && prev <= stop))) { exit_code = 4; break; }      // :362-376
            if (!(check_median && (q.jm_ref == 0 || minc < 2 * prev) && minc > ((3 * stop) >> 1) && (q.flags & JMB_EPZS_DUAL))) break;   // :379-384
            if ((tx == 0 && ty == 0) || (tx == sx && ty == sy))                                         // :391-399
              pat = (abs(tx - sx) < 10 && abs(ty - sy) < 10) ? JMB_EPZS_PAT_SDIAMOND : JMB_EPZS_PAT_SQUARE;
            else pat = q.pattern_dual;
            cx = t2x; cy = t2y;
            check_median = 0;
          }
        }
      }
    }
    if (!exit_code) { if (q.jm_ref == 0 || prev > minc) prev = minc; exit_code = 5; }                    // :409-410
    // empty the visited set for the next search of this warp
    const int nt = __shfl_sync(0xffffffffu, vis.n, 0);
    if (__shfl_sync(0xffffffffu, full, 0) && lane == 0) jmb_req_report(err, EPZS_ERR_FIELD, ri);      // (more positions than the set holds: never silently wrong)
    if (nt > TCAP) { for (int i = lane; i < HS; i += 32) vis.tab[i] = 0; }
    else for (int i = lane; i < nt; i += 32) vis.tab[ws.touched[i]] = 0;
    __syncwarp();
    if (lane == 0) {
      jmb_epzs_res o;
      o.mv_x = o.imv_x = (int16_t)tx; o.mv_y = o.imv_y = (int16_t)ty;
      o.cost = o.icost = minc; o.prev_sad = prev; o.exit_code = exit_code; o.n_evals = evals;
      res[ri] = o;
    }
  }
}

// ---- sub-pel stage: BlockMotionSearch's gate (mv_search.c:964-976), then EPZS_sub_pel_motion_estimation (me_epzs_sub.c:30-213);
// a second launch so that the Hadamard code's registers do not limit how many integer searches are in flight ----------------
__global__ void __launch_bounds__(EW * 32, 3)
k_epzs_sub(const jmb_epzs_req *__restrict__ reqs, int n, jmb_epzs_res *__restrict__ res, const uint8_t *__restrict__ cur, int cur_pitch,
           const uint8_t *const *__restrict__ ref_planes, size_t plane_bytes, int ref_pitch, int w, int h, jmb_me_config me, int nref) {
  __shared__ WarpS wss[EW];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpS &ws = wss[warp];
  const long long BIG = (long long)0x7fffffff << 5;
  for (int ri = blockIdx.x * EW + warp; ri < n; ri += gridDim.x * EW) {
    const jmb_epzs_req q = reqs[ri];
    if (!(q.flags & JMB_EPZS_SUBPEL) || q.blocktype < 1 || q.blocktype > 7 || q.ref >= nref) continue;      // (rejected requests were reported by the integer stage)
    const jmb_epzs_res r0 = res[ri];
    if (r0.exit_code == 0 && !(q.flags & JMB_EPZS_SKIP_INT)) continue;                                        // integer stage did not run: rejected
    const bool gt0 = (q.flags & JMB_EPZS_REF_GT0_FRAME) != 0;
    long long minc = r0.icost;
    if (!((q.flags & JMB_EPZS_SKIP_INT) || !gt0 || 2 * minc < 7 * r0.prev_sad)) continue;
    Blk b{RefView{ref_planes[q.ref], plane_bytes, ref_pitch, w, h}, cur, cur_pitch, q.pos_x, q.pos_y, c_bsx[q.blocktype], c_bsy[q.blocktype]};
    const int px = q.pred_x, py = q.pred_y;
    int mvx = r0.imv_x, mvy = r0.imv_y, evals = r0.n_evals;
    const int t8 = (q.flags & JMB_EPZS_TEST8X8) != 0;
    const int max_pos2 = (!me.start_hp || !me.start_qp) ? max(1, me.search_pos2) : me.search_pos2;
    int lam = q.lambda[1], best = 0, second_pos = 0;
    long long second = BIG;
    const long long sub_thr = q.subthres + 2ll * lam;
    bool done = false;
    if (!(q.flags & JMB_EPZS_SKIP_INT) && !me.start_hp) minc = BIG;
    // one batch + replay; `wide`: the first loop of a stage (second-best tracking), else the refinement loop
    auto stage = [&](int p0, int p1, int div, int metric, bool wide) {
      const int nb = p1 - p0;
      if (nb <= 0) return;
      if (lane < nb) ws.mv[lane] = make_short2((short)(mvx + c_hp[p0 + lane][0] / div), (short)(mvy + c_hp[p0 + lane][1] / div));
      __syncwarp();
      eval_sub(b, ws, nb, metric, t8, lane);
      evals += nb;
      if (lane == 0)
        for (int c = 0; c < nb; c++) {
          const int pos = p0 + c;
          long long mcost = mv_cost(lam, ws.mv[c].x, ws.mv[c].y, px, py);
          if (wide) {
            if (mcost < second) {
              mcost += (long long)ws.dist[c] << 5;
              if (mcost < minc) { second = minc; second_pos = best; minc = mcost; best = pos; }
              else if (mcost < second) { second = mcost; second_pos = pos; }
            }
          } else if (mcost < minc) {
            mcost += (long long)ws.dist[c] << 5;
            if (mcost < minc) { minc = mcost; best = pos; }
          }
        }
      __syncwarp();
      best = __shfl_sync(0xffffffffu, best, 0); second_pos = __shfl_sync(0xffffffffu, second_pos, 0);
      minc = shfl_ll(minc); second = shfl_ll(second);
    };
    stage(me.start_hp, min(5, max_pos2), 1, me.metric[1], true);                                       // me_epzs_sub.c:66-90
    if (best == 0 && px == mvx && py == mvy && minc < sub_thr) done = true;                           // :92-95
    if (!done) {
      if (me.search_pos2 >= 9 && (best != 0 || (abs(px - mvx) + abs(py - mvy))))                       // :97-122
        stage(c_ns[best][second_pos], c_ne[best][second_pos], 1, me.metric[1], false);
      if (best) { mvx += c_hp[best][0]; mvy += c_hp[best][1]; }
      const int end_pos = (minc < sub_thr) ? 1 : 5;                                                    // :135-170
      second = BIG; best = 0;      // start_me_refinement_qp == 1 (checked by the host side); second_pos carries over as in JM
      lam = q.lambda[2];
      stage(me.start_qp, end_pos, 2, me.metric[2], true);
      if (minc > sub_thr && (best != 0 || (abs(px - mvx) + abs(py - mvy))))                            // :173-200
        stage(c_ns[best][second_pos], c_ne[best][second_pos], 2, me.metric[2], false);
      if (best > 0) { mvx += c_hp[best][0] / 2; mvy += c_hp[best][1] / 2; }
    }
    if (lane == 0) { res[ri].mv_x = (int16_t)mvx; res[ri].mv_y = (int16_t)mvy; res[ri].cost = minc; res[ri].n_evals = evals; }
  }
}

// requests of a whole picture for jmb_epzs_search_frame
__global__ void k_gen_epzs(const jmb_mb_mvpred *__restrict__ pred, int n_mb, int mb_w, jmb_epzs_frame_params fp, jmb_epzs_req *__restrict__ reqs) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_mb * 41) return;
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
  reqs[t] = q;
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
  if (!ctx->epzs_grid[0]) {      // persistent warps: as many CTAs as fit the device at once
    int sms = 148, per_sm = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    JMB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_epzs_int, EW * 32, 0));
    ctx->epzs_grid[0] = sms * max(1, per_sm);
    JMB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_epzs_sub, EW * 32, 0));
    ctx->epzs_grid[1] = sms * max(1, per_sm);
  }
  const int blocks = (n + EW - 1) / EW;
  jmb_time_begin(ctx, JMB_K_EPZS);
  k_epzs_int<<<min(blocks, ctx->epzs_grid[0]), EW * 32, 0, ctx->stream>>>(d_reqs, n, (const short2 *)d_cands, n_cands, d_res, ctx->cur, ctx->cur_pitch,
                                                                           (const uint8_t *const *)ctx->d_reftab, r0.plane_bytes, r0.pitch, ctx->cur_w,
                                                                           ctx->cur_h, ctx->nref, max_range, ctx->d_err);
  jmb_time_end(ctx, JMB_K_EPZS);
  JMB_LAUNCH_CHECK(ctx);
  if (any_subpel) {
    jmb_time_begin(ctx, JMB_K_REFINE);      // reported as "subpel_refine", like the refinement that follows the full search
    k_epzs_sub<<<min(blocks, ctx->epzs_grid[1]), EW * 32, 0, ctx->stream>>>(d_reqs, n, d_res, ctx->cur, ctx->cur_pitch, (const uint8_t *const *)ctx->d_reftab,
                                                                             r0.plane_bytes, r0.pitch, ctx->cur_w, ctx->cur_h, ctx->me, ctx->nref);
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
